"""The tile plan is the lensmap's layout in HBM — what the warp kernels actually read.
Here it is interpreted on the CPU, tile by tile the way the ring kernel does (source box cut out of the
faces with zero fill outside the plate; 16-bit entries in the kernel's lane order indexing into the
box, one tint per tile applied where the tile's flag words say so; 32-bit entries for gather tiles; background for empty tiles and
unmapped pixels; entry blocks addressed by tile index alone), and the result must be the reference's
render_lensmap.  This pins the planner and the layout contract without a GPU."""
import numpy as np
import pytest

EMPTY, BOX, GATHER, BOX_FULL = 0, 1, 2, 3


BOX_BLOCK, GATHER_BLOCK = 2048 + 128, 4096


def lane_pixel(lane, i):
    """pixel i (0..31) of lane `lane` in a BOX tile: 8 quads of 4 consecutive pixels, 4 rows apart"""
    return (lane >> 3) + 4 * (i >> 2), 4 * (lane & 7) + (i & 3)


def render_from_plan(fe, faces, palmaps, bg, rubix, max_box=8192):
    tiles, entries = fe.tile_plan()
    W, H, ps, P = fe.width, fe.height, fe.platesize, fe.numplates
    faces = faces.reshape(P, ps, ps)
    out = bg.copy()
    covered = np.zeros((H, W), bool)
    lut = np.concatenate([palmaps, np.arange(256, dtype=np.uint8)[None], np.arange(256, dtype=np.uint8)[None]])  # tint 6/7 -> identity
    types = tiles["type"] & 3
    is_box = np.isin(types, (BOX, BOX_FULL))
    nbox, ngather = int(is_box.sum()), int((types == GATHER).sum())
    # class order: BOX*, GATHER*, EMPTY* — block addresses then follow from the index alone
    assert is_box[:nbox].all() and (types[nbox:nbox + ngather] == GATHER).all() and (types[nbox + ngather:] == EMPTY).all()
    assert entries.size >= nbox * BOX_BLOCK + ngather * GATHER_BLOCK
    lanes = np.arange(32)
    shapes = {}
    for n, t in enumerate(tiles):
        x0, y0 = int(t["px"]), int(t["py"])
        assert x0 % 32 == 0 and y0 % 32 == 0 and x0 < W and y0 < H
        ys, xs = min(32, H - y0), min(32, W - x0)
        assert not covered[y0:y0 + ys, x0:x0 + xs].any(), "tiles overlap"
        covered[y0:y0 + ys, x0:x0 + xs] = True
        ty = int(t["type"]) & 3          # upper six bits: index of the box shape (one TMA descriptor per shape)
        if ty in (BOX, BOX_FULL):
            shape = int(t["type"]) >> 2
            shapes.setdefault(shape, (int(t["box_w16"]), int(t["box_h8"])))
            assert shapes[shape] == (int(t["box_w16"]), int(t["box_h8"])) and shape < 64
            bw, bh = int(t["box_w16"]) * 16, int(t["box_h8"]) * 8
            bx, by, plate, tile_tint = int(t["box_x"]), int(t["box_y"]), int(t["plate"]) & 7, (int(t["plate"]) >> 3) & 7
            assert 16 <= bw <= 256 and 8 <= bh <= 256 and bw * bh <= max_box and bx % 16 == 0  # TMA constraints
            assert int(t["entry_offset"]) == n * BOX_BLOCK
            box = np.zeros((bh, bw), np.uint8)  # TMA zero-fills what lies outside the tensor
            sy0, sy1 = max(by, 0), min(by + bh, ps)
            sx0, sx1 = max(bx, 0), min(bx + bw, ps)
            if sy1 > sy0 and sx1 > sx0:
                box[sy0 - by:sy1 - by, sx0 - bx:sx1 - bx] = faces[plate, sy0:sy1, sx0:sx1]
            blk = entries[n * BOX_BLOCK:(n + 1) * BOX_BLOCK]
            ent = blk[:2048].view("<u2").reshape(4, 32, 8)     # [load k][lane][j]: pixel i = 8k + j
            flags = blk[2048:].view("<u4")                     # [lane]: bit i = the lane's pixel i carries the tile's tint
            assert tile_tint <= 5 or (tile_tint == 7 and not flags.any())
            e = np.zeros((32, 32), np.uint16)
            tint = np.zeros((32, 32), np.int64)
            for i in range(32):
                r, c = lane_pixel(lanes, i)
                e[r, c] = ent[i >> 3, lanes, i & 7]
                tint[r, c] = np.where((flags >> i) & 1, tile_tint, 6)
            valid = (e & 0x8000) != 0
            if ty == BOX_FULL:
                assert valid.all() and ys == 32 and xs == 32
            off = (e & 0x3FFF).astype(np.int64)
            assert (off[valid] < bw * bh).all() and (tint <= 6).all() and (tint[~valid] == 6).all()
            px = box.reshape(-1)[np.where(valid, off, 0)]
        else:
            if ty == EMPTY:
                continue
            assert ty == GATHER and int(t["entry_offset"]) == nbox * BOX_BLOCK + (n - nbox) * GATHER_BLOCK
            e = entries[int(t["entry_offset"]):int(t["entry_offset"]) + 4096].view("<u4").reshape(32, 32)
            valid = (e & 0x80000000) != 0
            px = faces.reshape(-1)[np.where(valid, e & 0x0FFFFFFF, 0).astype(np.int64)]
            tint = ((e >> 28) & 7).astype(np.int64)
        if rubix:
            px = lut[np.minimum(tint, 7), px]
        assert not valid[ys:, :].any() and not valid[:, xs:].any(), "entries beyond the frame edge must be unmapped"
        sub = out[y0:y0 + ys, x0:x0 + xs]
        sub[valid[:ys, :xs]] = px[:ys, :xs][valid[:ys, :xs]]
    assert covered.all(), "every pixel belongs to exactly one tile"
    return out


CASES = [
    ("cube", "panini", "f_fov 180", (640, 480, 256)),       # BASELINE C1
    ("cube", "quincuncial", "f_cover", (333, 201, 128)),    # ragged edges, many gather tiles
    ("tetra", "stereographic", "f_fov 200", (257, 131, 96)),
    ("cube", "fisheye1", "f_contain", (320, 200, 208)),     # empty tiles, minification
    ("trism", "hammer", "f_contain", (256, 160, 64)),
    ("cube", "panini", "f_fov 170", (200, 120, 100)),       # platesize not a multiple of 16: no BOX tiles allowed
    ("cube", "sinusoidal", "f_contain", (300, 150, 96)),    # forward-built map
]


@pytest.mark.parametrize("rubix", [False, True])
@pytest.mark.parametrize("globe,lens,zoom,size", CASES)
def test_plan_interpreted_on_the_cpu_equals_reference_render(bb, host, restate, palette, globe, lens, zoom, size, rubix):
    w, h, ps = size
    host.command(f"f_globe {globe}")
    host.command(f"f_lens {lens}")
    host.command(zoom)
    host.set_rubix(rubix)
    host.build_lensmap(w, h, ps, threads=2)
    idx, tint = host.lensmap()
    faces = bb.synthetic_faces(host.numplates, ps, 5)
    bg = bb.synthetic_background(w, h)
    pm = restate.palmaps(palette)
    want = restate.render(idx, tint, faces, pm, rubix, background=bg)
    got = render_from_plan(host, faces, pm, bg, rubix)
    assert np.array_equal(got, want), (globe, lens, int((got != want).sum()))
    tiles, _ = host.tile_plan()
    if ps % 16:
        assert not np.isin(tiles["type"] & 3, (BOX, BOX_FULL)).any()
    elif lens == "panini":
        assert np.isin(tiles["type"] & 3, (BOX, BOX_FULL)).mean() > 0.7  # the point of the layout
