"""The device lensmap builder's whole pipeline on the CPU.

The text NVRTC compiles — translated lens + the fixed per-pixel kernel (`lens_source(with_kernel=True)`)
— is compiled here with g++ behind a small shim (blockIdx / threadIdx as globals, a serial launch
loop), once with the host libm and once with every libm result moved by pseudo-random ulps, the way
another libm (CUDA's) would.  The candidates it produces are merged exactly like
FisheyeHost::build_inverse_device does (risk bit -> take the interpreter's pixel), and the finished
lensmap must equal the interpreter build.  No GPU involved: this pins the kernel tail's arithmetic
(plate argmax, u/v, texel, rubix grid), the candidate encoding and the "undecided pixels go to the
host" rule against the reference-equivalent builder for every translatable lens."""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from test_transpile import TRANSLATABLE, perturbed

SHIM = r"""
#include <math.h>
#include <string.h>
#define __global__
#define __device__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(x)
#define __restrict__
struct uint3_ { unsigned x, y, z; };
static uint3_ blockIdx, blockDim, threadIdx;
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 r; r.x = x; r.y = y; return r; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p += v; return o; }
"""

RUN_INVERSE = r"""
extern "C" void run_lt_build(const LtParams *P, unsigned *cand) {
    blockDim.x = 128; blockDim.y = blockDim.z = 1;
    for (unsigned by = 0; by < (unsigned)P->height; ++by)
        for (unsigned bx = 0; bx * 128 < (unsigned)P->width; ++bx)
            for (unsigned t = 0; t < 128; ++t) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0; threadIdx.x = t;
                lt_build(*P, cand);
            }
}
"""

RUN_FORWARD = r"""
extern "C" void run_lt_forward_points(const LtParams *P, int2 *grid, unsigned char *status, unsigned *undecided, unsigned *counters, unsigned cap) {
    blockDim.x = 128; blockDim.y = blockDim.z = 1;
    const unsigned n1 = P->platesize + 1;
    for (unsigned bz = 0; bz < (unsigned)P->numplates; ++bz)
        for (unsigned by = 0; by < n1; ++by)
            for (unsigned bx = 0; bx * 128 < n1; ++bx)
                for (unsigned t = 0; t < 128; ++t) {
                    blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz; threadIdx.x = t;
                    lt_forward_points(*P, grid, status, undecided, counters, cap);
                }
}
"""


class PlateF(ctypes.Structure):
    _fields_ = [("forward", ctypes.c_float * 3), ("right", ctypes.c_float * 3), ("up", ctypes.c_float * 3), ("dist", ctypes.c_float)]


class LtParams(ctypes.Structure):  # mirrors blinky::LensBuildParams / the kernel's LtParams
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("platesize", ctypes.c_int), ("numplates", ctypes.c_int),
                ("scale", ctypes.c_double), ("rubix_block", ctypes.c_double), ("rubix_pad", ctypes.c_double),
                ("rubix_unit_px", ctypes.c_double), ("uv_dist", ctypes.c_double * 6), ("plates", PlateF * 6)]


GRID = (10, 4.0, 1.0)  # f_rubixgrid numcells cell pad (F_Init's default, fisheye.c:672)


def params_of(host, w, h, ps):
    p = LtParams()
    p.width, p.height, p.platesize, p.numplates = w, h, ps, host.numplates
    p.scale = host.scale
    numcells, cell, pad = GRID
    p.rubix_block = pad + cell
    p.rubix_pad = pad
    p.rubix_unit_px = float(ps) / (numcells * p.rubix_block + pad)
    for i, row in enumerate(host.plates()):
        for k in range(3):
            p.plates[i].forward[k], p.plates[i].right[k], p.plates[i].up[k] = row[k], row[3 + k], row[6 + k]
        p.plates[i].dist = row[10]
        p.uv_dist[i] = 0.5 / math.tan(float(np.float32(row[9]) / np.float32(2)))  # float halving, double tan (fisheye.c:2060)
    return p


def build_lib(src, run, path):
    with open(path + ".cpp", "w") as f:
        f.write(SHIM + src.replace("#include <math.h>", "") + run)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-builtin", "-shared", "-fPIC", "-o", path + ".so", path + ".cpp"],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[:3000]
    return ctypes.CDLL(path + ".so")


@pytest.mark.parametrize("scale", [0, 1, 1 << 20])  # 0 = host libm, else libm results off by up to 3*scale ulp
@pytest.mark.parametrize("lens", TRANSLATABLE)
def test_emulated_device_build_equals_interpreter_build(host, tmp_path, lens, scale):
    w, h, ps = 96, 64, 48
    host.set_rubixgrid(*GRID)
    for globe in ("cube", "tetra"):
        host.command(f"f_globe {globe}")
        host.command(f"f_lens {lens}")
        try:
            host.build_lensmap(w, h, ps, threads=1)
        except Exception:  # noqa: BLE001 — e.g. a zoom this lens cannot do on this globe
            continue
        idx, tint = host.lensmap()
        src = host.lens_source(with_kernel=True)
        if scale:
            src = perturbed(src, scale)
        lib = build_lib(src, RUN_INVERSE, str(tmp_path / f"{lens}_{globe}_{scale}"))
        p = params_of(host, w, h, ps)
        cand = np.zeros(w * h, np.uint32)
        lib.run_lt_build(ctypes.byref(p), cand.ctypes.data_as(ctypes.c_void_p))
        cand = cand.reshape(h, w)
        risk = (cand & 0x20000000) != 0
        valid = (cand & 0x80000000) != 0
        ongrid = (cand & 0x40000000) != 0
        c_idx = np.where(valid, (cand & 0x0FFFFFFF).astype(np.int64), -1)
        c_tint = np.where(valid & ~ongrid, c_idx // (ps * ps), 255)
        # the merge of FisheyeHost::build_inverse_device: undecided pixels are the interpreter's
        got_idx = np.where(risk, idx, c_idx)
        got_tint = np.where(risk, tint, c_tint)
        assert np.array_equal(got_idx, idx), (lens, globe, scale, int((got_idx != idx).sum()))
        assert np.array_equal(got_tint, tint), (lens, globe, scale)
        if scale == 0:
            # same libm: nothing should depend on the flags at all
            assert np.array_equal(c_idx, idx) and np.array_equal(c_tint, tint), (lens, globe)
        assert risk.mean() < (0.25 if scale <= 1 else 0.98), (lens, globe, scale, float(risk.mean()))


FORWARD_CHECK = ["eckert1", "gins8", "sinusoidal", "winkel2", "polyconic", "wagner6", "panini", "hammer"]


@pytest.mark.parametrize("scale", [0, 1 << 20])
@pytest.mark.parametrize("lens", FORWARD_CHECK)
def test_emulated_forward_points_equal_interpreter(host, tmp_path, lens, scale):
    """the grid-point kernel of the forward builder: screen position of every plate corner point
    (fisheye.c:2227-2243), decided points must equal lens_forward -> (int)(x/scale + W/2)"""
    w, h, ps = 120, 80, 12
    host.command("f_globe cube")
    host.command(f"f_lens {lens}")
    host.build_lensmap(w, h, ps, threads=1)  # fixes the scale
    src = host.lens_source(forward=True, with_kernel=True)
    if scale:
        src = perturbed(src, scale)
    lib = build_lib(src, RUN_FORWARD, str(tmp_path / f"{lens}_{scale}"))
    p = params_of(host, w, h, ps)
    n1 = ps + 1
    npts = host.numplates * n1 * n1
    grid = np.zeros((npts, 2), np.int32)
    status = np.zeros(npts, np.uint8)
    undecided = np.zeros(npts, np.uint32)
    counters = np.zeros(16, np.uint32)
    lib.run_lt_forward_points(ctypes.byref(p), grid.ctypes.data_as(ctypes.c_void_p), status.ctypes.data_as(ctypes.c_void_p),
                              undecided.ctypes.data_as(ctypes.c_void_p), counters.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(npts))
    plates = host.plates()
    decided = 0
    for pt in range(npts):
        if status[pt] == 2:
            continue
        i, j, plate = pt % n1, pt // n1 % n1, pt // n1 // n1
        # plate_uv_to_ray in float32 (fisheye.c:1198-1214)
        f, r, u = (plates[plate][k:k + 3].astype(np.float32) for k in (0, 3, 6))
        uu = np.float32((i - 0.5) / ps - 0.5)
        vv = np.float32(-((j - 0.5) / ps - 0.5))
        ray = np.float32(plates[plate][10]) * f
        ray = ray + uu * r
        ray = ray + vv * u
        ln = np.float32(math.sqrt(float(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2])))
        if ln:
            ray = ray * (np.float32(1) / ln)
        st, (x, y) = host.lens_forward(float(ray[0]), float(ray[1]), float(ray[2]))
        assert st == int(status[pt]), (lens, scale, pt, st, int(status[pt]))
        decided += 1
        if st == 1:
            want = (int(x / host.scale + w // 2), int(-y / host.scale + h // 2))
            assert want == (int(grid[pt][0]), int(grid[pt][1])), (lens, scale, pt, want, grid[pt].tolist())
    assert decided > 0.3 * npts, (lens, scale, decided, npts)
    assert int(counters[0]) == int((status == 2).sum())


# ----------------------------------------------------------------------------- forward builder, steps 2-4

FWD_HARNESS = r"""
#include <vector>
#include <cstring>
#include "forward_raster.h"
using namespace blinky;
extern "C" int fwd_sizeof_geom() { return (int)sizeof(FwdGeom); }
// patches -> stale replay -> rasterise every texel in the order given -> resolve
extern "C" void fwd_run(const FwdGeom *g, FwdPoint *grid, unsigned char *status, const ForwardPatch *patches, unsigned npatch,
                        int any_nil, const unsigned *order, unsigned ntexels, int32_t *idx, uint8_t *tint, int *display,
                        FwdMessage *messages, unsigned *nmsg) {
    for (unsigned k = 0; k < npatch; ++k) fwd_apply_patch(grid, status, patches[k]);
    if (any_nil)
        for (int t = 2 * (g->ps + 1) - 1; t >= 0; --t) fwd_stale_chain(grid, status, g->ps, g->numplates, t);
    const size_t npix = (size_t)g->width * g->height;
    std::vector<unsigned> keys(2 * npix, 0u);
    unsigned counters[16];
    memset(counters, 0, sizeof counters);
    FwdOut o{keys.data(), keys.data() + npix, counters, messages};
    for (unsigned k = 0; k < ntexels; ++k) {
        const unsigned t = order[k];
        const int px = t % g->ps, py = t / g->ps % g->ps, plate = t / g->ps / g->ps;
        fwd_raster_texel(*g, grid, o, plate, py, px);
    }
    for (size_t at = 0; at < npix; ++at) fwd_resolve_pixel(keys.data(), keys.data() + npix, idx, tint, at, g->ps);
    for (int i = 0; i < 6; ++i) display[i] = counters[3 + i] ? 1 : 0;
    *nmsg = counters[2];
}
"""


class FwdGeom(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int), ("height", ctypes.c_int), ("ps", ctypes.c_int), ("numplates", ctypes.c_int),
                ("rubix_block", ctypes.c_double), ("rubix_pad", ctypes.c_double), ("rubix_unit_px", ctypes.c_double),
                ("plates", PlateF * 6)]


class ForwardPatch(ctypes.Structure):
    _fields_ = [("point", ctypes.c_uint32), ("status", ctypes.c_int32), ("lx", ctypes.c_int32), ("ly", ctypes.c_int32)]


@pytest.fixture(scope="module")
def fwd_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("fwd")
    src = d / "fwd_harness.cpp"
    src.write_text(FWD_HARNESS)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(root, "blinky_b200", "csrc"),
                        "-o", str(d / "fwd_harness.so"), str(src)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[:3000]
    lib = ctypes.CDLL(str(d / "fwd_harness.so"))
    assert lib.fwd_sizeof_geom() == ctypes.sizeof(FwdGeom)
    return lib


def x86_int(v):
    """(int) of a double as cvttsd2si does it: out of range / NaN -> INT_MIN"""
    if not math.isfinite(v) or abs(v) >= 2147483648.0:
        return -2147483648
    return int(v)


HOLES = """
map = "lens_forward"
max_fov = 360
max_vfov = 180
lens_width = 2*pi
lens_height = pi
onload = "f_contain"
function lens_forward(x, y, z)
  local lat, lon = ray_to_latlon(x, y, z)
  if lat > 0.9 or (lon > 0.5 and lon < 0.7) or x*x < 0.0004 then
    return nil
  end
  return lon, lat
end
"""


@pytest.mark.parametrize("scale", [0, 1 << 20])
@pytest.mark.parametrize("lens", ["eckert1", "sinusoidal", "winkel2", "polyconic", "larrivee", "holes"])
def test_emulated_forward_build_equals_serial_builder_in_any_thread_order(host, fwd_lib, tmp_path, lens, scale):
    """grid points (NVRTC text behind the shim, exact or perturbed libm) -> undecided points from the
    interpreter -> stale-slot replay -> quads rasterised in RANDOM texel orders with max-key writes ->
    resolve: must equal the reference-equivalent serial scanline builder, including the display flags
    and the order of its "> maxdiff" console messages"""
    host.set_rubixgrid(*GRID)
    for globe, (w, h, ps) in (("cube", (128, 80, 20)), ("trism", (97, 60, 14))):
        host.command(f"f_globe {globe}")
        if lens == "holes":
            host.load_lens("holes", HOLES)
        else:
            host.command(f"f_lens {lens}")
        host.clear_log()
        host.build_lensmap(w, h, ps, threads=1)
        want_idx, want_tint = host.lensmap()
        want_disp = host.display()
        want_msgs = [int(l.split()[0]) for l in host.log.splitlines() if l.endswith("> maxdiff")]

        src = host.lens_source(forward=True, with_kernel=True)
        if scale:
            src = perturbed(src, scale)
        lib = build_lib(src, RUN_FORWARD, str(tmp_path / f"{lens}_{globe}_{scale}"))
        p = params_of(host, w, h, ps)
        n1 = ps + 1
        P = host.numplates
        npts = P * n1 * n1
        grid = np.zeros((npts, 2), np.int32)
        status = np.zeros(npts, np.uint8)
        undecided = np.zeros(npts, np.uint32)
        counters = np.zeros(16, np.uint32)
        lib.run_lt_forward_points(ctypes.byref(p), grid.ctypes.data_as(ctypes.c_void_p), status.ctypes.data_as(ctypes.c_void_p),
                                  undecided.ctypes.data_as(ctypes.c_void_p), counters.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(npts))
        # the interpreter settles the undecided points (FisheyeHost::build_forward_device)
        plates = host.plates()
        und = undecided[: int(counters[0])]
        patches = (ForwardPatch * max(1, len(und)))()
        for k, pt in enumerate(und.tolist()):
            i, j, plate = pt % n1, pt // n1 % n1, pt // n1 // n1
            f, r, u = (plates[plate][a:a + 3].astype(np.float32) for a in (0, 3, 6))
            uu = np.float32((i - 0.5) / ps - 0.5)
            vv = np.float32(-((j - 0.5) / ps - 0.5))
            ray = np.float32(plates[plate][10]) * f
            ray = ray + uu * r
            ray = ray + vv * u
            ln = np.float32(math.sqrt(float(ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2])))
            if ln:
                ray = ray * (np.float32(1) / ln)
            st, (x, y) = host.lens_forward(float(ray[0]), float(ray[1]), float(ray[2]))
            patches[k].point, patches[k].status = pt, st
            if st == 1:
                patches[k].lx, patches[k].ly = x86_int(x / host.scale + w // 2), x86_int(-y / host.scale + h // 2)
        any_nil = int(counters[1] > 0 or any(patches[k].status != 1 for k in range(len(und))))

        g = FwdGeom()
        g.width, g.height, g.ps, g.numplates = w, h, ps, P
        g.rubix_block, g.rubix_pad, g.rubix_unit_px = p.rubix_block, p.rubix_pad, p.rubix_unit_px
        for i in range(6):
            g.plates[i] = p.plates[i]
        ntex = P * ps * ps
        rng = np.random.default_rng(5)
        for order in (np.arange(ntex), np.arange(ntex)[::-1].copy(), rng.permutation(ntex), rng.permutation(ntex)):
            order = order.astype(np.uint32)
            gcopy, scopy = grid.copy(), status.copy()
            idx = np.zeros(w * h, np.int32)
            tint = np.zeros(w * h, np.uint8)
            disp = (ctypes.c_int * 6)()
            msgs = np.zeros((4096, 2), np.uint32)
            nmsg = ctypes.c_uint()
            fwd_lib.fwd_run(ctypes.byref(g), gcopy.ctypes.data_as(ctypes.c_void_p), scopy.ctypes.data_as(ctypes.c_void_p), patches,
                            ctypes.c_uint(len(und)), any_nil, order.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(ntex),
                            idx.ctypes.data_as(ctypes.c_void_p), tint.ctypes.data_as(ctypes.c_void_p), disp,
                            msgs.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nmsg))
            assert np.array_equal(idx.reshape(h, w), want_idx), (lens, globe, scale, int((idx.reshape(h, w) != want_idx).sum()))
            assert np.array_equal(tint.reshape(h, w), want_tint), (lens, globe, scale)
            assert list(disp)[:P] == want_disp[:P], (lens, globe)
            got_msgs = [int(v) for _, v in sorted(map(tuple, msgs[: nmsg.value].tolist()))]
            assert got_msgs == want_msgs, (lens, globe, scale)
