#include "tile_plan.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/blinky_b200.h"
#include "parallel.h"

namespace blinky {

namespace {

// one row of tiles; entry offsets are relative to the row's own entry buffer
void plan_tile_row(const uint32_t *packed, int width, int height, int platesize, bool allow_box, bool odd_pitch, int ty, TilePlan &plan) {
    const uint32_t ps = static_cast<uint32_t>(platesize);
    const uint32_t ps2 = ps * ps;
    std::vector<uint32_t> tile(kTilePixels);
    {
        for (int tx = 0; tx < plan.tiles_x; ++tx) {
            const int x0 = tx * kTileW, y0 = ty * kTileH;
            // collect the tile (pixels beyond the frame edge are unmapped)
            bool any = false, one_plate = true;
            int nvalid = 0;
            uint32_t plate = 0;
            uint32_t minx = ~0u, miny = ~0u, maxx = 0, maxy = 0;
            for (int r = 0; r < kTileH; ++r) {
                for (int c = 0; c < kTileW; ++c) {
                    uint32_t e = 0;
                    if (y0 + r < height && x0 + c < width) e = packed[static_cast<size_t>(y0 + r) * width + x0 + c];
                    tile[static_cast<size_t>(r) * kTileW + c] = e;
                    if (!(e & BLINKY_LM_VALID)) continue;
                    ++nvalid;
                    const uint32_t idx = e & BLINKY_LM_INDEX_MASK;
                    const uint32_t p = idx / ps2, rem = idx % ps2;
                    const uint32_t py = rem / ps, px = rem % ps;
                    if (!any) {
                        any = true;
                        plate = p;
                    } else if (p != plate) {
                        one_plate = false;
                    }
                    minx = std::min(minx, px);
                    maxx = std::max(maxx, px);
                    miny = std::min(miny, py);
                    maxy = std::max(maxy, py);
                }
            }
            TileDesc d;
            memset(&d, 0, sizeof d);
            d.px = static_cast<uint16_t>(x0);
            d.py = static_cast<uint16_t>(y0);
            if (!any) {
                d.type = TILE_EMPTY;
                ++plan.n_empty;
                plan.tiles.push_back(d);
                continue;
            }
            bool box = allow_box && one_plate;
            uint32_t bw = 0, bh = 0;
            if (box) {
                // TMA faults ("illegal instruction") unless the innermost coordinate is a
                // multiple of 16 bytes (measured on B200, scripts/tma_probe.cu): start the
                // box on a 16-texel boundary.  The row coordinate is unconstrained.
                minx &= ~15u;
                bw = ((maxx - minx + 1) + 15) / 16 * 16;
                bh = ((maxy - miny + 1) + 7) / 8 * 8;
                if (bw > 128 || bh > 256 || bw * bh > static_cast<uint32_t>(kMaxBoxBytes)) box = false;
                // The box pitch in shared memory is bw bytes.  With an even number of 16-byte columns
                // (pitch 32/64/96/128 B) source rows one or two apart start on the same banks, and a
                // consumer warp (4 tile rows) reads several source rows at the same x: bank conflicts.
                // One more column (read from L2, never referenced) makes the pitch an odd multiple of
                // 16 B, which staggers 8 consecutive rows over the banks.  (Optional: it did not pay.)
                if (box && odd_pitch && (bw / 16) % 2 == 0 && bw + 16 <= 128 && (bw + 16) * bh <= static_cast<uint32_t>(kMaxBoxBytes)) bw += 16;
            }
            // entry blocks start 16-byte aligned
            plan.entries.resize((plan.entries.size() + 15) / 16 * 16);
            d.entry_offset = static_cast<uint32_t>(plan.entries.size());
            if (box) {
                d.type = nvalid == kTilePixels ? TILE_BOX_FULL : TILE_BOX;
                if (nvalid == kTilePixels) ++plan.n_box_full;
                d.plate = static_cast<uint8_t>(plate);
                d.box_x = static_cast<int16_t>(minx);
                d.box_y = static_cast<int16_t>(miny);
                d.box_w16 = static_cast<uint8_t>(bw / 16);
                d.box_h8 = static_cast<uint8_t>(bh / 8);
                plan.entries.resize(plan.entries.size() + kTilePixels * 2);
                uint16_t *out = reinterpret_cast<uint16_t *>(plan.entries.data() + d.entry_offset);
                for (int i = 0; i < kTilePixels; ++i) {
                    const uint32_t e = tile[static_cast<size_t>(i)];
                    if (!(e & BLINKY_LM_VALID)) {
                        out[i] = static_cast<uint16_t>(BLINKY_LM_TINT_NONE << kBoxTintShift);
                        continue;
                    }
                    const uint32_t rem = (e & BLINKY_LM_INDEX_MASK) % ps2;
                    const uint32_t py = rem / ps, px = rem % ps;
                    const uint32_t off = (py - miny) * bw + (px - minx);
                    const uint32_t tint = (e >> BLINKY_LM_TINT_SHIFT) & 7u;
                    out[i] = static_cast<uint16_t>(kBoxValid | (tint << kBoxTintShift) | off);
                }
                const uint16_t shape = static_cast<uint16_t>((d.box_w16 << 8) | d.box_h8);
                if (std::find(plan.shapes.begin(), plan.shapes.end(), shape) == plan.shapes.end()) plan.shapes.push_back(shape);
                plan.box_bytes += static_cast<uint64_t>(bw) * bh;
                ++plan.n_box;
            } else {
                d.type = TILE_GATHER;
                plan.entries.resize(plan.entries.size() + kTilePixels * 4);
                memcpy(plan.entries.data() + d.entry_offset, tile.data(), kTilePixels * 4);
                ++plan.n_gather;
            }
            plan.tiles.push_back(d);
        }
    }
}

}  // namespace

TilePlan make_tile_plan(const uint32_t *packed, int width, int height, int platesize, bool allow_box, int threads) {
    TilePlan plan;
    // BLINKY_BOX_PITCH=odd pads boxes to odd multiples of 16 bytes (see plan_tile_row).  Measured on
    // B200 (scripts/pitch_ab.sh, profiles/r1e_box_pitch_ab.txt): no difference within noise on any
    // BASELINE lens while staging 13 % more bytes, so the tightest box stays the default.
    const char *pitch_env = getenv("BLINKY_BOX_PITCH");
    const bool odd_pitch = pitch_env && strcmp(pitch_env, "odd") == 0;
    plan.width = width;
    plan.height = height;
    plan.platesize = platesize;
    plan.tiles_x = (width + kTileW - 1) / kTileW;
    plan.tiles_y = (height + kTileH - 1) / kTileH;
    std::vector<TilePlan> rows(static_cast<size_t>(plan.tiles_y));
    parallel_for(plan.tiles_y, threads, [&](int ty) {
        TilePlan &r = rows[static_cast<size_t>(ty)];
        r.tiles_x = plan.tiles_x;
        r.tiles.reserve(static_cast<size_t>(plan.tiles_x));
        plan_tile_row(packed, width, height, platesize, allow_box, odd_pitch, ty, r);
    });
    // stitch the rows together in order (every entry block is a multiple of 16 bytes)
    size_t total = 0;
    for (const TilePlan &r : rows) total += r.entries.size();
    plan.entries.resize(total + 16);
    plan.tiles.reserve(static_cast<size_t>(plan.tiles_x) * plan.tiles_y);
    size_t base = 0;
    for (const TilePlan &r : rows) {
        if (!r.entries.empty()) memcpy(plan.entries.data() + base, r.entries.data(), r.entries.size());
        for (TileDesc d : r.tiles) {
            if (d.type != TILE_EMPTY) d.entry_offset += static_cast<uint32_t>(base);
            plan.tiles.push_back(d);
        }
        base += r.entries.size();
        plan.n_empty += r.n_empty;
        plan.n_box += r.n_box;
        plan.n_box_full += r.n_box_full;
        plan.n_gather += r.n_gather;
        plan.box_bytes += r.box_bytes;
        for (uint16_t shape : r.shapes)
            if (std::find(plan.shapes.begin(), plan.shapes.end(), shape) == plan.shapes.end()) plan.shapes.push_back(shape);
    }
    // BOX tiles first: the TMA ring kernel walks [0, n_box), the gather kernel the rest
    std::stable_partition(plan.tiles.begin(), plan.tiles.end(),
                          [](const TileDesc &d) { return d.type == TILE_BOX || d.type == TILE_BOX_FULL; });
    return plan;
}

}  // namespace blinky
