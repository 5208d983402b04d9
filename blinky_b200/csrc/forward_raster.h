// Forward lensmap builder, steps 2-4: the per-thread bodies of the static kernels in
// lens_device.cu, written host/device so that the CPU test-suite can execute them (in any thread
// order) against the serial host builder.  Reference: fisheye.c:2126-2338 (resume_lensmap_forward,
// draw_quad), :1963-1982 (set_lensmap_from_plate), :1922-1960 (rubix grid).
//
// Everything here is integer or IEEE float/double arithmetic in the host's operation order; the
// translation unit that includes it must be compiled without FMA contraction.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>

#include "lens_device.h"

#if defined(__CUDACC__)
#define FWD_HD __host__ __device__ __forceinline__
#else
#define FWD_HD inline
#endif

namespace blinky {

struct FwdPoint {  // screen position of one plate grid point (int2 on the device)
    int x, y;
};
struct FwdMessage {  // one "%d > maxdiff" the reference prints, with the writer key that orders it
    unsigned key, value;
};
constexpr unsigned kFwdMessageCap = 4096;

struct FwdGeom {
    int width, height, ps, numplates;
    double rubix_block, rubix_pad, rubix_unit_px;
    LensBuildParams::PlateF plates[6];
};

struct FwdOut {
    unsigned *idxkey, *tintkey;   // [W*H] highest writer key (+1) overall / among off-grid writers
    unsigned *counters;           // [2] message count, [3..8] display flags
    FwdMessage *messages;
};

FWD_HD unsigned fwd_atomic_max(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__)
    return atomicMax(p, v);
#else
    const unsigned o = *p;
    if (v > o) *p = v;
    return o;
#endif
}
FWD_HD unsigned fwd_atomic_inc(unsigned *p) {
#if defined(__CUDA_ARCH__)
    return atomicAdd(p, 1u);
#else
    return (*p)++;
#endif
}

FWD_HD void fwd_apply_patch(FwdPoint *grid, unsigned char *status, const ForwardPatch &pt) {
    status[pt.point] = static_cast<unsigned char>(pt.status);
    if (pt.status == 1) {
        grid[pt.point].x = pt.lx;
        grid[pt.point].y = pt.ly;
    }
}

// The reference keeps two row buffers and `continue`s over nil results (fisheye.c:2151-2189), so a
// nil slot shows whatever the buffer held before: the row two steps earlier, the previous plate's last
// rows at a plate start, zero at the very beginning.  Thread t = (column t/2, buffer t%2) replays its chain.
FWD_HD void fwd_stale_chain(FwdPoint *grid, const unsigned char *status, int ps, int numplates, int t) {
    const int n1 = ps + 1;
    const int i = t >> 1;
    int j = (t & 1) ? ps - 1 : ps;  // buffer `bot` starts with row ps, buffer `top` with row ps-1
    FwdPoint last;
    last.x = last.y = 0;
    for (int p = 0; p < numplates;) {
        const size_t row = (static_cast<size_t>(p) * n1 + j) * n1;
        // slot 1 is skipped together with slot 0 (the `continue` in the px == 0 branch)
        const bool valid = status[row + i] == 1 && !(i == 1 && status[row] != 1);
        if (valid) last = grid[row + i];
        else grid[row + i] = last;
        if (j >= 2) {
            j -= 2;
        } else {
            j = j == 1 ? ps : ps - 1;  // the buffer that ended as `bot` (row 1) takes row ps of the next plate
            ++p;
        }
    }
}

FWD_HD float fwd_dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

FWD_HD void fwd_set(const FwdGeom &g, const FwdOut &o, int lx, int ly, unsigned key, bool ongrid, int plate) {
    if (lx < 0 || lx >= g.width || ly < 0 || ly >= g.height) return;  // set_lensmap_from_plate's screen check
    o.counters[3 + plate] = 1u;                                         // display flag (benign race: all writers store 1)
    const size_t at = static_cast<size_t>(lx) + static_cast<size_t>(ly) * g.width;
    fwd_atomic_max(&o.idxkey[at], key);
    if (!ongrid) fwd_atomic_max(&o.tintkey[at], key);
}

// draw_quad (fisheye.c:2246-2338) for the texel (plate, px, py); key orders the writers like the
// reference's loops do (plate ascending, py descending, px ascending): the highest key wins.
FWD_HD void fwd_raster_texel(const FwdGeom &g, const FwdPoint *grid, const FwdOut &o, int plate, int py, int px) {
    const int ps = g.ps, n1 = ps + 1;
    // the texel belongs to this plate only if the plate wins the ray's argmax (:2193-2199)
    {
        const LensBuildParams::PlateF &P = g.plates[plate];
        double u = static_cast<double>(px) / ps, v = static_cast<double>(py) / ps;
        u -= 0.5;
        v -= 0.5;
        v = -v;
        float r[3] = {0.0f, 0.0f, 0.0f};
        const float fu = static_cast<float>(u), fv = static_cast<float>(v);
        for (int k = 0; k < 3; ++k) r[k] = r[k] + P.dist * P.forward[k];
        for (int k = 0; k < 3; ++k) r[k] = r[k] + fu * P.right[k];
        for (int k = 0; k < 3; ++k) r[k] = r[k] + fv * P.up[k];
        float len = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
        len = static_cast<float>(sqrt(static_cast<double>(len)));
        if (len) {
            const float inv = 1 / len;
            r[0] *= inv;
            r[1] *= inv;
            r[2] *= inv;
        }
        int best = 0;
        double best_dp = -2;
        for (int k = 0; k < g.numplates; ++k) {
            const double dp = static_cast<double>(fwd_dot3(r, g.plates[k].forward));
            if (dp > best_dp) {
                best_dp = dp;
                best = k;
            }
        }
        if (best != plate) return;
    }
    const unsigned key = (static_cast<unsigned>(plate) * ps + (ps - 1 - py)) * ps + px + 1u;
    const double ux = static_cast<double>(px) / g.rubix_unit_px, uy = static_cast<double>(py) / g.rubix_unit_px;
    const bool ongrid = fmod(ux, g.rubix_block) < g.rubix_pad || fmod(uy, g.rubix_block) < g.rubix_pad;

    const size_t top = (static_cast<size_t>(plate) * n1 + py) * n1, bot = top + n1;
    const FwdPoint c0 = grid[top + px], c1 = grid[top + px + 1], c2 = grid[bot + px + 1], c3 = grid[bot + px];  // tl, tr, br, bl: clockwise
    const int cx[4] = {c0.x, c1.x, c2.x, c3.x}, cy[4] = {c0.y, c1.y, c2.y, c3.y};
    int x = cx[0], y = cy[0];
    int minx = x, maxx = x, miny = y, maxy = y;
    for (int i = 1; i < 4; ++i) {
        if (cx[i] < minx) minx = cx[i]; else if (cx[i] > maxx) maxx = cx[i];
        if (cy[i] < miny) miny = cy[i]; else if (cy[i] > maxy) maxy = cy[i];
    }
    const int maxdiff = 20;
    // abs() of an int difference, computed like the host does (wraps the same way on overflow)
    const int ddx = static_cast<int>(static_cast<unsigned>(minx) - static_cast<unsigned>(maxx));
    const int ddy = static_cast<int>(static_cast<unsigned>(miny) - static_cast<unsigned>(maxy));
    if ((ddx < 0 ? -ddx : ddx) > maxdiff || (ddy < 0 ? -ddy : ddy) > maxdiff) return;
    if (miny == maxy && minx == maxx) {
        fwd_set(g, o, x, y, key, ongrid, plate);
        return;
    }
    if (miny == maxy) {
        for (int tx = minx; tx <= maxx; ++tx) fwd_set(g, o, tx, miny, key, ongrid, plate);
        return;
    }
    if (minx == maxx) {
        for (int ty = miny; ty <= maxy; ++ty) fwd_set(g, o, x, ty, key, ongrid, plate);
        return;
    }
    for (y = miny; y <= maxy; ++y) {
        int tx[2] = {minx, maxx};
        int found = 0;
        int j = 3;
        for (int i = 0; i < 4; ++i) {
            const int ix = cx[i], iy = cy[i], jx = cx[j], jy = cy[j];
            if ((iy < y && y <= jy) || (jy < y && y <= iy)) {
                const double dy = jy - iy;
                const double dx = jx - ix;
                tx[found] = static_cast<int>(ix + (y - iy) / dy * dx);
                if (++found == 2) break;
            }
            j = i;
        }
        if (tx[0] > tx[1]) {
            const int t = tx[0];
            tx[0] = tx[1];
            tx[1] = t;
        }
        if (tx[1] - tx[0] > maxdiff) {
            const unsigned at = fwd_atomic_inc(&o.counters[2]);
            if (at < kFwdMessageCap) {
                o.messages[at].key = key;
                o.messages[at].value = static_cast<unsigned>(tx[1] - tx[0]);
            }
            return;
        }
        for (x = tx[0]; x <= tx[1]; ++x) fwd_set(g, o, x, y, key, ongrid, plate);
    }
}

FWD_HD void fwd_resolve_pixel(const unsigned *idxkey, const unsigned *tintkey, int32_t *idx, uint8_t *tint, size_t at, int ps) {
    const unsigned k = idxkey[at];
    if (k) {
        const unsigned key = k - 1, px = key % ps, t = key / ps, py = ps - 1 - t % ps, plate = t / ps;
        idx[at] = static_cast<int32_t>(plate * ps * ps + py * ps + px);
    } else {
        idx[at] = -1;
    }
    const unsigned tk = tintkey[at];
    tint[at] = tk ? static_cast<uint8_t>((tk - 1) / ps / ps) : 255;
}

}  // namespace blinky
