// Device side of the B200 lens-warp path — kernels, resident state, frame pipeline.
//
// The reference's hot loop (/root/reference/engine/NQ/fisheye.c:2406-2424) is,
// per screen pixel: load an 8-byte pointer, load the source byte through it,
// optionally map it through a 256-entry tint LUT chosen by a second per-pixel
// byte, store one byte.  Here one packed 32-bit lensmap entry per pixel
// replaces pointer + tint byte (4 B instead of 9 B read per pixel), entries
// are fetched as 128-bit vectors, source bytes are gathered through the
// read-only path, and four output pixels are written per 32-bit store so that
// every warp-level store instruction covers one full 128-byte line.
//
// sm_100a only.  HBM-bound byte work: no tensor cores on purpose.
#include "warp_device.h"

#include <cuda.h>  // CUtensorMap types only; the driver entry point is fetched at run time
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "../../include/blinky_b200.h"
#include "tile_plan.h"

namespace blinky {

namespace {

constexpr int kThreads = 256;
constexpr int kChunks = 4;                            // quads (4-pixel groups) per thread
constexpr int kQuadsPerBlock = kThreads * kChunks;    // 1024 quads = 4096 pixels per CTA
constexpr int kPixelsPerBlock = kQuadsPerBlock * 4;

struct WarpParams {
    const uint4 *lensmap4;      // packed entries, 4 per element; padded to whole CTAs
    const uint8_t *faces;       // frame 0
    size_t face_stride;         // bytes between frames
    const uint32_t *bg32;       // background, 4 pixels per element (padded like the lensmap)
    const uint8_t *lut;         // [6][256] rubix tint LUTs
    const uint32_t *rgba;       // [256] palette expansion table (RGBA mode)
    void *out;                  // frame 0
    size_t out_stride;          // bytes between frames
    uint32_t nquads;            // ceil(W*H / 4)
    uint32_t npix;              // W*H
};

__device__ __forceinline__ uint4 ld_lensmap(const uint4 *p) {
    // streamed once per frame by this SM: keep it out of L1 so the gathers own L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t ld_face(const uint8_t *p) {
    // read-only path, allocate in L1: neighbouring pixels hit the same sectors
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void st_stream_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void st_stream_v4(uint4 *p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// --------------------------------------------------------------------------
// K1: direct gather.  grid = (ceil(nquads/1024), nframes), block = 256.
// Thread t of CTA b owns quads b*1024 + c*256 + t, c = 0..3, so each of the
// four 128-bit lensmap loads of a warp covers 512 contiguous bytes and each
// 32-bit output store of a warp covers 128 contiguous bytes.
// --------------------------------------------------------------------------
template <bool RUBIX, bool RGBA>
__global__ void __launch_bounds__(kThreads) warp_gather_kernel(const WarpParams p) {
    __shared__ uint8_t s_lut[RUBIX ? 6 * 256 : 4];
    __shared__ uint32_t s_rgba[RGBA ? 256 : 1];
    if (RUBIX) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.lut);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
        for (int i = threadIdx.x; i < 6 * 256 / 4; i += kThreads) dst[i] = __ldg(src + i);
    }
    if (RGBA) {
        for (int i = threadIdx.x; i < 256; i += kThreads) s_rgba[i] = __ldg(p.rgba + i);
    }
    if (RUBIX || RGBA) __syncthreads();

    const uint8_t *__restrict__ faces = p.faces + static_cast<size_t>(blockIdx.y) * p.face_stride;
    const uint32_t q0 = blockIdx.x * kQuadsPerBlock + threadIdx.x;

    uint4 e[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) e[c] = ld_lensmap(p.lensmap4 + q0 + c * kThreads);  // padded: always in range

    uint32_t v[kChunks][4];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const uint32_t ent[4] = {e[c].x, e[c].y, e[c].z, e[c].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v[c][k] = 0;
            if (ent[k] & BLINKY_LM_VALID) v[c][k] = ld_face(faces + (ent[k] & BLINKY_LM_INDEX_MASK));
        }
    }

#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const uint32_t q = q0 + c * kThreads;
        if (q >= p.nquads) continue;
        const uint32_t ent[4] = {e[c].x, e[c].y, e[c].z, e[c].w};
        const uint32_t all_valid = ent[0] & ent[1] & ent[2] & ent[3] & BLINKY_LM_VALID;
        uint32_t bgw = 0;
        if (!all_valid) bgw = __ldg(p.bg32 + q);
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t b = v[c][k];
            if (RUBIX) {
                const uint32_t t = (ent[k] >> BLINKY_LM_TINT_SHIFT) & 7u;
                if (t != BLINKY_LM_TINT_NONE) b = s_lut[t * 256 + b];
            }
            if (!(ent[k] & BLINKY_LM_VALID)) b = (bgw >> (8 * k)) & 0xffu;
            px[k] = b;
        }
        if (RGBA) {
            uint4 w = make_uint4(s_rgba[px[0]], s_rgba[px[1]], s_rgba[px[2]], s_rgba[px[3]]);
            uint4 *o = reinterpret_cast<uint4 *>(static_cast<uint8_t *>(p.out) + static_cast<size_t>(blockIdx.y) * p.out_stride);
            st_stream_v4(o + q, w);
        } else {
            uint32_t w = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
            uint32_t *o = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(p.out) + static_cast<size_t>(blockIdx.y) * p.out_stride);
            st_stream_u32(o + q, w);
        }
    }
}

// --------------------------------------------------------------------------
// K0: scalar kernel, one pixel per thread, byte stores.  Used only when the
// frame size or the caller's pointers rule out 32-bit stores (W*H % 4 != 0 or
// unaligned strides) — a correctness path for ragged sizes, not a fast path.
// --------------------------------------------------------------------------
template <bool RUBIX, bool RGBA>
__global__ void __launch_bounds__(kThreads) warp_scalar_kernel(const WarpParams p) {
    const uint32_t i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= p.npix) return;
    const uint32_t ent = __ldg(reinterpret_cast<const uint32_t *>(p.lensmap4) + i);
    const uint8_t *faces = p.faces + static_cast<size_t>(blockIdx.y) * p.face_stride;
    uint32_t b;
    if (ent & BLINKY_LM_VALID) {
        b = ld_face(faces + (ent & BLINKY_LM_INDEX_MASK));
        if (RUBIX) {
            const uint32_t t = (ent >> BLINKY_LM_TINT_SHIFT) & 7u;
            if (t != BLINKY_LM_TINT_NONE) b = __ldg(p.lut + t * 256 + b);
        }
    } else {
        b = __ldg(reinterpret_cast<const uint8_t *>(p.bg32) + i);
    }
    uint8_t *o = static_cast<uint8_t *>(p.out) + static_cast<size_t>(blockIdx.y) * p.out_stride;
    if (RGBA) reinterpret_cast<uint32_t *>(o)[i] = __ldg(p.rgba + b);
    else o[i] = static_cast<uint8_t>(b);
}


// --------------------------------------------------------------------------
// K2: tiled, warp-specialised kernel with TMA-staged source boxes.
//
// Persistent CTAs (1 producer warp + 8 consumer warps) walk work items
// (tile, frame) in tile-major order through a kStages-deep shared-memory ring.
//   producer (one lane): waits for a free stage, copies the 16-byte TileDesc into
//     the stage, and issues (a) one 4-D TMA tensor load (x, y, plate, frame) of the
//     tile's source box and (b) one bulk copy of the tile's entry block — both
//     complete on the stage's `full` mbarrier.  It runs up to kStages items ahead,
//     so every global-memory latency is off the consumers' critical path.
//   consumers (256 threads, one 4-pixel quad each): wait on `full`, read the
//     descriptor, their four entries and the four source bytes from SHARED memory,
//     apply the tint LUT, write one 32-bit word (one 128-bit word in RGBA mode),
//     then release the stage on its `empty` mbarrier.  No block-wide barrier.
// GATHER tiles (plate seams, singular points) carry 32-bit entries and read the
// globe directly, with the same 2-D thread layout (a warp covers 32x4 pixels).
// --------------------------------------------------------------------------
constexpr int kStages = 4;
constexpr int kConsumerWarps = 4;
constexpr int kConsumerThreads = kConsumerWarps * 32;   // 128: each owns two 4-pixel quads of a 32x32 tile
constexpr int kTiledThreads = kConsumerThreads + 32;    // + producer warp
constexpr int kTmapRows = 32;  // descriptor table index = (w16-1)*kTmapRows + (h8-1)

struct TiledParams {
    const TileDesc *tiles;
    const uint8_t *entries;
    const CUtensorMap *tmaps;
    const uint8_t *faces;
    size_t face_stride;
    const uint8_t *bg;
    const uint8_t *lut;
    const uint32_t *rgba;
    void *out;
    size_t out_stride;
    uint32_t ntiles, nframes, total;
    int width, height;
    uint32_t zero;  // always 0, but only the host knows: see stage_release()
};

struct __align__(128) TiledStage {
    uint8_t box[kMaxBoxBytes];          // TMA destination; GATHER tiles put their 4 KB of 32-bit entries here instead
    uint8_t entries[kTilePixels * 2];   // 16-bit entries of BOX tiles
    uint4 desc;                         // TileDesc
    uint32_t frame;
    uint32_t pad[3];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Releases a pipeline stage AFTER the values read from it have really arrived in
// registers.  A shared-memory load is only complete when its destination register is
// written; `mbarrier.arrive` does not wait for in-flight LDS (seen in SASS: the arrive
// issued right behind eight pending LDS, and the producer's next TMA then raced them).
// Passing a value computed from every loaded register as an (unused) asm input makes
// the scoreboard hold the arrive until those loads have landed.
// Hands a ring stage back to the producer once EVERY lane of the warp has the stage's data in
// registers.  `mbarrier.arrive` does not wait for shared-memory loads that are still in flight, and
// when the LSU queue is backed up (slow stores to a peer GPU or to host memory) an LDS can sit there
// long enough for the producer's next TMA to overwrite the stage under it.  An unused asm operand is
// not a dependency either — ptxas drops the computation feeding it.  So the loaded values are folded,
// through a kernel parameter that is always zero but unknown to the compiler, into the barrier's
// ADDRESS: the warp-wide OR reduction needs every lane's loaded registers, the arrive needs its result.
// CONVERGED: every lane executed the same load instructions (no divergence since the stage became
// visible).  A warp-level load completes as a whole, so lane 0's registers stand for all lanes and the
// reduction can be skipped.
template <bool CONVERGED>
__device__ __forceinline__ void stage_release(uint64_t *bar, uint32_t loaded_values, uint32_t zero, uint32_t lane) {
    uint32_t dep = loaded_values & zero;
    if (CONVERGED) __syncwarp();
    else dep = __reduce_or_sync(0xffffffffu, dep);
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar) + dep) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// NB (measured on B200, scripts/tma_probe.cu): the innermost coordinate must be a
// multiple of 16 bytes or the TMA unit raises "illegal instruction".
__device__ __forceinline__ void tma_load_box(void *smem_dst, const CUtensorMap *tmap, int x, int y, int plate, int frame,
                                             uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(x), "r"(y), "r"(plate), "r"(frame), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, d, 0x0040), 0x5410);
}

// one quad (4 consecutive pixels) of a BOX tile: e2 holds its four 16-bit entries
template <bool RUBIX, bool FULL>
__device__ __forceinline__ void box_quad(const uint8_t *__restrict__ box, const uint8_t *__restrict__ s_lut, uint2 e2,
                                         uint32_t (&px)[4], uint32_t &valid) {
    const uint32_t ent[4] = {e2.x & 0xffffu, e2.x >> 16, e2.y & 0xffffu, e2.y >> 16};
    valid = 0xfu;
    if (!FULL) valid = ((ent[0] >> 15) & 1u) | ((ent[1] >> 14) & 2u) | ((ent[2] >> 13) & 4u) | ((ent[3] >> 12) & 8u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // unmapped entries carry offset 0: the read is harmless and its value is replaced below
        uint32_t b = box[ent[k] & kBoxOffsetMask];
        if (RUBIX) {
            const uint32_t t = (ent[k] >> kBoxTintShift) & 7u;
            if (t != BLINKY_LM_TINT_NONE) b = s_lut[t * 256 + b];
        }
        px[k] = b;
    }
}

// GATHER tile (plate seams, singular points, strong minification): 32-bit entries,
// direct global reads.  Layout differs from the BOX path on purpose: a warp takes
// 8 tile rows and lane l is column l, so one warp-level load covers 32 consecutive
// screen pixels of ONE row — neighbouring texels of at most a few plate rows —
// instead of an 8x4 patch that touches 4x as many 32-byte sectors.
template <bool RUBIX, bool RGBA>
__device__ __forceinline__ void gather_tile(const TiledParams &p, const uint32_t *__restrict__ ent32, const uint8_t *__restrict__ faces,
                                            const uint8_t *__restrict__ s_lut, const uint32_t *__restrict__ s_rgba, uint8_t *out_frame,
                                            uint32_t tile_x, uint32_t tile_y, uint32_t warp, uint32_t lane, uint64_t *empty_bar,
                                            uint32_t stage_words) {
    uint32_t e[8], v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = ent32[(warp * 8 + j) * kTileW + lane];
    // entries are in registers: the stage can be refilled
    stage_release<true>(empty_bar, e[0] | e[1] | e[2] | e[3] | e[4] | e[5] | e[6] | e[7] | stage_words, p.zero, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = 0;
        if (e[j] & BLINKY_LM_VALID) v[j] = ld_face(faces + (e[j] & BLINKY_LM_INDEX_MASK));
    }
    const uint32_t x = tile_x + lane;
    if (x >= static_cast<uint32_t>(p.width)) return;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t y = tile_y + warp * 8 + j;
        if (y < static_cast<uint32_t>(p.height)) {
            const uint32_t pix = y * static_cast<uint32_t>(p.width) + x;
            uint32_t b = v[j];
            if (e[j] & BLINKY_LM_VALID) {
                if (RUBIX) {
                    const uint32_t t = (e[j] >> BLINKY_LM_TINT_SHIFT) & 7u;
                    if (t != BLINKY_LM_TINT_NONE) b = s_lut[t * 256 + b];
                }
            } else {
                b = __ldg(p.bg + pix);
            }
            if (RGBA) reinterpret_cast<uint32_t *>(out_frame)[pix] = s_rgba[b];
            else out_frame[pix] = static_cast<uint8_t>(b);
        }
    }
}

template <bool RGBA>
__device__ __forceinline__ void store_quad(void *out_frame, const uint32_t *__restrict__ s_rgba, uint32_t quad_index,
                                           const uint32_t (&px)[4]) {
    if (RGBA) {
        st_stream_v4(static_cast<uint4 *>(out_frame) + quad_index, make_uint4(s_rgba[px[0]], s_rgba[px[1]], s_rgba[px[2]], s_rgba[px[3]]));
    } else {
        st_stream_u32(static_cast<uint32_t *>(out_frame) + quad_index, pack4(px[0], px[1], px[2], px[3]));
    }
}

__device__ __forceinline__ void patch_background(const uint8_t *__restrict__ bg, uint32_t pix, uint32_t valid, uint32_t (&px)[4]) {
    if (valid != 0xfu) {
        const uint32_t bgw = __ldg(reinterpret_cast<const uint32_t *>(bg + pix));
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (!((valid >> k) & 1u)) px[k] = (bgw >> (8 * k)) & 0xffu;
    }
}

template <bool RUBIX, bool RGBA>
__global__ void __launch_bounds__(kTiledThreads, 8) warp_tiled_kernel(const TiledParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    TiledStage *stages = reinterpret_cast<TiledStage *>(smem_raw);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + sizeof(TiledStage) * kStages);
    uint64_t *empty_bar = full_bar + kStages;
    uint8_t *s_lut = reinterpret_cast<uint8_t *>(empty_bar + kStages);
    uint32_t *s_rgba = reinterpret_cast<uint32_t *>(s_lut + 6 * 256);

    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (RUBIX) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(p.lut);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_lut);
        for (int i = tid; i < 6 * 256 / 4; i += kTiledThreads) dst[i] = __ldg(src + i);
    }
    if (RGBA) {
        for (int i = tid; i < 256; i += kTiledThreads) s_rgba[i] = __ldg(p.rgba + i);
    }
    __syncthreads();

    const uint32_t G = gridDim.x;
    if (tid >= kConsumerThreads) {
        // ------------------------------ producer warp ------------------------------
        if (tid == kConsumerThreads) {
            uint32_t w = blockIdx.x;
            uint4 next = make_uint4(0, 0, 0, 0);
            if (w < p.total) next = __ldg(reinterpret_cast<const uint4 *>(p.tiles + w / p.nframes));
            for (uint32_t it = 0; w < p.total; w += G, ++it) {
                const uint32_t stage = it % kStages, round = it / kStages;
                const uint4 d = next;
                const uint32_t tile = w / p.nframes, frame = w - tile * p.nframes;
                const uint32_t wn = w + G;
                if (wn < p.total) next = __ldg(reinterpret_cast<const uint4 *>(p.tiles + wn / p.nframes));
                if (round > 0) mbar_wait(&empty_bar[stage], (round - 1) & 1u);
                TiledStage &st = stages[stage];
                st.desc = d;
                st.frame = frame;
                const uint32_t type = (d.z >> 8) & 0xffu;
                if (type == TILE_BOX || type == TILE_BOX_FULL) {
                    const uint32_t w16 = (d.z >> 16) & 0xffu, h8 = d.z >> 24;
                    const uint32_t box_bytes = w16 * 16u * h8 * 8u;
                    mbar_expect_tx(&full_bar[stage], box_bytes + kTilePixels * 2);
                    tma_load_box(st.box, p.tmaps + (w16 - 1) * kTmapRows + (h8 - 1), static_cast<int>(static_cast<int16_t>(d.y & 0xffffu)),
                                 static_cast<int>(static_cast<int16_t>(d.y >> 16)), static_cast<int>(d.z & 0xffu),
                                 static_cast<int>(frame), &full_bar[stage]);
                    bulk_copy_g2s(st.entries, p.entries + d.x, kTilePixels * 2, &full_bar[stage]);
                } else if (type == TILE_GATHER) {
                    mbar_expect_tx(&full_bar[stage], kTilePixels * 4);
                    bulk_copy_g2s(st.box, p.entries + d.x, kTilePixels * 4, &full_bar[stage]);
                } else {
                    mbar_arrive(&full_bar[stage]);
                }
            }
        }
        return;
    }

    // -------------------------------- consumers --------------------------------
    // thread t owns quad (row t/8, column t%8) and the quad 16 rows below it
    const uint32_t qx4 = (tid & 7u) * 4u, row = tid >> 3, lane = tid & 31u;
    const uint32_t width = static_cast<uint32_t>(p.width), height = static_cast<uint32_t>(p.height);
    uint32_t it = 0;
    for (uint32_t w = blockIdx.x; w < p.total; w += G, ++it) {
        const uint32_t stage = it % kStages, round = it / kStages;
        TiledStage &st = stages[stage];
        mbar_wait(&full_bar[stage], round & 1u);
        const uint4 d = st.desc;
        const uint32_t frame = st.frame;
        const uint32_t type = (d.z >> 8) & 0xffu;
        const uint32_t x = (d.w & 0xffffu) + qx4, y0 = (d.w >> 16) + row;
        const uint32_t pix0 = y0 * width + x, pix1 = pix0 + 16u * width;
        uint8_t *out_frame = static_cast<uint8_t *>(p.out) + static_cast<size_t>(frame) * p.out_stride;

        uint32_t pa[4], pb[4], va = 0, vb = 0;
        if (type == TILE_BOX_FULL) {
            const uint2 ea = reinterpret_cast<const uint2 *>(st.entries)[tid];
            const uint2 eb = reinterpret_cast<const uint2 *>(st.entries)[tid + kConsumerThreads];
            box_quad<RUBIX, true>(st.box, s_lut, ea, pa, va);
            box_quad<RUBIX, true>(st.box, s_lut, eb, pb, vb);
            stage_release<true>(&empty_bar[stage], pa[0] | pa[1] | pa[2] | pa[3] | pb[0] | pb[1] | pb[2] | pb[3] | d.w | frame, p.zero, lane);
            store_quad<RGBA>(out_frame, s_rgba, pix0 >> 2, pa);
            store_quad<RGBA>(out_frame, s_rgba, pix1 >> 2, pb);
            continue;
        }
        if (type == TILE_BOX) {
            const uint2 ea = reinterpret_cast<const uint2 *>(st.entries)[tid];
            const uint2 eb = reinterpret_cast<const uint2 *>(st.entries)[tid + kConsumerThreads];
            box_quad<RUBIX, false>(st.box, s_lut, ea, pa, va);
            box_quad<RUBIX, false>(st.box, s_lut, eb, pb, vb);
        } else if (type == TILE_GATHER) {
            gather_tile<RUBIX, RGBA>(p, reinterpret_cast<const uint32_t *>(st.box), p.faces + static_cast<size_t>(frame) * p.face_stride,
                                     s_lut, s_rgba, out_frame, d.w & 0xffffu, d.w >> 16, tid >> 5, lane, &empty_bar[stage],
                                     d.w | frame);  // the descriptor words read from the stage count too
            continue;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) pa[k] = pb[k] = 0;
        }
        // everything this warp needs from the stage is in registers: hand the stage back
        stage_release<false>(&empty_bar[stage], pa[0] | pa[1] | pa[2] | pa[3] | pb[0] | pb[1] | pb[2] | pb[3] | d.w | frame, p.zero, lane);
        if (x < width) {
            if (y0 < height) {
                patch_background(p.bg, pix0, va, pa);
                store_quad<RGBA>(out_frame, s_rgba, pix0 >> 2, pa);
            }
            if (y0 + 16u < height) {
                patch_background(p.bg, pix1, vb, pb);
                store_quad<RGBA>(out_frame, s_rgba, pix1 >> 2, pb);
            }
        }
    }
}


// --------------------------------------------------------------------------
// K3: companion of K2 for the tiles TMA staging cannot serve (GATHER: plate seams,
// singular points, strong minification; EMPTY: background only).  These are bound
// by global-memory latency, so instead of the ring they get plain parallelism:
// grid = (tiles, frame groups), 256 threads, a warp owns 4 tile rows and lane l is
// column l (one warp-level load = 32 consecutive screen pixels of one row).  The
// tile's entries are fetched once and reused for every frame of the group.
// --------------------------------------------------------------------------
constexpr int kGatherFramesPerCta = 4;

template <bool RUBIX, bool RGBA>
__global__ void __launch_bounds__(kThreads) warp_tile_gather_kernel(const TiledParams p, const uint32_t first_tile) {
    const uint32_t tid = threadIdx.x;
    const uint4 d = __ldg(reinterpret_cast<const uint4 *>(p.tiles + first_tile + blockIdx.x));
    const uint32_t type = (d.z >> 8) & 0xffu;
    const uint32_t tile_x = d.w & 0xffffu, tile_y = d.w >> 16;
    const uint32_t width = static_cast<uint32_t>(p.width), height = static_cast<uint32_t>(p.height);
    const uint32_t f0 = blockIdx.y * kGatherFramesPerCta;
    const uint32_t f1 = min(f0 + kGatherFramesPerCta, p.nframes);

    if (type == TILE_EMPTY) {
        // quad layout: thread t copies 4 background pixels of row t/8
        const uint32_t x = tile_x + (tid & 7u) * 4u, y = tile_y + (tid >> 3);
        if (x >= width || y >= height) return;
        const uint32_t pix = y * width + x;
        const uint32_t bgw = __ldg(reinterpret_cast<const uint32_t *>(p.bg + pix));
        const uint32_t px[4] = {bgw & 0xffu, (bgw >> 8) & 0xffu, (bgw >> 16) & 0xffu, bgw >> 24};
        for (uint32_t f = f0; f < f1; ++f) {
            uint8_t *o = static_cast<uint8_t *>(p.out) + static_cast<size_t>(f) * p.out_stride;
            if (RGBA) st_stream_v4(reinterpret_cast<uint4 *>(o) + (pix >> 2), make_uint4(__ldg(p.rgba + px[0]), __ldg(p.rgba + px[1]), __ldg(p.rgba + px[2]), __ldg(p.rgba + px[3])));
            else st_stream_u32(reinterpret_cast<uint32_t *>(o) + (pix >> 2), bgw);
        }
        return;
    }

    const uint32_t warp = tid >> 5, lane = tid & 31u;
    const uint32_t x = tile_x + lane;
    const uint32_t *__restrict__ ent32 = reinterpret_cast<const uint32_t *>(p.entries + d.x);
    uint32_t e[4], bgv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = __ldg(ent32 + (warp * 4 + j) * kTileW + lane);
    if (x >= width) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t y = tile_y + warp * 4 + j;
        bgv[j] = 0;
        if (!(e[j] & BLINKY_LM_VALID) && y < height) bgv[j] = __ldg(p.bg + y * width + x);
    }
    // all gathers of the CTA's frames are issued before the first store: 16 independent
    // loads in flight per thread instead of 4 (stores may alias loads as far as the compiler
    // knows, so a plain frame loop would serialise on them)
    uint32_t v[kGatherFramesPerCta][4];
#pragma unroll
    for (int g = 0; g < kGatherFramesPerCta; ++g) {
        const uint32_t f = f0 + g;
        const uint8_t *__restrict__ faces = p.faces + static_cast<size_t>(f < f1 ? f : f0) * p.face_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[g][j] = bgv[j];
            if (e[j] & BLINKY_LM_VALID) v[g][j] = ld_face(faces + (e[j] & BLINKY_LM_INDEX_MASK));
        }
    }
#pragma unroll
    for (int g = 0; g < kGatherFramesPerCta; ++g) {
        const uint32_t f = f0 + g;
        if (f >= f1) break;
        uint8_t *o = static_cast<uint8_t *>(p.out) + static_cast<size_t>(f) * p.out_stride;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t y = tile_y + warp * 4 + j;
            if (y < height) {
                uint32_t b = v[g][j];
                if (RUBIX && (e[j] & BLINKY_LM_VALID)) {
                    const uint32_t t = (e[j] >> BLINKY_LM_TINT_SHIFT) & 7u;
                    if (t != BLINKY_LM_TINT_NONE) b = __ldg(p.lut + t * 256 + b);
                }
                const uint32_t pix = y * width + x;
                if (RGBA) reinterpret_cast<uint32_t *>(o)[pix] = __ldg(p.rgba + b);
                else o[pix] = static_cast<uint8_t>(b);
            }
        }
    }
}

constexpr size_t kTiledSmemBytes = sizeof(TiledStage) * kStages + 2 * kStages * sizeof(uint64_t) + 6 * 256 + 256 * 4;

inline size_t round_up(size_t v, size_t m) { return (v + m - 1) / m * m; }

}  // namespace

// ---------------------------------------------------------------------------
// frame pipeline slot
// ---------------------------------------------------------------------------
// ---- host -> device plate upload without the copy engine --------------------------------------
// The rectangles a lens samples are strided (e.g. half a plate wide); 2-D DMA copies of them reach
// ~35 GB/s here while the link does 45-48.  Pinned host memory is mapped into the device address
// space (UVA), so the SMs can pull exactly those bytes themselves, 16 bytes per thread.
struct UploadRects {
    int n;
    uint32_t off[6];     // byte offset of the rectangle's first 16-byte column in the frame
    uint32_t vec_w[6];   // rectangle width in 16-byte vectors
    uint32_t rows[6];
    uint32_t first[7];   // prefix sum of vec_w * rows
    uint32_t pitch;      // platesize
};

__global__ void __launch_bounds__(256) upload_rects_kernel(const uint8_t *__restrict__ host_src, uint8_t *__restrict__ dst, const __grid_constant__ UploadRects R) {
    const uint32_t total = R.first[R.n];
    for (uint32_t v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
        int k = 0;
        while (v >= R.first[k + 1]) ++k;
        const uint32_t local = v - R.first[k];
        const uint32_t row = local / R.vec_w[k], col = local - row * R.vec_w[k];
        const size_t at = static_cast<size_t>(R.off[k]) + static_cast<size_t>(row) * R.pitch + static_cast<size_t>(col) * 16;
        const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(host_src + at));
        *reinterpret_cast<uint4 *>(dst + at) = x;
    }
}

struct WarpDevice::Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    uint8_t *d_faces = nullptr, *d_out = nullptr;
    uint8_t *h_faces = nullptr, *h_out = nullptr;  // pinned staging
    bool busy = false;
    // finalize info
    uint8_t *dst = nullptr;
    int dst_rowbytes = 0, x0 = 0, y0 = 0;
    bool keep_unmapped = false, direct = false;
};

struct WarpDevice::TmapSet {
    const void *faces = nullptr;
    size_t face_stride = 0;
    int nframes = 0;
    CUtensorMap *d_table = nullptr;  // [8 * kTmapRows]
    uint64_t last_use = 0;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

#define CK(call)                                      \
    do {                                              \
        cudaError_t e_ = (call);                      \
        if (e_ != cudaSuccess) return fail(#call, e_); \
    } while (0)

bool WarpDevice::fail(const char *what, int cuda_err) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(static_cast<cudaError_t>(cuda_err)));
    err_ = buf;
    return false;
}

WarpDevice::WarpDevice(int device) : device_(device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
    if (prop.major != 10) {
        char buf[256];
        snprintf(buf, sizeof buf, "blinky_b200 needs an sm_100 (B200) device; device %d is sm_%d%d (%s)", device, prop.major,
                 prop.minor, prop.name);
        throw std::runtime_error(buf);
    }
    sm_count_ = prop.multiProcessorCount;
    if (const char *e = getenv("BLINKY_E2E_UPLOAD")) upload_by_kernel_ = strcmp(e, "kernel") == 0;
    if (const char *e = getenv("BLINKY_E2E_OUT")) out_by_kernel_ = strcmp(e, "direct") == 0;
    cudaStream_t s;
    e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    stream_ = s;
    e = cudaMalloc(&d_lut_, 6 * 256);
    if (e == cudaSuccess) e = cudaMemset(d_lut_, 0, 6 * 256);
    if (e == cudaSuccess) e = cudaMalloc(&d_rgba_, 256 * 4);
    if (e == cudaSuccess) e = cudaMemset(d_rgba_, 0, 256 * 4);
    if (e != cudaSuccess) throw std::runtime_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
}

WarpDevice::~WarpDevice() {
    cudaSetDevice(device_);
    cudaDeviceSynchronize();
    for (Slot *s : slots_) {
        if (s->stream) cudaStreamDestroy(s->stream);
        if (s->done) cudaEventDestroy(s->done);
        cudaFree(s->d_faces);
        cudaFree(s->d_out);
        if (s->h_faces) cudaFreeHost(s->h_faces);
        if (s->h_out) cudaFreeHost(s->h_out);
        delete s;
    }
    cudaFree(d_lensmap_);
    cudaFree(d_lut_);
    cudaFree(d_bg_);
    cudaFree(d_rgba_);
    cudaFree(d_tiles_);
    cudaFree(d_entries_);
    for (TmapSet *t : tmap_sets_) {
        cudaFree(t->d_table);
        delete t;
    }
    if (stream_) cudaStreamDestroy(static_cast<cudaStream_t>(stream_));
}

bool WarpDevice::upload_lensmap(const LensmapUpload &lm) {
    CK(cudaSetDevice(device_));
    CK(cudaDeviceSynchronize());  // nothing may still be reading the old map
    const size_t npix = static_cast<size_t>(lm.width) * lm.height;
    const size_t npad = round_up(npix, kPixelsPerBlock);
    if (npad != npix_pad_) {
        cudaFree(d_lensmap_);
        cudaFree(d_bg_);
        d_lensmap_ = nullptr;
        d_bg_ = nullptr;
        CK(cudaMalloc(&d_lensmap_, npad * sizeof(uint32_t)));
        CK(cudaMalloc(&d_bg_, npad));
        CK(cudaMemset(d_bg_, 0, npad));
    } else if (lm.width != width_ || lm.height != height_) {
        CK(cudaMemset(d_bg_, 0, npad));
    }
    // padding entries are "unmapped"
    CK(cudaMemset(d_lensmap_, 0, npad * sizeof(uint32_t)));
    CK(cudaMemcpy(d_lensmap_, lm.packed, npix * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_lut_, lm.palmaps, 6 * 256, cudaMemcpyHostToDevice));
    width_ = lm.width;
    height_ = lm.height;
    platesize_ = lm.platesize;
    numplates_ = lm.numplates;
    npix_ = npix;
    npix_pad_ = npad;
    rubix_ = lm.rubix;
    memcpy(display_, lm.display, sizeof display_);
    memcpy(plate_rect_, lm.plate_rect, sizeof plate_rect_);
    span_off_.assign(lm.span_off, lm.span_off + lm.height + 1);
    spans_.assign(lm.spans, lm.spans + lm.nspans * 2);
    have_lensmap_ = true;
    // tiled layout
    have_plan_ = false;
    plan_has_box_ = false;
    cudaFree(d_tiles_);
    cudaFree(d_entries_);
    d_tiles_ = nullptr;
    d_entries_ = nullptr;
    for (TmapSet *t : tmap_sets_) {
        cudaFree(t->d_table);
        delete t;
    }
    tmap_sets_.clear();
    if (lm.plan && !lm.plan->tiles.empty()) {
        const TilePlan &pl = *lm.plan;
        CK(cudaMalloc(&d_tiles_, pl.tiles.size() * sizeof(TileDesc)));
        CK(cudaMemcpy(d_tiles_, pl.tiles.data(), pl.tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_entries_, pl.entries.size()));
        CK(cudaMemcpy(d_entries_, pl.entries.data(), pl.entries.size(), cudaMemcpyHostToDevice));
        ntiles_ = static_cast<uint32_t>(pl.tiles.size());
        nbox_tiles_ = static_cast<uint32_t>(pl.n_box);
        shapes_ = pl.shapes;
        plan_has_box_ = pl.n_box > 0;
        have_plan_ = true;
        char buf[256];
        snprintf(buf, sizeof buf, "tiles %dx%d of %dx%d px: %d box (TMA, %.2f B/px staged), %d gather, %d empty; entries %.2f B/px",
                 pl.tiles_x, pl.tiles_y, kTileW, kTileH, pl.n_box, static_cast<double>(pl.box_bytes) / static_cast<double>(npix),
                 pl.n_gather, pl.n_empty, static_cast<double>(pl.entries.size()) / static_cast<double>(npix));
        plan_summary_ = buf;
    }
    // frame slots depend on the sizes: rebuild lazily
    for (Slot *s : slots_) {
        cudaStreamDestroy(s->stream);
        cudaEventDestroy(s->done);
        cudaFree(s->d_faces);
        cudaFree(s->d_out);
        cudaFreeHost(s->h_faces);
        cudaFreeHost(s->h_out);
        delete s;
    }
    slots_.clear();
    return true;
}

bool WarpDevice::set_background(const uint8_t *bg_host) {
    if (!have_lensmap_) {
        err_ = "set_background: build a lensmap first (the background has the view's size)";
        return false;
    }
    CK(cudaSetDevice(device_));
    CK(cudaDeviceSynchronize());
    if (bg_host) CK(cudaMemcpy(d_bg_, bg_host, npix_, cudaMemcpyHostToDevice));
    else CK(cudaMemset(d_bg_, 0, npix_pad_));
    return true;
}

bool WarpDevice::set_rgba_table(const uint32_t table[256]) {
    CK(cudaSetDevice(device_));
    CK(cudaMemcpy(d_rgba_, table, 256 * 4, cudaMemcpyHostToDevice));
    have_rgba_ = true;
    return true;
}

bool WarpDevice::warp(const void *d_faces, size_t face_stride, void *d_out, size_t out_stride, int nframes, void *stream,
                      bool rgba) {
    if (!have_lensmap_) {
        err_ = "warp: no lensmap on the device (call blinky_build_lensmap)";
        return false;
    }
    if (nframes <= 0) return true;
    if (nframes > 65535) {
        err_ = "warp: at most 65535 frames per launch";
        return false;
    }
    const size_t opx = rgba ? 4 : 1;
    // the tiled kernel writes 4-pixel words at (y*W + x): needs W % 4 == 0 and aligned frames
    const bool tiled_ok = have_plan_ && (width_ % 4 == 0) && (reinterpret_cast<uintptr_t>(d_out) % (4 * opx) == 0) &&
                          (out_stride % (4 * opx) == 0 || nframes == 1) &&
                          (!plan_has_box_ || (reinterpret_cast<uintptr_t>(d_faces) % 16 == 0 && (face_stride % 16 == 0 || nframes == 1)));
    if (variant_ == BLINKY_KERNEL_GATHER || !tiled_ok) return launch_flat(d_faces, face_stride, d_out, out_stride, nframes, stream, rgba);
    return launch_tiled(d_faces, face_stride, d_out, out_stride, nframes, stream, rgba);
}

WarpDevice::TmapSet *WarpDevice::get_tmaps(const void *d_faces, size_t face_stride, int nframes, void *stream) {
    static uint64_t tick = 0;
    ++tick;
    for (TmapSet *t : tmap_sets_)
        if (t->faces == d_faces && t->face_stride == face_stride && t->nframes == nframes) {
            t->last_use = tick;
            return t;
        }
    if (!encode_fn_) {
        cudaDriverEntryPointQueryResult qres;
        void *fn = nullptr;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            fail("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)", e);
            return nullptr;
        }
        encode_fn_ = fn;
    }
    TmapSet *t = nullptr;
    if (tmap_sets_.size() >= 8) {  // recycle the least recently used table (its kernels are long gone)
        t = tmap_sets_[0];
        for (TmapSet *c : tmap_sets_)
            if (c->last_use < t->last_use) t = c;
        cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    } else {
        t = new TmapSet();
        if (cudaMalloc(&t->d_table, sizeof(CUtensorMap) * 8 * kTmapRows) != cudaSuccess) {
            delete t;
            fail("cudaMalloc(tensor maps)", cudaGetLastError());
            return nullptr;
        }
        tmap_sets_.push_back(t);
    }
    t->faces = d_faces;
    t->face_stride = face_stride;
    t->nframes = nframes;
    t->last_use = tick;
    std::vector<CUtensorMap> host(8 * kTmapRows);
    memset(host.data(), 0, host.size() * sizeof(CUtensorMap));
    const cuuint64_t ps = static_cast<cuuint64_t>(platesize_);
    // 4-D view of the globe: (x, y, plate, frame)
    const cuuint64_t dims[4] = {ps, ps, static_cast<cuuint64_t>(numplates_), static_cast<cuuint64_t>(nframes)};
    const cuuint64_t fstride = nframes > 1 ? static_cast<cuuint64_t>(face_stride) : ps * ps * static_cast<cuuint64_t>(numplates_);
    const cuuint64_t strides[3] = {ps, ps * ps, fstride};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (uint16_t shape : shapes_) {
        const uint32_t w16 = shape >> 8, h8 = shape & 0xff;
        const cuuint32_t box[4] = {w16 * 16, h8 * 8, 1, 1};
        CUresult r = reinterpret_cast<EncodeTiledFn>(encode_fn_)(
            &host[(w16 - 1) * kTmapRows + (h8 - 1)], CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<void *>(d_faces), dims, strides,
            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            char buf[160];
            snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(box %ux%u, ps %d) failed with CUresult %d", w16 * 16, h8 * 8, platesize_, static_cast<int>(r));
            err_ = buf;
            t->faces = nullptr;
            return nullptr;
        }
    }
    if (cudaMemcpyAsync(t->d_table, host.data(), host.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice,
                        static_cast<cudaStream_t>(stream)) != cudaSuccess) {
        fail("cudaMemcpyAsync(tensor maps)", cudaGetLastError());
        t->faces = nullptr;
        return nullptr;
    }
    return t;
}

bool WarpDevice::launch_tiled(const void *d_faces, size_t face_stride, void *d_out, size_t out_stride, int nframes,
                              void *stream, bool rgba) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TiledParams p;
    p.tiles = static_cast<const TileDesc *>(d_tiles_);
    p.entries = d_entries_;
    p.tmaps = nullptr;
    if (plan_has_box_) {
        TmapSet *t = get_tmaps(d_faces, face_stride, nframes, stream);
        if (!t) return false;
        p.tmaps = t->d_table;
    }
    p.faces = static_cast<const uint8_t *>(d_faces);
    p.face_stride = face_stride;
    p.bg = d_bg_;
    p.lut = d_lut_;
    p.rgba = d_rgba_;
    p.out = d_out;
    p.out_stride = out_stride;
    // Tiles [0, nbox) are BOX tiles.  When only a few tiles are GATHER/EMPTY they ride along
    // in the TMA ring (one launch; their global latency hides among the BOX tiles); when
    // they are many (strong minification, big unmapped borders) they go to the gather
    // kernel K3, which hides that latency with plain parallelism.
    const bool split = (ntiles_ - nbox_tiles_) * 100u > ntiles_ * static_cast<uint32_t>(split_percent_);
    const uint32_t ring_tiles = split ? nbox_tiles_ : ntiles_;
    p.ntiles = ring_tiles;
    p.nframes = static_cast<uint32_t>(nframes);
    p.total = ring_tiles * static_cast<uint32_t>(nframes);
    p.width = width_;
    p.height = height_;
    p.zero = 0;
    const bool rubix = rubix_;
    const int vi = (rubix ? 1 : 0) | (rgba ? 2 : 0);
    if (tiled_ctas_per_sm_[vi] == 0) {
        int n = 0;
        cudaError_t e;
        const void *fn = rubix && rgba ? reinterpret_cast<const void *>(warp_tiled_kernel<true, true>)
                         : rubix       ? reinterpret_cast<const void *>(warp_tiled_kernel<true, false>)
                         : rgba        ? reinterpret_cast<const void *>(warp_tiled_kernel<false, true>)
                                       : reinterpret_cast<const void *>(warp_tiled_kernel<false, false>);
        e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kTiledSmemBytes));
        if (e != cudaSuccess) return fail("cudaFuncSetAttribute(max dynamic smem)", e);
        if (rubix && rgba) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, warp_tiled_kernel<true, true>, kTiledThreads, kTiledSmemBytes);
        else if (rubix) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, warp_tiled_kernel<true, false>, kTiledThreads, kTiledSmemBytes);
        else if (rgba) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, warp_tiled_kernel<false, true>, kTiledThreads, kTiledSmemBytes);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, warp_tiled_kernel<false, false>, kTiledThreads, kTiledSmemBytes);
        if (e != cudaSuccess || n < 1) n = 4;
        tiled_ctas_per_sm_[vi] = n;
    }
    uint32_t grid = static_cast<uint32_t>(sm_count_ * tiled_ctas_per_sm_[vi]);
    if (grid > p.total) grid = p.total;
    char buf[320];
    int n = 0;
    if (p.total > 0) {
        if (rubix && rgba) warp_tiled_kernel<true, true><<<grid, kTiledThreads, kTiledSmemBytes, st>>>(p);
        else if (rubix) warp_tiled_kernel<true, false><<<grid, kTiledThreads, kTiledSmemBytes, st>>>(p);
        else if (rgba) warp_tiled_kernel<false, true><<<grid, kTiledThreads, kTiledSmemBytes, st>>>(p);
        else warp_tiled_kernel<false, false><<<grid, kTiledThreads, kTiledSmemBytes, st>>>(p);
        ++launches_;
        n = snprintf(buf, sizeof buf, "warp_tiled_kernel<rubix=%d,rgba=%d> grid=%u block=%d (%d CTAs/SM, persistent, %d-stage TMA ring)", rubix, rgba,
                     grid, kTiledThreads, tiled_ctas_per_sm_[vi], kStages);
    }
    const uint32_t nother = ntiles_ - ring_tiles;
    if (nother > 0) {
        dim3 g2(nother, static_cast<unsigned>((nframes + kGatherFramesPerCta - 1) / kGatherFramesPerCta));
        if (rubix && rgba) warp_tile_gather_kernel<true, true><<<g2, kThreads, 0, st>>>(p, ring_tiles);
        else if (rubix) warp_tile_gather_kernel<true, false><<<g2, kThreads, 0, st>>>(p, ring_tiles);
        else if (rgba) warp_tile_gather_kernel<false, true><<<g2, kThreads, 0, st>>>(p, ring_tiles);
        else warp_tile_gather_kernel<false, false><<<g2, kThreads, 0, st>>>(p, ring_tiles);
        ++launches_;
        snprintf(buf + n, sizeof buf - static_cast<size_t>(n), "%swarp_tile_gather_kernel<rubix=%d,rgba=%d> grid=(%u,%u) block=%d", n ? " + " : "", rubix, rgba,
                 g2.x, g2.y, kThreads);
    }
    last_kernel_ = buf;
    CK(cudaGetLastError());
    return true;
}

bool WarpDevice::launch_flat(const void *d_faces, size_t face_stride, void *d_out, size_t out_stride, int nframes, void *stream,
                             bool rgba) {
    // NULL is CUDA's default stream (what torch.cuda.current_stream() hands out
    // unless the caller made its own) — NOT this context's private stream.
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    WarpParams p;
    p.lensmap4 = reinterpret_cast<const uint4 *>(d_lensmap_);
    p.faces = static_cast<const uint8_t *>(d_faces);
    p.face_stride = face_stride;
    p.bg32 = reinterpret_cast<const uint32_t *>(d_bg_);
    p.lut = d_lut_;
    p.rgba = d_rgba_;
    p.out = d_out;
    p.out_stride = out_stride;
    p.nquads = static_cast<uint32_t>((npix_ + 3) / 4);
    p.npix = static_cast<uint32_t>(npix_);

    const size_t opx = rgba ? 4 : 1;  // output bytes per pixel
    const bool vector_ok = (npix_ % 4 == 0) && (reinterpret_cast<uintptr_t>(d_out) % (4 * opx) == 0) &&
                           (out_stride % (4 * opx) == 0 || nframes == 1);
    const bool rubix = rubix_;
    if (vector_ok) {
        dim3 grid(static_cast<unsigned>(npix_pad_ / kPixelsPerBlock), static_cast<unsigned>(nframes));
        if (rubix && rgba) warp_gather_kernel<true, true><<<grid, kThreads, 0, st>>>(p);
        else if (rubix) warp_gather_kernel<true, false><<<grid, kThreads, 0, st>>>(p);
        else if (rgba) warp_gather_kernel<false, true><<<grid, kThreads, 0, st>>>(p);
        else warp_gather_kernel<false, false><<<grid, kThreads, 0, st>>>(p);
        char buf[160];
        snprintf(buf, sizeof buf, "warp_gather_kernel<rubix=%d,rgba=%d> grid=(%u,%u) block=%d", rubix, rgba, grid.x, grid.y, kThreads);
        last_kernel_ = buf;
    } else {
        dim3 grid(static_cast<unsigned>((npix_ + kThreads - 1) / kThreads), static_cast<unsigned>(nframes));
        if (rubix && rgba) warp_scalar_kernel<true, true><<<grid, kThreads, 0, st>>>(p);
        else if (rubix) warp_scalar_kernel<true, false><<<grid, kThreads, 0, st>>>(p);
        else if (rgba) warp_scalar_kernel<false, true><<<grid, kThreads, 0, st>>>(p);
        else warp_scalar_kernel<false, false><<<grid, kThreads, 0, st>>>(p);
        char buf[160];
        snprintf(buf, sizeof buf, "warp_scalar_kernel<rubix=%d,rgba=%d> grid=(%u,%u) block=%d", rubix, rgba, grid.x, grid.y, kThreads);
        last_kernel_ = buf;
    }
    ++launches_;
    CK(cudaGetLastError());
    return true;
}

bool WarpDevice::ensure_slots() {
    if (!slots_.empty()) return true;
    int kSlots = 3;
    if (const char *e = getenv("BLINKY_HOST_SLOTS")) {  // pipeline depth of blinky_warp_host (experiments)
        const int v = atoi(e);
        if (v >= 1 && v <= 16) kSlots = v;
    }
    slot_face_bytes_ = static_cast<size_t>(numplates_) * platesize_ * platesize_;
    slot_out_bytes_ = round_up(npix_, 16);
    for (int i = 0; i < kSlots; ++i) {
        Slot *s = new Slot();
        slots_.push_back(s);
        CK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
        CK(cudaMalloc(&s->d_faces, slot_face_bytes_));
        CK(cudaMemset(s->d_faces, 0, slot_face_bytes_));
        CK(cudaMalloc(&s->d_out, slot_out_bytes_));
        CK(cudaMallocHost(&s->h_faces, slot_face_bytes_));
        CK(cudaMallocHost(&s->h_out, slot_out_bytes_));
    }
    return true;
}

void WarpDevice::finalize_slot(Slot &s) {
    if (!s.busy) return;
    cudaEventSynchronize(s.done);
    s.busy = false;
    if (s.direct) return;  // the copy engine already wrote the caller's buffer
    const int W = width_, H = height_;
    uint8_t *dst = s.dst + static_cast<size_t>(s.y0) * s.dst_rowbytes + s.x0;
    if (s.keep_unmapped) {
        // only mapped pixels are written, like `if (*lmap)` in render_lensmap (:2413)
        for (int y = 0; y < H; ++y) {
            const uint8_t *src = s.h_out + static_cast<size_t>(y) * W;
            uint8_t *row = dst + static_cast<size_t>(y) * s.dst_rowbytes;
            for (int32_t k = span_off_[static_cast<size_t>(y)]; k < span_off_[static_cast<size_t>(y) + 1]; ++k) {
                const int32_t a = spans_[static_cast<size_t>(k) * 2], b = spans_[static_cast<size_t>(k) * 2 + 1];
                memcpy(row + a, src + a, static_cast<size_t>(b - a));
            }
        }
    } else if (s.dst_rowbytes == W) {
        memcpy(dst, s.h_out, static_cast<size_t>(W) * H);
    } else {
        for (int y = 0; y < H; ++y) memcpy(dst + static_cast<size_t>(y) * s.dst_rowbytes, s.h_out + static_cast<size_t>(y) * W, static_cast<size_t>(W));
    }
}

static bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

bool WarpDevice::warp_host(const uint8_t *faces_host, size_t face_stride, uint8_t *dst_host, size_t dst_frame_stride,
                           int dst_rowbytes, int x0, int y0, int nframes, bool keep_unmapped) {
    if (!have_lensmap_) {
        err_ = "warp_host: no lensmap on the device (call blinky_build_lensmap)";
        return false;
    }
    CK(cudaSetDevice(device_));
    if (!ensure_slots()) return false;
    const size_t ps2 = static_cast<size_t>(platesize_) * platesize_;
    const bool src_pinned = is_pinned(faces_host);
    const bool dst_pinned = is_pinned(dst_host);
    const int W = width_, H = height_;
    bool ok = true;
    for (int f = 0; f < nframes && ok; ++f) {
        Slot &s = *slots_[static_cast<size_t>(f) % slots_.size()];
        finalize_slot(s);  // frees the slot (waits for frame f-3)
        const uint8_t *src = faces_host + static_cast<size_t>(f) * face_stride;
        // Only what the lens looks at is uploaded: plates with display != 0 (:764-766), and of
        // those only the texel rectangle the lensmap samples.  (TMA boxes may overhang the
        // rectangle; those texels are staged but never referenced by an entry.)
        UploadRects ur;
        ur.n = 0;
        ur.first[0] = 0;
        ur.pitch = static_cast<uint32_t>(platesize_);
        const bool by_kernel = upload_by_kernel_ && platesize_ % 16 == 0 && reinterpret_cast<uintptr_t>(src) % 16 == 0 && face_stride % 16 == 0;
        for (int pl = 0; pl < numplates_; ++pl) {
            if (!display_[pl]) continue;
            const int *r = plate_rect_[pl];
            if (r[0] > r[2] || r[1] > r[3]) continue;
            if (by_kernel && src_pinned) {
                const int xa = r[0] & ~15, xb = (r[2] + 16) & ~15;  // 16-byte columns covering [r0, r2]
                ur.off[ur.n] = static_cast<uint32_t>(pl * ps2 + static_cast<size_t>(r[1]) * platesize_ + xa);
                ur.vec_w[ur.n] = static_cast<uint32_t>((xb - xa) / 16);
                ur.rows[ur.n] = static_cast<uint32_t>(r[3] - r[1] + 1);
                ur.first[ur.n + 1] = ur.first[ur.n] + ur.vec_w[ur.n] * ur.rows[ur.n];
                ++ur.n;
                continue;
            }
            const size_t rw = static_cast<size_t>(r[2] - r[0] + 1), rh = static_cast<size_t>(r[3] - r[1] + 1);
            const size_t off = pl * ps2 + static_cast<size_t>(r[1]) * platesize_ + r[0];
            const uint8_t *from = src + off;
            if (!src_pinned) {
                for (size_t y = 0; y < rh; ++y) memcpy(s.h_faces + off + y * platesize_, from + y * platesize_, rw);
                from = s.h_faces + off;
            }
            cudaError_t e = cudaMemcpy2DAsync(s.d_faces + off, static_cast<size_t>(platesize_), from, static_cast<size_t>(platesize_), rw, rh,
                                              cudaMemcpyHostToDevice, s.stream);
            if (e != cudaSuccess) { ok = fail("cudaMemcpy2DAsync(H2D faces)", e); break; }
        }
        if (ok && ur.n > 0) {
            const uint32_t total = ur.first[ur.n];
            const unsigned blocks = std::min<unsigned>((total + 255) / 256, static_cast<unsigned>(sm_count_) * 8u);
            upload_rects_kernel<<<blocks, 256, 0, s.stream>>>(src, s.d_faces, ur);
            ++launches_;
        }
        if (!ok) break;
        s.dst = dst_host + static_cast<size_t>(f) * dst_frame_stride;
        // the warp kernel can store straight into the caller's pinned frame (posted PCIe writes, no
        // staging copy) when the frame is tightly packed
        const bool zero_copy_out = out_by_kernel_ && dst_pinned && !keep_unmapped && dst_rowbytes == W && x0 == 0 && y0 == 0 &&
                                   reinterpret_cast<uintptr_t>(s.dst) % 16 == 0;
        if (!warp(s.d_faces, slot_face_bytes_, zero_copy_out ? s.dst : s.d_out, slot_out_bytes_, 1, s.stream, false)) { ok = false; break; }
        s.dst_rowbytes = dst_rowbytes;
        s.x0 = x0;
        s.y0 = y0;
        s.keep_unmapped = keep_unmapped;
        s.direct = dst_pinned && !keep_unmapped;
        cudaError_t e;
        if (zero_copy_out) {
            e = cudaSuccess;
        } else if (s.direct) {
            e = cudaMemcpy2DAsync(s.dst + static_cast<size_t>(y0) * dst_rowbytes + x0, static_cast<size_t>(dst_rowbytes), s.d_out,
                                  static_cast<size_t>(W), static_cast<size_t>(W), static_cast<size_t>(H), cudaMemcpyDeviceToHost, s.stream);
        } else {
            e = cudaMemcpyAsync(s.h_out, s.d_out, static_cast<size_t>(W) * H, cudaMemcpyDeviceToHost, s.stream);
        }
        if (e != cudaSuccess) { ok = fail("cudaMemcpyAsync(D2H frame)", e); break; }
        e = cudaEventRecord(s.done, s.stream);
        if (e != cudaSuccess) { ok = fail("cudaEventRecord", e); break; }
        s.busy = true;
    }
    // drain in submission order
    for (size_t k = 0; k < slots_.size(); ++k) {
        Slot &s = *slots_[(static_cast<size_t>(nframes) + k) % slots_.size()];
        finalize_slot(s);
    }
    if (ok) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) ok = fail("warp_host", e);
    }
    return ok;
}

size_t WarpDevice::upload_bytes_per_frame() const {
    size_t n = 0;
    for (int pl = 0; pl < numplates_; ++pl) {
        const int *r = plate_rect_[pl];
        if (display_[pl] && r[0] <= r[2] && r[1] <= r[3]) n += static_cast<size_t>(r[2] - r[0] + 1) * static_cast<size_t>(r[3] - r[1] + 1);
    }
    return n;
}

bool WarpDevice::alloc_device(size_t bytes, void **out) {
    CK(cudaSetDevice(device_));
    CK(cudaMalloc(out, bytes));
    return true;
}

bool WarpDevice::free_device(void *p) {
    CK(cudaSetDevice(device_));
    CK(cudaFree(p));
    return true;
}

bool WarpDevice::ipc_export(void *p, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaSetDevice(device_));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle, &h, 64);
    return true;
}

bool WarpDevice::ipc_open(const unsigned char handle[64], void **out) {
    CK(cudaSetDevice(device_));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));  // maps the peer GPU's memory over NVLink
    return true;
}

bool WarpDevice::ipc_close(void *p) {
    CK(cudaSetDevice(device_));
    CK(cudaIpcCloseMemHandle(p));
    return true;
}

bool WarpDevice::alloc_pinned(size_t bytes, void **out) {
    CK(cudaSetDevice(device_));
    CK(cudaMallocHost(out, bytes));
    return true;
}

bool WarpDevice::free_pinned(void *p) {
    CK(cudaFreeHost(p));
    return true;
}

bool WarpDevice::sync() {
    CK(cudaSetDevice(device_));
    CK(cudaDeviceSynchronize());
    return true;
}

}  // namespace blinky
