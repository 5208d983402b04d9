// Diagnostic: which TMA box shapes / descriptor placements work for the 4-D uint8
// globe view (x, y, plate, frame).  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_probe scripts/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ void do_load(const CUtensorMap *tm, unsigned char *box, unsigned long long *bar, int bytes, int x, int y, int pl, int fr,
                        unsigned char *out) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(box)), "l"(tm), "r"(x), "r"(y), "r"(pl), "r"(fr), "r"(smem_u32(bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = box[i];
}

__global__ void k_param(const __grid_constant__ CUtensorMap tm, int bytes, int x, int y, int pl, int fr, unsigned char *out) {
    __shared__ __align__(128) unsigned char box[4096];
    __shared__ __align__(8) unsigned long long bar;
    do_load(&tm, box, &bar, bytes, x, y, pl, fr, out);
}
__global__ void k_global(const CUtensorMap *tm, int bytes, int x, int y, int pl, int fr, unsigned char *out) {
    __shared__ __align__(128) unsigned char box[4096];
    __shared__ __align__(8) unsigned long long bar;
    do_load(tm, box, &bar, bytes, x, y, pl, fr, out);
}


// usage: tma_probe rank bw bh x y mode(0=param,1=global) [l2promo] -> one load, verified; exit code 0 ok
int main(int argc, char **argv) {
    if (argc < 7) { printf("usage\n"); return 9; }
    const int rank = atoi(argv[1]), bw = atoi(argv[2]), bh = atoi(argv[3]), x = atoi(argv[4]), y = atoi(argv[5]), mode = atoi(argv[6]);
    const int promo = argc > 7 ? atoi(argv[7]) : 0;
    const int ps = 256, P = 6, F = 2;
    std::vector<unsigned char> h((size_t)ps * ps * P * F);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d, *dout;
    cudaMalloc(&d, h.size());
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&dout, 4096);
    CUtensorMap *dtm;
    cudaMalloc(&dtm, sizeof(CUtensorMap));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (!fn) { printf("no entry point\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t dims[4] = {ps, ps, P, F}, strides[3] = {ps, (cuuint64_t)ps * ps, (cuuint64_t)ps * ps * P};
    if (rank == 3) dims[2] = P * F;
    if (rank == 2) dims[1] = (cuuint64_t)ps * P * F;
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                     CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    const int pl = 4, fr = 1;
    cudaMemset(dout, 0xEE, 4096);
    if (rank != 4) { printf("only rank 4 kernels built; encode ok\n"); return 0; }
    if (mode == 0) k_param<<<1, 128>>>(tm, bw * bh, x, y, pl, fr, dout);
    else { cudaMemcpy(dtm, &tm, sizeof tm, cudaMemcpyHostToDevice); k_global<<<1, 128>>>(dtm, bw * bh, x, y, pl, fr, dout); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("FAIL: %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<unsigned char> got(bw * bh);
    cudaMemcpy(got.data(), dout, got.size(), cudaMemcpyDeviceToHost);
    int mism = 0;
    for (int r2 = 0; r2 < bh; ++r2)
        for (int c = 0; c < bw; ++c) {
            int yy = y + r2, xx = x + c;
            unsigned char want = (yy >= 0 && xx >= 0 && yy < ps && xx < ps) ? h[(((size_t)fr * P + pl) * ps + yy) * ps + xx] : 0;
            if (got[r2 * bw + c] != want) ++mism;
        }
    printf(mism ? "MISMATCH %d\n" : "ok\n", mism);
    return mism ? 4 : 0;
}
