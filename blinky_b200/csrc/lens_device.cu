// Device-side lensmap construction: NVRTC compile of the translated lens + launch.
// See lens_device.h.
#include "lens_device.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace blinky {

namespace {

// The per-pixel tail of the lensmap build, appended to the translated lens.  Operation
// order and types follow fisheye_host.cpp (which follows fisheye.c:2023-2066 ray_to_plate_index /
// ray_to_plate_uv, :1984-2013 set_from_ray, :1922-1960 rubix grid) line by line: after the ray is
// narrowed to float32 everything is IEEE +,-,*,/ and sqrt, compiled with --fmad=false, so it is
// bit-identical to the host.
const char *kKernelSource = R"KRN(
struct LtParams {
    int width, height, platesize, numplates;
    double scale;
    double rubix_block, rubix_pad, rubix_unit_px;
    double uv_dist[6];
    LtPlate plates[6];
};

static __device__ __forceinline__ float lt_dot3(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

extern "C" __global__ void __launch_bounds__(128) lt_build(const __grid_constant__ LtParams P, unsigned *__restrict__ cand) {
    const int lx = blockIdx.x * blockDim.x + threadIdx.x, ly = blockIdx.y;
    if (lx >= P.width) return;
    const double x = (lx - P.width / 2) * P.scale;
    const double y = -(ly - P.height / 2) * P.scale;
    Ctx c;
    c.flag = 0;
    c.steps = 0;
    c.plates = P.plates;
    c.numplates = P.numplates;
    lt_init_mut(c);
    LtD r[3];
    unsigned out = 0;
    if (lt_entry(c, x, y, r)) {
        float ray[3] = {lt_f32(c, r[0]), lt_f32(c, r[1]), lt_f32(c, r[2])};
        float len = ray[0] * ray[0] + ray[1] * ray[1] + ray[2] * ray[2];
        len = (float)sqrt((double)len);
        if (len) {
            const float inv = 1 / len;
            ray[0] *= inv; ray[1] *= inv; ray[2] *= inv;
        }
        int best = 0;
        double best_dp = -2;
        for (int i = 0; i < P.numplates; ++i) {
            const double dp = (double)lt_dot3(ray, P.plates[i].forward);
            if (dp > best_dp) { best_dp = dp; best = i; }
        }
        const LtPlate &p = P.plates[best];
        const double px_ = (double)lt_dot3(p.right, ray);
        const double py_ = (double)lt_dot3(p.up, ray);
        const double pz_ = (double)lt_dot3(p.forward, ray);
        const double dist = P.uv_dist[best];
        const double u = px_ / pz_ * dist + 0.5;
        const double v = -py_ / pz_ * dist + 0.5;
        if (u >= 0 && u <= 1 && v >= 0 && v <= 1) {
            const int ps = P.platesize;
            const int px = (int)(u * ps), py = (int)(v * ps);
            if (px >= 0 && px < ps && py >= 0 && py < ps) {
                const double ux = (double)px / P.rubix_unit_px, uy = (double)py / P.rubix_unit_px;
                const bool ongrid = fmod(ux, P.rubix_block) < P.rubix_pad || fmod(uy, P.rubix_block) < P.rubix_pad;
                out = 0x80000000u | (ongrid ? 0x40000000u : 0u) | (unsigned)(best * ps * ps + px + py * ps);
            }
        }
    }
    if (c.flag) out |= 0x20000000u;
    cand[(size_t)ly * P.width + lx] = out;
}
)KRN";

struct Nvrtc {
    void *lib = nullptr;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    std::string why;
};

Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
        for (const char *nm : names) {
            n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.lib) break;
        }
        if (!n.lib) {
            n.why = std::string("NVRTC not found: ") + dlerror();
            return;
        }
#define LOAD(sym)                                                          \
    n.sym = reinterpret_cast<decltype(n.sym)>(dlsym(n.lib, "nvrtc" #sym)); \
    if (!n.sym) n.why = "NVRTC lacks nvrtc" #sym;
        LOAD(CreateProgram)
        LOAD(CompileProgram)
        LOAD(GetCUBINSize)
        LOAD(GetCUBIN)
        LOAD(GetProgramLogSize)
        LOAD(GetProgramLog)
        LOAD(DestroyProgram)
        LOAD(GetErrorString)
#undef LOAD
    });
    return n;
}

struct Driver {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
    bool ok = false;
};

Driver &driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        d.ok = get("cuModuleLoadData", reinterpret_cast<void **>(&d.ModuleLoadData)) &&
               get("cuModuleGetFunction", reinterpret_cast<void **>(&d.ModuleGetFunction)) &&
               get("cuModuleUnload", reinterpret_cast<void **>(&d.ModuleUnload)) &&
               get("cuLaunchKernel", reinterpret_cast<void **>(&d.LaunchKernel));
    });
    return d;
}

}  // namespace

struct LensDevice::Module {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

LensDevice::~LensDevice() {
    for (auto &kv : cache_) {
        if (kv.second->mod && driver().ok) driver().ModuleUnload(kv.second->mod);
        delete kv.second;
    }
}

bool LensDevice::compile(const std::string &lens_source, std::vector<char> *cubin, std::string *log) {
    Nvrtc &n = nvrtc();
    if (!n.why.empty()) {
        *log = n.why;
        return false;
    }
    const std::string src = lens_source + kKernelSource;
    nvrtcProgram prog;
    nvrtcResult rc = n.CreateProgram(&prog, src.c_str(), "lens.cu", 0, nullptr, nullptr);
    if (rc != NVRTC_SUCCESS) {
        *log = std::string("nvrtcCreateProgram: ") + n.GetErrorString(rc);
        return false;
    }
    // --fmad=false: the host computes with -ffp-contract=off; parity needs unfused arithmetic
    const char *opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "--disable-warnings"};
    rc = n.CompileProgram(prog, static_cast<int>(sizeof(opts) / sizeof(opts[0])), opts);
    size_t logsz = 0;
    n.GetProgramLogSize(prog, &logsz);
    if (logsz > 1) {
        log->resize(logsz);
        n.GetProgramLog(prog, &(*log)[0]);
    }
    if (rc != NVRTC_SUCCESS) {
        *log = std::string("NVRTC: ") + n.GetErrorString(rc) + "\n" + *log;
        n.DestroyProgram(&prog);
        return false;
    }
    size_t sz = 0;
    n.GetCUBINSize(prog, &sz);
    cubin->resize(sz);
    n.GetCUBIN(prog, cubin->data());
    n.DestroyProgram(&prog);
    return sz > 0;
}

bool LensDevice::build(const std::string &lens_source, const LensBuildParams &p, uint32_t *cand, std::string *err) {
    compile_ms_ = kernel_ms_ = 0;
    if (cudaSetDevice(device_) != cudaSuccess) {
        *err = "cudaSetDevice failed";
        return false;
    }
    Driver &d = driver();
    if (!d.ok) {
        *err = "CUDA driver entry points unavailable";
        return false;
    }
    Module *m = nullptr;
    auto it = cache_.find(lens_source);
    if (it != cache_.end()) {
        m = it->second;
    } else {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<char> cubin;
        std::string log;
        if (!compile(lens_source, &cubin, &log)) {
            *err = log;
            return false;
        }
        cudaFree(nullptr);  // make sure the primary context is current
        m = new Module;
        CUresult cr = d.ModuleLoadData(&m->mod, cubin.data());
        if (cr == CUDA_SUCCESS) cr = d.ModuleGetFunction(&m->fn, m->mod, "lt_build");
        if (cr != CUDA_SUCCESS) {
            if (m->mod) d.ModuleUnload(m->mod);
            delete m;
            *err = "loading the compiled lens failed (CUresult " + std::to_string(static_cast<int>(cr)) + ")";
            return false;
        }
        if (cache_.size() >= 16) {  // lenses are few; keep the cache from growing without bound
            for (auto &kv : cache_) {
                d.ModuleUnload(kv.second->mod);
                delete kv.second;
            }
            cache_.clear();
        }
        cache_[lens_source] = m;
        compile_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }

    const size_t npix = static_cast<size_t>(p.width) * p.height;
    uint32_t *d_cand = nullptr;
    cudaError_t ce = cudaMalloc(&d_cand, npix * sizeof(uint32_t));
    if (ce != cudaSuccess) {
        *err = std::string("cudaMalloc: ") + cudaGetErrorString(ce);
        return false;
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    LensBuildParams params = p;
    void *args[] = {&params, &d_cand};
    const unsigned block = 128;
    cudaEventRecord(e0, nullptr);
    CUresult cr = d.LaunchKernel(m->fn, (p.width + block - 1) / block, static_cast<unsigned>(p.height), 1, block, 1, 1, 0, nullptr, args, nullptr);
    cudaEventRecord(e1, nullptr);
    bool ok = cr == CUDA_SUCCESS;
    if (!ok) *err = "cuLaunchKernel failed (CUresult " + std::to_string(static_cast<int>(cr)) + ")";
    if (ok) {
        ce = cudaMemcpy(cand, d_cand, npix * sizeof(uint32_t), cudaMemcpyDeviceToHost);  // synchronises
        if (ce != cudaSuccess) {
            *err = std::string("lens kernel: ") + cudaGetErrorString(ce);
            ok = false;
        } else {
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            kernel_ms_ = ms;
            ++launches_;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_cand);
    return ok;
}

}  // namespace blinky
