// rt_dsl.cu -- run-time compilation and launch of user kernels (rendering/_core.py:247-299: the reference builds one
// OpenCL program from every @kernel_struct / @kernel_function / @kernel_main at first dispatch, :283-286).
//
// The Python side (rendertoy_b200/rendering/_dsl.py) turns the accumulated OpenCL C text into CUDA C++ over a small
// device prelude; here NVRTC compiles it for sm_100a (no FMA contraction, like the rest of the library) and the CUDA
// runtime's library API loads and launches it.  NVRTC is dlopen'ed on first use so librendertoy_b200.so itself has no
// hard dependency on it.
#include <dlfcn.h>
#include <string>
#include <vector>

#include "rt_common.cuh"

namespace {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;

struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(nvrtcResult) = nullptr;
};

Nvrtc *nvrtc()
{
    static Nvrtc n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            n.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (n.h) {
#define RT_SYM(f) n.f = (decltype(n.f))dlsym(n.h, "nvrtc" #f)
            RT_SYM(CreateProgram); RT_SYM(CompileProgram); RT_SYM(GetProgramLogSize); RT_SYM(GetProgramLog);
            RT_SYM(GetCUBINSize); RT_SYM(GetCUBIN); RT_SYM(DestroyProgram); RT_SYM(GetErrorString);
#undef RT_SYM
            if (!n.CreateProgram || !n.CompileProgram || !n.GetCUBIN) { dlclose(n.h); n.h = nullptr; }
        }
    }
    return n.h ? &n : nullptr;
}

struct rt_dsl_module {
    std::vector<char> cubin;
    cudaLibrary_t lib = nullptr; // loaded on first launch (needs a device); compile alone works without a GPU
};

} // namespace

extern "C" {

// Compile CUDA C++ source to an sm_100a cubin.  On failure the compiler log is copied to `log` (NUL-terminated,
// truncated to log_cap) and RT_ERR_UNSUPPORTED is returned.
int rt_dsl_compile(const char *cuda_source, uint64_t *out_module, char *log, int log_cap)
{
    RT_REQUIRE(cuda_source && out_module, "source / out handle");
    if (log && log_cap > 0) log[0] = 0;
    Nvrtc *n = nvrtc();
    if (!n) {
        rt_set_error("libnvrtc.so.12 could not be loaded: run-time kernels are unavailable");
        return RT_ERR_UNSUPPORTED;
    }
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n->CreateProgram(&prog, cuda_source, "rendertoy_dsl.cu", 0, nullptr, nullptr);
    if (r != 0) { rt_set_error("nvrtcCreateProgram: %s", n->GetErrorString ? n->GetErrorString(r) : "?"); return RT_ERR_UNSUPPORTED; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "--prec-div=true", "--prec-sqrt=true",
                          "--ftz=false", "-lineinfo", "-w"};
    r = n->CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts);
    size_t log_size = 0;
    if (n->GetProgramLogSize && n->GetProgramLogSize(prog, &log_size) == 0 && log_size > 1 && log && log_cap > 1) {
        std::vector<char> full(log_size + 1, 0);
        n->GetProgramLog(prog, full.data());
        snprintf(log, (size_t)log_cap, "%s", full.data());
    }
    if (r != 0) {
        rt_set_error("NVRTC compilation failed: %s", n->GetErrorString ? n->GetErrorString(r) : "?");
        n->DestroyProgram(&prog);
        return RT_ERR_UNSUPPORTED;
    }
    size_t sz = 0;
    n->GetCUBINSize(prog, &sz);
    rt_dsl_module *m = new rt_dsl_module;
    m->cubin.resize(sz);
    n->GetCUBIN(prog, m->cubin.data());
    n->DestroyProgram(&prog);
    *out_module = (uint64_t)(uintptr_t)m;
    return RT_OK;
}

// Launch `kernel` of a compiled module over n_threads work-items (1-D, 128 threads per block; the kernels guard
// thread_id >= number_of_threads themselves, rendering/_core.py:252-253).  args[i] points at the i-th argument's
// value (a device pointer for buffers, the struct/scalar bytes for by-value arguments).
int rt_dsl_launch(uint64_t module, const char *kernel, int64_t n_threads, void **args, void *stream)
{
    RT_REQUIRE(module && kernel && n_threads >= 0, "module / kernel / thread count");
    rt_dsl_module *m = (rt_dsl_module *)(uintptr_t)module;
    if (!m->lib) RT_CUDA(cudaLibraryLoadData(&m->lib, m->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    cudaKernel_t k = nullptr;
    RT_CUDA(cudaLibraryGetKernel(&k, m->lib, kernel));
    if (n_threads == 0) return RT_OK;
    const unsigned block = 128;
    const long long grid = (n_threads + block - 1) / block;
    RT_REQUIRE(grid < (1ll << 31), "too many threads for a 1-D grid");
    RT_CUDA(cudaLaunchKernel((const void *)k, dim3((unsigned)grid), dim3(block), args, 0, (cudaStream_t)stream));
    return RT_OK;
}

int rt_dsl_unload(uint64_t module)
{
    if (!module) return RT_OK;
    rt_dsl_module *m = (rt_dsl_module *)(uintptr_t)module;
    if (m->lib) cudaLibraryUnload(m->lib);
    delete m;
    return RT_OK;
}

} // extern "C"
