// osl_b200_jit.h — NVRTC + driver-API helpers shared by the host translation
// units of libosl_b200.so (product code, internal).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

extern const int osl_b200_num_device_headers;
extern const char* const osl_b200_device_header_names[];
extern const char* const osl_b200_device_header_sources[];

namespace oslb200 {

typedef int CUresult_;
typedef void* CUmodule_;
typedef void* CUfunction_;

struct Driver {
    void* lib                                                              = nullptr;
    CUresult_ (*cuInit)(unsigned)                                          = nullptr;
    CUresult_ (*cuModuleLoadData)(CUmodule_*, const void*)                 = nullptr;
    CUresult_ (*cuModuleUnload)(CUmodule_)                                 = nullptr;
    CUresult_ (*cuModuleGetFunction)(CUfunction_*, CUmodule_, const char*) = nullptr;
    CUresult_ (*cuLaunchKernel)(CUfunction_, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                void*, void**, void**)                     = nullptr;
    CUresult_ (*cuGetErrorString)(CUresult_, const char**)                 = nullptr;
    bool ok = false;
    std::string why;
    std::string err(CUresult_ r) const
    {
        const char* s = nullptr;
        if (cuGetErrorString)
            cuGetErrorString(r, &s);
        return s ? s : ("CUDA driver error " + std::to_string(r));
    }
};

// libcuda is resolved at run time so the library loads (and JIT-compiles) on
// machines without a GPU; executing without a driver is a hard error.
inline Driver&
jit_driver()
{
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!d.lib)
            d.lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
        if (!d.lib) {
            d.why = "libcuda.so.1 not found: no NVIDIA driver on this machine";
            return;
        }
#define OSLB200_SYM(n)                                           \
    *(void**)(&d.n) = dlsym(d.lib, #n);                          \
    if (!d.n) {                                                  \
        d.why = std::string("libcuda is missing symbol ") + #n; \
        return;                                                  \
    }
        OSLB200_SYM(cuInit)
        OSLB200_SYM(cuModuleLoadData)
        OSLB200_SYM(cuModuleUnload)
        OSLB200_SYM(cuModuleGetFunction)
        OSLB200_SYM(cuLaunchKernel)
        OSLB200_SYM(cuGetErrorString)
#undef OSLB200_SYM
        if (d.cuInit(0) != 0) {
            d.why = "cuInit failed";
            return;
        }
        d.ok = true;
    });
    return d;
}

// CUDA C++ text -> sm_100a cubin.  Returns "" on success, else the error + log.
inline std::string
jit_compile(const std::string& src, const char* name, bool fma, std::vector<char>& cubin)
{
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), name, osl_b200_num_device_headers, osl_b200_device_header_sources,
                           osl_b200_device_header_names)
        != NVRTC_SUCCESS)
        return "nvrtcCreateProgram failed";
    std::vector<const char*> opts = { "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--prec-div=true",
                                      "--prec-sqrt=true", "--ftz=false", "-default-device" };
    opts.push_back(fma ? "--fmad=true" : "--fmad=false");
    nvrtcResult r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    std::string log;
    size_t logsz = 0;
    nvrtcGetProgramLogSize(prog, &logsz);
    if (logsz > 1) {
        log.resize(logsz);
        nvrtcGetProgramLog(prog, &log[0]);
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return "NVRTC: " + std::string(nvrtcGetErrorString(r)) + "\n" + log;
    }
    size_t sz = 0;
    if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
        nvrtcDestroyProgram(&prog);
        return "NVRTC produced no cubin";
    }
    cubin.resize(sz);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    return "";
}

}  // namespace oslb200
