// osl_b200_jit.h — NVRTC + driver-API helpers shared by the host translation
// units of libosl_b200.so (product code, internal).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include "host/osl_b200_texture.h"

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
    CUresult_ (*cuModuleGetGlobal)(unsigned long long*, size_t*, CUmodule_, const char*) = nullptr;
    CUresult_ (*cuFuncSetAttribute)(CUfunction_, int, int)                 = nullptr;
    CUresult_ (*cuFuncGetAttribute)(int*, int, CUfunction_)                = nullptr;
    CUresult_ (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction_, int, size_t) = nullptr;
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
        OSLB200_SYM(cuFuncSetAttribute)
        OSLB200_SYM(cuFuncGetAttribute)
        OSLB200_SYM(cuOccupancyMaxActiveBlocksPerMultiprocessor)
        *(void**)(&d.cuModuleGetGlobal) = dlsym(d.lib, "cuModuleGetGlobal_v2");
        if (!d.cuModuleGetGlobal) {
            d.why = "libcuda is missing symbol cuModuleGetGlobal_v2";
            return;
        }
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

// Uploads the images a module's texture() calls name to the CURRENT device and fills the
// module's `osl_tex_` table (device/osl_b200_texture.cuh).  Returns "" or the error.
// `allocations` receives the device buffers (owned by the caller, freed with cudaFree).
inline std::string
bind_module_textures_impl(CUmodule_ mod, const std::vector<std::string>& names, const std::string& searchpath,
                     std::vector<void*>& allocations)
{
    if (names.empty())
        return "";
    Driver& d = jit_driver();
    if (!d.ok)
        return "CUDA driver unavailable: " + d.why;
    unsigned long long dptr = 0;
    size_t bytes            = 0;
    CUresult_ r             = d.cuModuleGetGlobal(&dptr, &bytes, mod, "osl_tex_");
    if (r != 0)
        return "cuModuleGetGlobal(osl_tex_): " + d.err(r);
    std::vector<TexDescHost> table(names.size());
    if (bytes < table.size() * sizeof(TexDescHost))
        return "texture table of the module is smaller than its texture list";
    for (size_t i = 0; i < names.size(); ++i) {
        std::string err;
        const TextureImage* im = texture_get(names[i], searchpath, err);
        if (!im)
            return err;
        void* p   = nullptr;
        size_t nb = im->rgba.size() * sizeof(float);
        if (cudaMalloc(&p, nb) != cudaSuccess || cudaMemcpy(p, im->rgba.data(), nb, cudaMemcpyHostToDevice) != cudaSuccess)
            return "uploading texture '" + names[i] + "': " + cudaGetErrorString(cudaGetLastError());
        allocations.push_back(p);
        table[i] = TexDescHost { p, im->w, im->h, im->nch, 0 };
    }
    if (cudaMemcpy((void*)dptr, table.data(), table.size() * sizeof(TexDescHost), cudaMemcpyHostToDevice) != cudaSuccess)
        return std::string("filling osl_tex_: ") + cudaGetErrorString(cudaGetLastError());
    return "";
}

}  // namespace oslb200
