// osl_b200_cabi.cu — C ABI, NVRTC JIT, kernel launch and host<->device staging
// for libosl_b200.so (product code).
//
// Stands in for the reference's JIT + execute plumbing:
//   BackendLLVM::run / JIT            src/liboslexec/llvm_instance.cpp:2083-2572
//   ShadingContext::execute*          src/liboslexec/context.cpp:92-263, 268-440
//   BatchedExecutor<W>::execute       src/include/OSL/oslexec.h:982-1033
// The generated CUDA text is compiled with NVRTC straight to an sm_100a cubin
// (no PTX JIT at load time), loaded through the driver API (resolved with
// dlopen so the library itself loads on machines without a GPU) and launched
// with one thread per shading point.
#include "../../include/osl_b200.h"
#include "host/osl_b200_group.h"
#include "host/osl_b200_texture.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include "device/osl_b200_device.cuh"

// device header texts (generated at build from csrc/device/*.cuh)
extern const int osl_b200_num_device_headers;
extern const char* const osl_b200_device_header_names[];
extern const char* const osl_b200_device_header_sources[];

using namespace oslb200;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches { 0 };

int
fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

// ---- driver API via dlopen ---------------------------------------------------
typedef int CUresult_;
typedef void* CUmodule_;
typedef void* CUfunction_;
struct Driver {
    void* lib = nullptr;
    CUresult_ (*cuInit)(unsigned)                                                          = nullptr;
    CUresult_ (*cuModuleLoadData)(CUmodule_*, const void*)                                 = nullptr;
    CUresult_ (*cuModuleUnload)(CUmodule_)                                                 = nullptr;
    CUresult_ (*cuModuleGetFunction)(CUfunction_*, CUmodule_, const char*)                 = nullptr;
    CUresult_ (*cuLaunchKernel)(CUfunction_, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
                                unsigned, void*, void**, void**)                           = nullptr;
    CUresult_ (*cuGetErrorString)(CUresult_, const char**)                                 = nullptr;
    CUresult_ (*cuOccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction_, int, size_t) = nullptr;
    bool ok = false;
    std::string why;
};
Driver&
driver()
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
#define LOADSYM(n)                                                   \
    *(void**)(&d.n) = dlsym(d.lib, #n);                              \
    if (!d.n) {                                                      \
        d.why = std::string("libcuda is missing symbol ") + #n;     \
        return;                                                      \
    }
        LOADSYM(cuInit)
        LOADSYM(cuModuleLoadData)
        LOADSYM(cuModuleUnload)
        LOADSYM(cuModuleGetFunction)
        LOADSYM(cuLaunchKernel)
        LOADSYM(cuGetErrorString)
        LOADSYM(cuOccupancyMaxActiveBlocksPerMultiprocessor)
#undef LOADSYM
        if (d.cuInit(0) != 0) {
            d.why = "cuInit failed";
            return;
        }
        d.ok = true;
    });
    return d;
}
std::string
cu_err(CUresult_ r)
{
    const char* s = nullptr;
    if (driver().cuGetErrorString)
        driver().cuGetErrorString(r, &s);
    return s ? s : ("CUDA driver error " + std::to_string(r));
}

struct LaunchBlock {  // must match B200Launch in the generated code
    const float* varying[B200_SG_NFIELDS];
    float uniform[B200_SG_NFIELDS][4];
    long long plane_stride;
    const int* shadeindex;
    void* output_base;
    const void* userdata_base;
    long long npoints;
    long long shadeindex_base;
    long long out_adjust[B200_MAX_OUTPUTS];
    int stage_outputs;
    int pad_;
    float xf[B200_MAX_SPACES][2][16];
    int xf_ok[B200_MAX_SPACES];
    unsigned* journal;
    unsigned journal_words;
    unsigned pad2_;
};

}  // namespace

// shared with the other host translation units of the library
namespace oslb200 {
int
set_error(int code, const std::string& msg)
{
    return fail(code, msg);
}
void
count_launches(long long n)
{
    g_launches.fetch_add(n);
}
}  // namespace oslb200

struct b200_group {
    Group g;
    std::string source;
    std::vector<char> cubin;
    int block = 256;
    int gridcap = 0;        // option gridcap=N: at most N CTAs per SM in the grid (0: waves x the measured residency)
    int waves = 12;         // option waves=N: grid = at most N full waves of resident CTAs; CTAs loop over the remaining tiles
    int minblocks = 0;      // option minblocks=N: resident CTAs per SM the register allocator must allow (0: its own choice)
    int stage_outputs = -1; // option stage=0|1: direct stores / shared-memory staging + TMA bulk stores (-1: by record size)
    std::mutex mu;
    std::map<int, std::pair<CUmodule_, CUfunction_>> loaded;  // per device
    int sm_count[64] = { 0 };
    int resident[64] = { 0 };  // CTAs of the kernel one SM holds (cuOccupancyMaxActiveBlocksPerMultiprocessor)
    std::vector<void*> texture_allocs;  // device images of the textures the group reads (all devices)
    // printf journal (only for groups that have printf sites): one device buffer per device
    unsigned journal_words = 4u << 20;   // option journal=WORDS
    std::map<int, unsigned*> journal_dev;
    std::string journal_text;
    // staging buffers for execute_host: one set per device (its streams belong to that device's
    // context), grown on demand.  host_mu serialises execute_host callers of this group: the
    // reference lets many threads execute one group, each with its own context; here they
    // share the group's staging area, so they take turns (the device path has no such state).
    struct Stage {
        char* d_in       = nullptr;
        char* d_out      = nullptr;
        char* d_ud       = nullptr;
        size_t in_bytes  = 0, out_bytes = 0, ud_bytes = 0;
        cudaStream_t streams[3] = { nullptr, nullptr, nullptr };
    };
    std::map<int, Stage> stages;
    std::mutex host_mu;
    long long userdata_record = 0;   // option userdata=record: bytes per point
    bool error_repeats = false;      // option error_repeats=1: report identical error()/warning() texts again
    std::set<std::string> messages_seen;
};

static std::map<std::string, std::string>
parse_options(const char* s)
{
    std::map<std::string, std::string> m;
    if (!s)
        return m;
    std::istringstream in(s);
    std::string kv;
    while (std::getline(in, kv, ',')) {
        size_t e = kv.find('=');
        if (e == std::string::npos)
            m[kv] = "1";
        else
            m[kv.substr(0, e)] = kv.substr(e + 1);
    }
    return m;
}

static int
nvrtc_compile(b200_group* G, std::string& log)
{
    std::string src = G->source;
    for (size_t p; (p = src.find("%BLOCK%")) != std::string::npos;)
        src.replace(p, 7, std::to_string(G->block));
    for (size_t p; (p = src.find("%MINBLOCKS%")) != std::string::npos;)
        src.replace(p, 11, G->minblocks > 0 ? ", " + std::to_string(G->minblocks) : std::string());
    G->source = src;
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, src.c_str(), "osl_b200_group.cu", osl_b200_num_device_headers,
                           osl_b200_device_header_sources, osl_b200_device_header_names)
        != NVRTC_SUCCESS)
        return fail(B200_ERR_COMPILE, "nvrtcCreateProgram failed");
    std::vector<const char*> opts = { "--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo",
                                      "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
                                      "-default-device" };
    opts.push_back(G->g.fma ? "--fmad=true" : "--fmad=false");
    nvrtcResult r = nvrtcCompileProgram(prog, (int)opts.size(), opts.data());
    size_t logsz  = 0;
    nvrtcGetProgramLogSize(prog, &logsz);
    if (logsz > 1) {
        log.resize(logsz);
        nvrtcGetProgramLog(prog, &log[0]);
    }
    if (r != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return fail(B200_ERR_COMPILE, "NVRTC: " + std::string(nvrtcGetErrorString(r)) + "\n" + log);
    }
    size_t sz = 0;
    if (nvrtcGetCUBINSize(prog, &sz) != NVRTC_SUCCESS || sz == 0) {
        nvrtcDestroyProgram(&prog);
        return fail(B200_ERR_COMPILE, "NVRTC produced no cubin");
    }
    G->cubin.resize(sz);
    nvrtcGetCUBIN(prog, G->cubin.data());
    nvrtcDestroyProgram(&prog);
    return B200_OK;
}

extern "C" {

int
b200_abi_version(void)
{
    return OSL_B200_ABI_VERSION;
}
const char*
b200_last_error(void)
{
    return g_err.c_str();
}
long long
b200_launch_count(void)
{
    return g_launches.load();
}

int
b200_group_compile(const b200_group_desc* desc, b200_group** out)
{
    if (!desc || !out || desc->nlayers <= 0 || !desc->layers)
        return fail(B200_ERR_INVALID, "b200_group_compile: null or empty group description");
    *out = nullptr;
    std::unique_ptr<b200_group> G(new b200_group);
    try {
        G->g.name = desc->name ? desc->name : "group";
        auto opt  = parse_options(desc->options);
        if (opt.count("fma"))
            G->g.fma = atoi(opt["fma"].c_str()) != 0;
        if (opt.count("colorspace"))
            G->g.colorspace = opt["colorspace"];
        if (opt.count("journal")) {
            // journal=1 asks for the default size, any larger value is the size in words
            unsigned long jw     = strtoul(opt["journal"].c_str(), nullptr, 10);
            G->g.journal_enabled = jw > 0;
            if (jw > 1)
                G->journal_words = (unsigned)jw;
        }
        if (opt.count("minblocks"))
            G->minblocks = std::max(0, atoi(opt["minblocks"].c_str()));
        if (opt.count("gridcap"))
            G->gridcap = std::max(1, atoi(opt["gridcap"].c_str()));
        if (opt.count("waves"))
            G->waves = std::max(1, atoi(opt["waves"].c_str()));
        if (opt.count("error_repeats"))
            G->error_repeats = atoi(opt["error_repeats"].c_str()) != 0;
        if (opt.count("block"))
            G->block = atoi(opt["block"].c_str());
        if (opt.count("stage"))
            G->stage_outputs = atoi(opt["stage"].c_str()) != 0;
        if (opt.count("texturepath"))
            G->g.texturepath = opt["texturepath"];
        if (G->block < 32 || G->block > 1024 || (G->block % 32))
            return fail(B200_ERR_INVALID, "option block must be a multiple of 32 in [32,1024]");
        for (int i = 0; i < desc->nlayers; ++i) {
            const b200_layer& l = desc->layers[i];
            if (!l.oso_text || !l.layername)
                return fail(B200_ERR_INVALID, "layer without oso text or name");
            std::vector<ParamValue> pvs;
            for (int p = 0; p < l.nparams; ++p) {
                const b200_param& bp = l.params[p];
                ParamValue pv;
                pv.name = bp.name ? bp.name : "";
                for (int k = 0; k < bp.nvalues; ++k) {
                    if (bp.type == 0)
                        pv.ivals.push_back(((const int*)bp.values)[k]);
                    else if (bp.type == 1)
                        pv.fvals.push_back(((const float*)bp.values)[k]);
                    else if (bp.type == 2)
                        pv.svals.push_back(((const char* const*)bp.values)[k]);
                    else
                        return fail(B200_ERR_INVALID, "bad parameter type code");
                }
                pvs.push_back(std::move(pv));
            }
            G->g.add_layer(l.oso_text, l.layername, pvs);
        }
        for (int i = 0; i < desc->nconnections; ++i) {
            const b200_connection& c = desc->connections[i];
            G->g.connect(c.srclayer, c.srcparam, c.dstlayer, c.dstparam);
        }
        for (int i = 0; i < desc->noutputs; ++i) {
            const b200_symloc& s = desc->outputs[i];
            G->g.add_output(s.name, s.offset, s.stride, s.derivs != 0);
        }
        for (int i = 0; i < desc->nattributes && desc->attributes; ++i) {
            const b200_attribute& a = desc->attributes[i];
            if (!a.name || a.nvalues < 0 || a.type < 0 || a.type > 2 || (a.nvalues && !a.values))
                return fail(B200_ERR_INVALID, "bad attribute description");
            Attribute at;
            at.name = a.name;
            at.type = a.type;
            for (int k = 0; k < a.nvalues; ++k) {
                if (a.type == 0)
                    at.ivals.push_back(((const int*)a.values)[k]);
                else if (a.type == 1)
                    at.fvals.push_back(((const float*)a.values)[k]);
                else
                    at.svals.push_back(((const char* const*)a.values)[k] ? ((const char* const*)a.values)[k] : "");
            }
            G->g.attributes.push_back(at);
        }
        for (int i = 0; i < desc->nuserdata && desc->userdata; ++i) {
            const b200_userdata& u = desc->userdata[i];
            if (!u.name || (u.ncomp != 1 && u.ncomp != 3) || u.stride < 0 || (u.is_int && u.ncomp != 1))
                return fail(B200_ERR_INVALID, "bad userdata description");
            UserData d;
            d.name = u.name; d.ncomp = u.ncomp; d.is_int = u.is_int != 0; d.derivs = u.derivs != 0;
            d.offset = u.offset; d.stride = u.stride; d.valid_offset = u.valid_offset; d.valid_stride = u.valid_stride;
            G->g.userdata.push_back(d);
        }
        G->g.block = G->block;
        G->g.finalize();
        if (opt.count("userdata") && opt["userdata"] == "record") {
            // The library lays the userdata out itself: one record per point holding, for every
            // interpolated parameter of the group, a validity word and the value with derivatives
            // (what RendererServices::get_userdata(derivatives=true, ...) fills).  The caller asks
            // b200_group_userdata_field() where each field went and passes records at execute.
            long long off = 0;
            for (Layer& l : G->g.layers)
                for (Symbol& sy : l.m.syms) {
                    if (!sy.is_param() || !sy.interpolated || sy.conn_layer >= 0 || sy.type.arraylen)
                        continue;
                    const bool is_int = sy.type.base == Base::Int;
                    if (!is_int && !(sy.type.base == Base::Float || sy.type.is_triple()))
                        continue;
                    bool dup = false;
                    for (const UserData& u : G->g.userdata)
                        dup |= u.name == sy.name && u.ncomp == sy.type.ncomp() && u.is_int == is_int;
                    if (dup)
                        continue;
                    UserData d;
                    d.name = sy.name; d.ncomp = sy.type.ncomp(); d.is_int = is_int; d.derivs = !is_int;
                    d.valid_offset = off;
                    d.offset       = off + 4;
                    off += 4 + 4 * d.ncomp * (d.derivs ? 3 : 1);
                    G->g.userdata.push_back(d);
                }
            for (UserData& u : G->g.userdata)
                u.stride = u.valid_stride = off;
            G->userdata_record = off;
        }
        G->source = generate_cuda(G->g);
    } catch (const std::exception& e) {
        return fail(B200_ERR_COMPILE, e.what());
    }
    std::string log;
    int rc = nvrtc_compile(G.get(), log);
    if (rc != B200_OK)
        return rc;
    *out = G.release();
    return B200_OK;
}

void
b200_group_destroy(b200_group* g)
{
    if (!g)
        return;
    for (auto& kv : g->loaded)
        if (driver().ok)
            driver().cuModuleUnload(kv.second.first);
    for (void* p : g->texture_allocs)
        cudaFree(p);
    for (auto& kv : g->stages) {
        b200_group::Stage& st = kv.second;
        cudaSetDevice(kv.first);
        if (st.d_in) cudaFree(st.d_in);
        if (st.d_out) cudaFree(st.d_out);
        if (st.d_ud) cudaFree(st.d_ud);
        for (cudaStream_t q : st.streams)
            if (q) cudaStreamDestroy(q);
    }
    for (auto& kv : g->journal_dev)
        if (kv.second)
            cudaFree(kv.second);
    delete g;
}

// ---- printf journal: decode + format on the host -----------------------------------
// One conversion per argument component, components and array elements separated by a
// blank, ints through %d-family conversions print as ints and otherwise as floats, floats
// through %d/%i/%x print truncated (llvm_gen_printf, llvm_gen.cpp:350-700).
static void
journal_format(const Group& g, const JournalFormat& jf, const unsigned* w, size_t nw, std::string& out)
{
    const std::string& f = jf.fmt;
    size_t ai = 0, wi = 0;
    char buf[1100];
    for (size_t i = 0; i < f.size();) {
        if (f[i] != '%') {
            out += f[i++];
            continue;
        }
        if (i + 1 < f.size() && f[i + 1] == '%') {
            out += '%';
            i += 2;
            continue;
        }
        size_t j = i + 1;
        while (j < f.size() && !strchr("cdefgimopsuxXEG", f[j]))
            ++j;
        std::string spec = f.substr(i, j + 1 - i);
        char conv        = j < f.size() ? f[j] : 'g';
        i                = j + 1;
        if (ai >= jf.args.size())
            continue;
        const JournalArg& a = jf.args[ai++];
        int nel             = a.arraylen ? a.arraylen : 1;
        bool first          = true;
        for (int e = 0; e < nel; ++e)
            for (int c = 0; c < a.ncomp; ++c, ++wi) {
                if (wi >= nw)
                    return;
                if (!first)
                    out += ' ';
                first = false;
                if (a.base == Base::String) {
                    std::string s = spec.substr(0, spec.size() - 1) + "s";
                    unsigned id   = w[wi];
                    snprintf(buf, sizeof buf, s.c_str(), id < g.strings.size() ? g.strings[id].c_str() : "");
                } else if (a.base == Base::Int) {
                    int v = (int)w[wi];
                    if (strchr("dioxXuc", conv))
                        snprintf(buf, sizeof buf, spec.c_str(), v);
                    else
                        snprintf(buf, sizeof buf, spec.c_str(), (double)v);
                } else {
                    float v;
                    memcpy(&v, &w[wi], 4);
                    if (conv == 'd' || conv == 'i' || conv == 'x' || conv == 'X')
                        snprintf(buf, sizeof buf, spec.c_str(), (int)v);
                    else
                        snprintf(buf, sizeof buf, spec.c_str(), (double)v);
                }
                out += buf;
            }
    }
}

/* Text printed by the group's printf ops since the previous call, ordered by shade index
 * and, within a point, by execution order - what single-threaded testshade prints.
 * Synchronises `device`.  The pointer stays valid until the next call on this group. */
int
b200_texture_add(const char* name, int width, int height, int nchannels, const float* pixels)
{
    if (!name || !pixels || width <= 0 || height <= 0 || nchannels < 1 || nchannels > 4)
        return fail(B200_ERR_INVALID, "b200_texture_add: bad arguments");
    texture_add(name, width, height, nchannels, pixels);
    return B200_OK;
}

const char*
b200_group_journal(b200_group* g, int device)
{
    if (!g)
        return "";
    g->journal_text.clear();
    auto it = g->journal_dev.find(device);
    if (it == g->journal_dev.end() || !it->second)
        return g->journal_text.c_str();
    cudaSetDevice(device);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        fail(B200_ERR_CUDA, std::string("journal: ") + cudaGetErrorString(cudaGetLastError()));
        return g->journal_text.c_str();
    }
    unsigned head[2] = { 0, 0 };
    if (cudaMemcpy(head, it->second, sizeof head, cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail(B200_ERR_CUDA, std::string("journal read-back: ") + cudaGetErrorString(cudaGetLastError()));
        return g->journal_text.c_str();
    }
    size_t used = head[0];
    if (used > (size_t)g->journal_words - 2)
        used = (size_t)g->journal_words - 2;
    std::vector<unsigned> rec(used);
    if (used && cudaMemcpy(rec.data(), it->second + 2, used * sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail(B200_ERR_CUDA, std::string("journal read-back: ") + cudaGetErrorString(cudaGetLastError()));
        return g->journal_text.c_str();
    }
    struct Ref {
        unsigned si, seq;
        size_t at;
    };
    std::vector<Ref> refs;
    for (size_t p = 0; p + 4 <= used;) {
        unsigned n = rec[p];
        if (n < 4 || p + n > used)
            break;  // unwritten space: a record that did not fit
        refs.push_back({ rec[p + 1], rec[p + 2], p });
        p += n;
    }
    std::stable_sort(refs.begin(), refs.end(),
                     [](const Ref& a, const Ref& b) { return a.si != b.si ? (int)a.si < (int)b.si : a.seq < b.seq; });
    for (const Ref& r : refs) {
        unsigned fmt = rec[r.at + 3];
        if (fmt >= g->g.jformats.size())
            continue;
        const JournalFormat& jf = g->g.jformats[fmt];
        if (jf.kind == 0) {
            journal_format(g->g, jf, rec.data() + r.at + 4, rec[r.at] - 4, g->journal_text);
            continue;
        }
        // error() / warning(): what the reference's error handler prints for testshade, a
        // message reported before is dropped (attribute error_repeats = 0; option error_repeats=1)
        std::string msg = jf.kind == 3 ? "ERROR: " : (jf.kind == 1 ? "ERROR: Shader error [" : "WARNING: Shader warning [");
        if (jf.kind != 3)
            msg += jf.shadername + "]: ";
        journal_format(g->g, jf, rec.data() + r.at + 4, rec[r.at] - 4, msg);
        msg += "\n";
        if (!g->error_repeats) {
            if (g->messages_seen.count(msg))
                continue;
            g->messages_seen.insert(msg);
        }
        g->journal_text += msg;
    }
    if (head[1])
        g->journal_text += "[journal overflow: output truncated; raise the group option journal=WORDS]\n";
    if (cudaMemset(it->second, 0, sizeof(unsigned) * (2 + used)) != cudaSuccess)
        fail(B200_ERR_CUDA, std::string("journal reset: ") + cudaGetErrorString(cudaGetLastError()));
    return g->journal_text.c_str();
}

const char*
b200_group_cuda_source(const b200_group* g)
{
    return g ? g->source.c_str() : "";
}
const void*
b200_group_cubin(const b200_group* g, long long* size)
{
    if (size)
        *size = g ? (long long)g->cubin.size() : 0;
    return g ? g->cubin.data() : nullptr;
}
int
b200_group_num_warnings(const b200_group* g)
{
    return g ? (int)g->g.warnings.size() : 0;
}
const char*
b200_group_warning(const b200_group* g, int i)
{
    return (g && i >= 0 && i < (int)g->g.warnings.size()) ? g->g.warnings[i].c_str() : "";
}
int
b200_group_reads_global(const b200_group* g, int field)
{
    return (g && g->g.globals_read.count(field)) ? 1 : 0;
}

static int
ensure_loaded(b200_group* g, int device, CUfunction_* fn)
{
    std::lock_guard<std::mutex> lock(g->mu);
    auto it = g->loaded.find(device);
    if (it != g->loaded.end()) {
        *fn = it->second.second;
        return B200_OK;
    }
    Driver& d = driver();
    if (!d.ok)
        return fail(B200_ERR_CUDA, "CUDA driver unavailable: " + d.why);
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess)
        ce = cudaFree(0);  // make the primary context current
    if (ce != cudaSuccess)
        return fail(B200_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
    CUmodule_ mod;
    CUresult_ r = d.cuModuleLoadData(&mod, g->cubin.data());
    if (r != 0)
        return fail(B200_ERR_CUDA, "cuModuleLoadData: " + cu_err(r));
    CUfunction_ f;
    r = d.cuModuleGetFunction(&f, mod, "osl_b200_group_kernel");
    if (r != 0) {
        d.cuModuleUnload(mod);
        return fail(B200_ERR_CUDA, "cuModuleGetFunction: " + cu_err(r));
    }
    std::string terr = bind_module_textures(mod, g->g.textures, g->g.texturepath, g->texture_allocs);
    if (!terr.empty()) {
        d.cuModuleUnload(mod);
        return fail(B200_ERR_INVALID, terr);
    }
    g->loaded[device] = { mod, f };
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (device < 64) {
        g->sm_count[device] = sms;
        int nb = 0;
        if (d.cuOccupancyMaxActiveBlocksPerMultiprocessor
            && d.cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, g->block, 0) == 0 && nb > 0)
            g->resident[device] = nb;
    }
    *fn = f;
    return B200_OK;
}

static int
launch_group(b200_group* g, int device, void* stream, long long npoints, const b200_globals* sg,
             const int* shadeindex, const void* userdata_base, void* output_base, long long sidx_base,
             const long long* out_adjust)
{
    CUfunction_ fn;
    int rc = ensure_loaded(g, device, &fn);
    if (rc != B200_OK)
        return rc;
    // the kernel belongs to the primary context of `device`: make it current for the launch
    // (the caller's own current device is put back afterwards)
    struct DeviceScope {
        int prev = -1;
        explicit DeviceScope(int d)
        {
            if (cudaGetDevice(&prev) != cudaSuccess)
                prev = -1;
            if (prev != d)
                cudaSetDevice(d);
            else
                prev = -1;
        }
        ~DeviceScope()
        {
            if (prev >= 0)
                cudaSetDevice(prev);
        }
    } scope(device);
    LaunchBlock L;
    memcpy(L.varying, sg->varying, sizeof L.varying);
    memcpy(L.uniform, sg->uniform, sizeof L.uniform);
    L.plane_stride    = sg->plane_stride;
    L.shadeindex      = shadeindex;
    L.output_base     = output_base;
    L.userdata_base   = userdata_base;
    L.npoints         = npoints;
    L.shadeindex_base = sidx_base;
    for (int k = 0; k < B200_MAX_OUTPUTS; ++k)
        L.out_adjust[k] = out_adjust ? out_adjust[k] : 0;
    // staging pays for records of two or more 16-byte rows (48 B layers record: 4x); a 12 B colour
    // record is already written in three nearly-full lines per warp and the staging only adds latency
    L.stage_outputs = g->stage_outputs >= 0 ? g->stage_outputs : (g->g.stage_record_bytes >= 32 ? 1 : 0);
    L.pad_          = 0;
    // named coordinate systems: resolve the names the generated code references and invert
    // once per launch (the reference inverts per call: rs_get_inverse_matrix_*)
    for (int k = 0; k < B200_MAX_SPACES; ++k) {
        osld::M44 m = osld::m44_diag(1.0f), mi = m;
        L.xf_ok[k]  = 0;
        if (k < (int)g->g.spaces.size()) {
            for (int t = 0; t < sg->ntransforms && sg->transforms; ++t)
                if (sg->transforms[t].name && g->g.spaces[k] == sg->transforms[t].name) {
                    m          = osld::m44_load(sg->transforms[t].m);
                    mi         = osld::m44_inverse(m);
                    L.xf_ok[k] = 1;
                    break;
                }
        }
        memcpy(L.xf[k][0], m.x, sizeof m.x);
        memcpy(L.xf[k][1], mi.x, sizeof mi.x);
    }
    L.journal       = nullptr;
    L.journal_words = 0;
    L.pad2_         = 0;
    if (!g->g.jformats.empty() && g->journal_words > 16) {
        std::lock_guard<std::mutex> lk(g->mu);
        unsigned*& jb = g->journal_dev[device];
        if (!jb) {
            cudaSetDevice(device);
            if (cudaMalloc(&jb, sizeof(unsigned) * (size_t)g->journal_words) != cudaSuccess
                || cudaMemset(jb, 0, sizeof(unsigned) * (size_t)g->journal_words) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMalloc(journal) failed");
        }
        L.journal       = jb;
        L.journal_words = g->journal_words;
    }
    int sms = (device >= 0 && device < 64 && g->sm_count[device]) ? g->sm_count[device] : 148;
    // grid: one CTA per tile of `block` points, capped at `waves` full waves of resident CTAs
    // (SMs x the kernel's measured residency: a cap that is not a multiple of the residency
    // leaves the last wave partly empty - 8 CTAs/SM against 5 resident cost 20 %); CTAs loop
    // over the remaining tiles.  Sweep: profiles/group_tune_r02.txt
    long long want = (npoints + g->block - 1) / g->block;
    const int res_ = (device >= 0 && device < 64 && g->resident[device]) ? g->resident[device] : 4;
    long long cap  = g->gridcap > 0 ? (long long)sms * g->gridcap : (long long)sms * res_ * g->waves;
    unsigned grid  = (unsigned)(want < cap ? want : cap);
    void* args[]   = { &L };
    CUresult_ r    = driver().cuLaunchKernel(fn, grid, 1, 1, (unsigned)g->block, 1, 1, 0, stream, args, nullptr);
    if (r != 0)
        return fail(B200_ERR_CUDA, "cuLaunchKernel: " + cu_err(r));
    g_launches.fetch_add(1);
    return B200_OK;
}

int
b200_group_execute(b200_group* g, int device, void* stream, long long npoints, const b200_globals* sg,
                   const int* shadeindex, const void* userdata_base, void* output_base)
{
    if (!g || !sg || npoints < 0)
        return fail(B200_ERR_INVALID, "b200_group_execute: bad arguments");
    if (npoints == 0)
        return B200_OK;
    return launch_group(g, device, stream, npoints, sg, shadeindex, userdata_base, output_base, 0, nullptr);
}

// Host-pointer path: stage only the planes the kernel reads, pipelined in
// chunks over three streams so H2D, kernel and D2H overlap.
static int
execute_host_impl(b200_group* g, int device, long long npoints, const b200_globals* sg, const void* userdata_host,
                  long long userdata_bytes, void* output_base, long long first)
{
    if (!g || !sg || npoints < 0 || first < 0)
        return fail(B200_ERR_INVALID, "b200_group_execute_host: bad arguments");
    if (npoints == 0)
        return B200_OK;
    CUfunction_ fn;
    int rc = ensure_loaded(g, device, &fn);
    if (rc != B200_OK)
        return rc;
    std::lock_guard<std::mutex> host_lock(g->host_mu);
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    cudaSetDevice(device);
    struct Restore {
        int d;
        ~Restore()
        {
            if (d >= 0)
                cudaSetDevice(d);
        }
    } restore { prev_device == device ? -1 : prev_device };
    struct Plane {
        int field, comps;
    };
    std::vector<Plane> planes;
    static const bool is_triple[B200_SG_NFIELDS]
        = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0 };
    size_t in_bpp = 0;
    for (int f = 0; f < B200_SG_NFIELDS; ++f)
        if (sg->varying[f] && g->g.globals_read.count(f)) {
            planes.push_back({ f, is_triple[f] ? 3 : 1 });
            in_bpp += 4 * (is_triple[f] ? 3 : 1);
        }
    typedef OutCluster Cluster;
    const std::vector<OutCluster>& clusters = g->g.clusters;
    size_t out_bpp = 0;
    for (const Cluster& c : clusters)
        out_bpp += (size_t)c.stride;
    if (!output_base && out_bpp)
        return fail(B200_ERR_INVALID, "execute_host: group has outputs but output_base is null");
    const long long CH    = 1 << 20;  // points per chunk
    const int NS          = 3;
    long long chunk       = npoints < CH ? npoints : CH;
    b200_group::Stage& st = g->stages[device];
    size_t need_in = (size_t)chunk * in_bpp * NS, need_out = (size_t)chunk * out_bpp * NS;
    if (st.in_bytes < need_in) {
        if (st.d_in) cudaFree(st.d_in);
        st.d_in = nullptr; st.in_bytes = 0;
        if (cudaMalloc(&st.d_in, need_in) != cudaSuccess)
            return fail(B200_ERR_CUDA, "cudaMalloc(staging in) failed");
        st.in_bytes = need_in;
    }
    if (st.out_bytes < need_out) {
        if (st.d_out) cudaFree(st.d_out);
        st.d_out = nullptr; st.out_bytes = 0;
        if (cudaMalloc(&st.d_out, need_out) != cudaSuccess)
            return fail(B200_ERR_CUDA, "cudaMalloc(staging out) failed");
        st.out_bytes = need_out;
    }
    for (int s = 0; s < NS; ++s)
        if (!st.streams[s] && cudaStreamCreateWithFlags(&st.streams[s], cudaStreamNonBlocking) != cudaSuccess)
            return fail(B200_ERR_CUDA, "cudaStreamCreate failed");
    // userdata arena: uploaded whole, once (the kernel indexes it by shade index)
    const void* d_userdata = nullptr;
    if (userdata_host && userdata_bytes > 0 && !g->g.userdata.empty()) {
        if (st.ud_bytes < (size_t)userdata_bytes) {
            if (st.d_ud) cudaFree(st.d_ud);
            st.d_ud = nullptr; st.ud_bytes = 0;
            if (cudaMalloc(&st.d_ud, (size_t)userdata_bytes) != cudaSuccess)
                return fail(B200_ERR_CUDA, "cudaMalloc(userdata) failed");
            st.ud_bytes = (size_t)userdata_bytes;
        }
        if (cudaMemcpy(st.d_ud, userdata_host, (size_t)userdata_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            return fail(B200_ERR_CUDA, "userdata upload failed");
        // userdata=record: the caller hands over the records of this batch only (record 0 = the
        // point with shade index `first`); the kernel indexes by shade index, so rebase
        d_userdata = g->userdata_record > 0 ? st.d_ud - g->userdata_record * first : st.d_ud;
    }
    auto out_bytes_of = [&](int k) -> size_t {
        const Symbol& sy = g->g.layers[g->g.outputs[k].first].m.syms[g->g.outputs[k].second];
        return (size_t)4 * sy.type.ncomp() * (sy.type.arraylen ? sy.type.arraylen : 1) * (sy.out.derivs ? 3 : 1);
    };
    cudaError_t ce = cudaSuccess;
    int slot = 0;
    for (long long b = 0; b < npoints && ce == cudaSuccess; b += chunk, slot = (slot + 1) % NS) {
        long long n     = (npoints - b) < chunk ? (npoints - b) : chunk;
        cudaStream_t s  = st.streams[slot];
        char* din       = st.d_in + (size_t)slot * chunk * in_bpp;
        char* dout      = st.d_out + (size_t)slot * chunk * out_bpp;
        b200_globals dg = *sg;
        dg.plane_stride = n;
        size_t off      = 0;
        for (const Plane& p : planes) {
            dg.varying[p.field] = (const float*)(din + off);
            for (int c = 0; c < p.comps; ++c) {
                const float* src = sg->varying[p.field] + (size_t)c * sg->plane_stride + b;
                cudaError_t e = cudaMemcpyAsync(din + off, src, (size_t)n * 4, cudaMemcpyHostToDevice, s);
                if (e != cudaSuccess) ce = e;
                off += (size_t)n * 4;
            }
        }
        for (int f = 0; f < B200_SG_NFIELDS; ++f)
            if (sg->varying[f] && !g->g.globals_read.count(f))
                dg.varying[f] = nullptr;
        // kernel sees shadeindex = b + i; each cluster's records are rebased
        // into a dense staging region:  addr = dout + S_c + (offset_k - lo_c) + stride*(sidx - b)
        long long adjust[B200_MAX_OUTPUTS] = { 0 };
        size_t region                      = 0;
        for (const Cluster& c : clusters) {
            for (int k : c.outs)
                adjust[k] = (long long)region - c.lo - c.stride * (first + b);
            region += (size_t)n * c.stride;
        }
        rc = launch_group(g, device, s, n, &dg, nullptr, d_userdata, dout, first + b, adjust);
        if (rc != B200_OK)
            return rc;
        region = 0;
        for (const Cluster& c : clusters) {
            char* hdst = (char*)output_base + c.lo + c.stride * (first + b);
            if (c.dense) {
                // the fields tile the record: one contiguous copy of n whole records
                size_t bytes  = (size_t)(n - 1) * c.stride + (size_t)(c.hi - c.lo);
                cudaError_t e = cudaMemcpyAsync(hdst, dout + region, bytes, cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess) ce = e;
            } else {
                // sparse record (the renderer owns the bytes between the fields): copy each
                // symbol's bytes only, the SymLocationDesc contract of the device path
                for (int k : c.outs) {
                    const Symbol& sy = g->g.layers[g->g.outputs[k].first].m.syms[g->g.outputs[k].second];
                    size_t foff      = (size_t)(sy.out.offset - c.lo);
                    cudaError_t e    = cudaMemcpy2DAsync(hdst + foff, (size_t)c.stride, dout + region + foff,
                                                         (size_t)c.stride, out_bytes_of(k), (size_t)n,
                                                         cudaMemcpyDeviceToHost, s);
                    if (e != cudaSuccess) ce = e;
                }
            }
            region += (size_t)n * c.stride;
        }
    }
    for (int s = 0; s < NS; ++s) {
        cudaError_t e = cudaStreamSynchronize(st.streams[s]);
        if (e != cudaSuccess) ce = e;
    }
    if (ce == cudaSuccess)
        ce = cudaGetLastError();
    if (ce != cudaSuccess)
        return fail(B200_ERR_CUDA, std::string("execute_host: ") + cudaGetErrorString(ce));
    return B200_OK;
}

int
b200_group_execute_host(b200_group* g, int device, long long npoints, const b200_globals* sg, void* output_base)
{
    return execute_host_impl(g, device, npoints, sg, nullptr, 0, output_base, 0);
}

int
b200_group_execute_host_userdata(b200_group* g, int device, long long npoints, const b200_globals* sg,
                                 const void* userdata_base, long long userdata_bytes, void* output_base)
{
    return execute_host_impl(g, device, npoints, sg, userdata_base, userdata_bytes, output_base, 0);
}

/* introspection of the userdata layout and of the named spaces (the C++ API mirror feeds them
 * from RendererServices::get_userdata / get_matrix) */
int
b200_group_userdata_fields(const b200_group* g, long long* record_bytes)
{
    if (record_bytes)
        *record_bytes = g ? g->userdata_record : 0;
    return g ? (int)g->g.userdata.size() : 0;
}
int
b200_group_userdata_field(const b200_group* g, int i, b200_userdata* out)
{
    if (!g || !out || i < 0 || i >= (int)g->g.userdata.size())
        return fail(B200_ERR_INVALID, "b200_group_userdata_field: bad index");
    const UserData& u = g->g.userdata[i];
    out->name = u.name.c_str(); out->ncomp = u.ncomp; out->is_int = u.is_int; out->offset = u.offset;
    out->stride = u.stride; out->derivs = u.derivs; out->valid_offset = u.valid_offset; out->valid_stride = u.valid_stride;
    return B200_OK;
}
int
b200_group_num_spaces(const b200_group* g)
{
    return g ? (int)g->g.spaces.size() : 0;
}
const char*
b200_group_space_name(const b200_group* g, int i)
{
    return (g && i >= 0 && i < (int)g->g.spaces.size()) ? g->g.spaces[i].c_str() : "";
}

int
b200_group_execute_host_at(b200_group* g, int device, long long npoints, const b200_globals* sg,
                           long long first_shadeindex, const void* userdata_base, long long userdata_bytes,
                           void* output_base)
{
    return execute_host_impl(g, device, npoints, sg, userdata_base, userdata_bytes, output_base, first_shadeindex);
}

}  // extern "C"

// ---------------------------------------------------------------------------
// AOT batch kernels of the device shadeop library (b200_shadeop_*)
// ---------------------------------------------------------------------------
namespace {
using namespace osld;

template<int KIND, int DIM, int NC, bool DERIV, bool PER>
__global__ void __launch_bounds__(256)
noise_kernel(long long n, const float* __restrict__ in, const float* __restrict__ period, float* __restrict__ out)
{
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x[4];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            x[d] = in[d * n + i];
        if (KIND == 2 || KIND == 3) {
            if (PER) {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    x[d] = pwrap(x[d], period[d]);
            }
            float r[3];
            ihnoise<KIND == 2, DIM, NC>(r, x);
#pragma unroll
            for (int c = 0; c < NC; ++c)
                out[c * n + i] = r[c];
            if (DERIV) {
#pragma unroll
                for (int c = 0; c < 2 * NC; ++c)
                    out[(NC + c) * n + i] = 0.0f;
            }
        } else {
            int per[4] = { 1, 1, 1, 1 };
            if (PER) {
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    per[d] = iperiod(period[d]);
            }
            if (DERIV) {
                Df xd[4], r[3];
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                    xd[d] = mkd(x[d], in[(DIM + d) * n + i], in[(2 * DIM + d) * n + i]);
                if (KIND >= 4)
                    simplex<DIM, NC, KIND == 5>(r, xd);
                else
                    perlin<Df, DIM, NC, KIND == 1, PER>(r, xd, per);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    out[c * n + i]            = r[c].val;
                    out[(NC + c) * n + i]     = r[c].dx;
                    out[(2 * NC + c) * n + i] = r[c].dy;
                }
            } else {
                float r[3];
                if (KIND >= 4)
                    simplex<DIM, NC, KIND == 5>(r, x);
                else
                    perlin<float, DIM, NC, KIND == 1, PER>(r, x, per);
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    out[c * n + i] = r[c];
            }
        }
    }
}

template<int DIM>
__global__ void __launch_bounds__(256) hash_kernel(long long n, const float* __restrict__ in, int* __restrict__ out)
{
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float x[4];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
            x[d] = in[d * n + i];
        int h = DIM == 1 ? hash_f(x[0])
                : DIM == 2 ? hash_ff(x[0], x[1])
                : DIM == 3 ? hash_v(mkv(x[0], x[1], x[2]))
                           : hash_vf(mkv(x[0], x[1], x[2]), x[3]);
        out[i] = h;
    }
}

template<int KIND, int DIM, int NC, bool DERIV>
void
launch_noise2(bool per, unsigned grid, cudaStream_t s, long long n, const float* in, const float* period, float* out)
{
    if (per)
        noise_kernel<KIND, DIM, NC, DERIV, true><<<grid, 256, 0, s>>>(n, in, period, out);
    else
        noise_kernel<KIND, DIM, NC, DERIV, false><<<grid, 256, 0, s>>>(n, in, period, out);
}
template<int KIND, int DIM>
void
launch_noise1(int nc, bool deriv, bool per, unsigned grid, cudaStream_t s, long long n, const float* in,
              const float* period, float* out)
{
    if (nc == 1 && !deriv) launch_noise2<KIND, DIM, 1, false>(per, grid, s, n, in, period, out);
    else if (nc == 1) launch_noise2<KIND, DIM, 1, true>(per, grid, s, n, in, period, out);
    else if (!deriv) launch_noise2<KIND, DIM, 3, false>(per, grid, s, n, in, period, out);
    else launch_noise2<KIND, DIM, 3, true>(per, grid, s, n, in, period, out);
}
template<int KIND>
void
launch_noise0(int dim, int nc, bool deriv, bool per, unsigned grid, cudaStream_t s, long long n, const float* in,
              const float* period, float* out)
{
    switch (dim) {
    case 1: launch_noise1<KIND, 1>(nc, deriv, per, grid, s, n, in, period, out); break;
    case 2: launch_noise1<KIND, 2>(nc, deriv, per, grid, s, n, in, period, out); break;
    case 3: launch_noise1<KIND, 3>(nc, deriv, per, grid, s, n, in, period, out); break;
    default: launch_noise1<KIND, 4>(nc, deriv, per, grid, s, n, in, period, out); break;
    }
}
unsigned
grid_for(long long n)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long want = (n + 255) / 256, cap = (long long)sms * 32;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}
}  // namespace

extern "C" int
b200_shadeop_noise(int kind, int outdim, int indim, int derivs, int fma, long long n, const float* in,
                   const float* period, float* out, void* stream)
{
    (void)fma;  // the AOT library is built strict (-fmad=false); generated groups choose per group
    if (kind < 0 || kind > 5 || (kind >= 4 && period) || (outdim != 1 && outdim != 3) || indim < 1 || indim > 4 || n < 0 || !in || !out)
        return fail(B200_ERR_INVALID, "b200_shadeop_noise: bad arguments");
    if (n == 0)
        return B200_OK;
    unsigned grid  = grid_for(n);
    cudaStream_t s = (cudaStream_t)stream;
    bool per       = period != nullptr;
    switch (kind) {
    case 0: launch_noise0<0>(indim, outdim, derivs != 0, per, grid, s, n, in, period, out); break;
    case 1: launch_noise0<1>(indim, outdim, derivs != 0, per, grid, s, n, in, period, out); break;
    case 2: launch_noise0<2>(indim, outdim, derivs != 0, per, grid, s, n, in, period, out); break;
    case 3: launch_noise0<3>(indim, outdim, derivs != 0, per, grid, s, n, in, period, out); break;
    case 4: launch_noise0<4>(indim, outdim, derivs != 0, false, grid, s, n, in, period, out); break;
    default: launch_noise0<5>(indim, outdim, derivs != 0, false, grid, s, n, in, period, out); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(B200_ERR_CUDA, std::string("b200_shadeop_noise launch: ") + cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return B200_OK;
}

extern "C" int
b200_shadeop_hash(int indim, long long n, const float* in, int* out, void* stream)
{
    if (indim < 1 || indim > 4 || n < 0 || !in || !out)
        return fail(B200_ERR_INVALID, "b200_shadeop_hash: bad arguments");
    if (n == 0)
        return B200_OK;
    unsigned grid  = grid_for(n);
    cudaStream_t s = (cudaStream_t)stream;
    switch (indim) {
    case 1: hash_kernel<1><<<grid, 256, 0, s>>>(n, in, out); break;
    case 2: hash_kernel<2><<<grid, 256, 0, s>>>(n, in, out); break;
    case 3: hash_kernel<3><<<grid, 256, 0, s>>>(n, in, out); break;
    default: hash_kernel<4><<<grid, 256, 0, s>>>(n, in, out); break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(B200_ERR_CUDA, std::string("b200_shadeop_hash launch: ") + cudaGetErrorString(e));
    g_launches.fetch_add(1);
    return B200_OK;
}
