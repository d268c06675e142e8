// osl_b200_render.cu — host side of the wavefront path tracer (product code).
//
// Plays the role of SimpleRaytracer::render (src/testrender/simpleraytracer.cpp:1424-1456)
// and OptixRaytracer::render (src/testrender/optixraytracer.cpp): upload the
// prepared scene once, JIT one module holding every material group plus the
// integrator kernels (csrc/device/osl_b200_render.cuh), then per batch of
// samples run  generate -> { intersect -> [sort by material] -> shade } until
// no path is alive -> resolve.
#include "../../include/osl_b200.h"
#include "host/osl_b200_group.h"
#include "osl_b200_jit.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>

namespace oslb200 {
int set_error(int code, const std::string& msg);
void count_launches(long long n);
}  // namespace oslb200
namespace oslb200 {
std::string
bind_module_textures(void* module, const std::vector<std::string>& names, const std::string& searchpath,
                     std::vector<void*>& allocations)
{
    return bind_module_textures_impl(module, names, searchpath, allocations);
}
}  // namespace oslb200
using namespace oslb200;

namespace {

// mirror of osld::RenderScene / RenderLaunch (device/osl_b200_render.cuh)
struct DevScene {
    int nverts, ntris, nnodes, nlightprims, nshaders, nmeshes;
    const float* verts;
    const float* normals;
    const float* uvs;
    const int* triangles;
    const int* n_triangles;
    const int* uv_triangles;
    const int* shaderids;
    const int* meshids;
    const float* mesh_surfacearea;
    const void* bvh_nodes;
    const unsigned* bvh_indices;
    const unsigned* lightprims;
    const int* shader_is_light;
    float eye[3], dir[3], up[3], fov;
    float cx[3], cy[3], invw, invh;
    int xres, yres;
    int aa, max_bounces, rr_depth, no_jitter, show_globals;
    int background_shader, background_resolution;
    float* bg_values;
    float* bg_rows;
    float* bg_cols;
    int bg_res;
    float bg_invres, bg_invjacobian;
    const void* leaf_tris;
    const float* bsdl_luts;
};
enum { PATH_QUADS = 8, SHADOW_QUADS = 4 };  // 128 B of path state + 64 B of pending shadow rays per slot
// device loop state (osld::C_*)
enum { C_LIVE = 0, C_OUT, C_SHADOW, C_SHADOW_OUT, C_NEXT, C_FETCH, C_ITER, C_FINISHED, C_WORDS = 8 + 128 };
struct DevLaunch {
    DevScene S;
    void* rec;
    void* shrec;
    int* queue_in;
    int* queue_out;
    int* queue_sh;
    int* counters;
    int* sort_keys;
    const int* shader_key;
    const int* pixmap;
    volatile int* host_state;
    int nslots, npix, s0, nsamples, total;
    float* result;
    float* accum;
    float* medium;
};
enum { MEDIUM_WORDS = 72 };  // OSLD_MEDIUM_WORDS: per-slot medium stack of modules with volume closures

const char* KERNELS[] = { "rt_camera", "rt_generate", "rt_trace", "rt_light", "rt_sort_scatter", "rt_shade",
                          "rt_swap", "rt_resolve", "rt_tail",
                          // only in modules generated for a scene with a background
                          "rt_bg_eval", "rt_bg_rows", "rt_bg_finish", "rt_bg_scale" };
enum { K_CAMERA, K_GENERATE, K_TRACE, K_LIGHT, K_SORT_SCATTER, K_SHADE, K_SWAP, K_RESOLVE, K_TAIL,
       K_BG_EVAL, K_BG_ROWS, K_BG_FINISH, K_BG_SCALE, K_N };
enum { K_FIRST_BG = K_BG_EVAL };
enum { TRACE_BLOCK = 128, SHADE_BLOCK = 128 };  // OSLD_TRACE_BLOCK / OSLD_SHADE_BLOCK

}  // namespace

struct b200_render {
    std::vector<std::unique_ptr<Group>> groups;
    std::string source;
    std::vector<char> cubin;           // empty until the module is compiled (option compile=0 defers it)
    std::mutex jit_mu;
    b200_render_scene host;  // host-pointer copy of the description
    bool fma = true, sort = true;
    long long slots_target = 0;         // option slots=N: path slots in the pool (0 = default)
    long long tail_paths   = 8192;      // option tail=N: at most N live paths -> rt_tail (0 = never)
    std::vector<std::string> textures;  // the module's texture table, in slot order
    std::string texturepath;            // option texturepath=dir[:dir...]
    RenderModuleInfo info;              // what the module was specialised to
    int bvh_stack = 64, stk_words = 3;  // OSLD_BVH_STACK, OSLD_STK_WORDS
    std::vector<int> shader_key;        // sort bucket per material
    // per-device state
    struct Dev {
        CUmodule_ mod = nullptr;
        CUfunction_ fn[K_N];
        int resident[K_N];  // CTAs per SM at the launch configuration used
        DevScene S;
        std::vector<void*> allocs;
        int sms = 148;
        const int* shader_key = nullptr;
        // path pool
        long long nslots_cap = 0;
        void* rec            = nullptr;  // nslots x 128 B path records
        void* shrec          = nullptr;  // nslots x 64 B shadow records
        int* queues          = nullptr;  // 4 x nslots
        int* sort_keys       = nullptr;
        float* medium        = nullptr;  // nslots x 288 B medium stacks (scenes with volume closures)
        int* counters        = nullptr;
        volatile int* host_state = nullptr;  // mapped pinned: {iter, live, shadow, next}
        // work set
        float* accum         = nullptr;
        int* pixmap          = nullptr;
        long long pix_cap    = 0;
        float* result        = nullptr;
        long long result_cap = 0;  // samples
    };
    std::map<int, Dev> devs;
};

static std::map<std::string, std::string>
parse_opts(const char* s)
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
ensure_compiled(b200_render* r)
{
    std::lock_guard<std::mutex> lock(r->jit_mu);
    if (!r->cubin.empty())
        return B200_OK;
    std::string err = jit_compile(r->source, "osl_b200_render.cu", r->fma, r->cubin);
    if (!err.empty()) {
        r->cubin.clear();
        return set_error(B200_ERR_COMPILE, err);
    }
    return B200_OK;
}

extern "C" int
b200_render_create(const b200_render_scene* scene, int nmaterials, const b200_group_desc* materials,
                   const char* options, b200_render** out)
{
    if (!scene || !materials || nmaterials <= 0 || !out)
        return set_error(B200_ERR_INVALID, "b200_render_create: bad arguments");
    if (nmaterials > 62)
        return set_error(B200_ERR_UNSUPPORTED, "b200_render_create: more than 62 materials not supported yet");
    if (scene->background_shader >= nmaterials)
        return set_error(B200_ERR_INVALID, "b200_render_create: background_shader is not a material index");
    *out = nullptr;
    std::unique_ptr<b200_render> R(new b200_render);
    auto opt = parse_opts(options);
    if (opt.count("fma"))
        R->fma = atoi(opt["fma"].c_str()) != 0;
    if (opt.count("sort"))
        R->sort = atoi(opt["sort"].c_str()) != 0;
    if (opt.count("slots"))
        R->slots_target = atoll(opt["slots"].c_str());
    if (opt.count("tail"))
        R->tail_paths = atoll(opt["tail"].c_str());
    R->host = *scene;
    try {
        std::vector<Group*> gs;
        for (int m = 0; m < nmaterials; ++m) {
            const b200_group_desc& d = materials[m];
            std::unique_ptr<Group> g(new Group);
            g->name = d.name ? d.name : "material";
            g->fma  = R->fma;
            for (int i = 0; i < d.nlayers; ++i) {
                const b200_layer& l = d.layers[i];
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
                        else
                            pv.svals.push_back(((const char* const*)bp.values)[k]);
                    }
                    pvs.push_back(std::move(pv));
                }
                g->add_layer(l.oso_text, l.layername, pvs);
            }
            for (int i = 0; i < d.nconnections; ++i)
                g->connect(d.connections[i].srclayer, d.connections[i].srcparam, d.connections[i].dstlayer,
                           d.connections[i].dstparam);
            g->finalize();
            gs.push_back(g.get());
            R->groups.push_back(std::move(g));
        }
        std::string body = generate_cuda_render(gs, scene->background_shader >= 0, &R->info);
        // Traversal stack: sized from the depth of THIS scene's BVH (a walk holds at most one
        // pending sibling per level plus the node in hand) instead of the reference's fixed 64
        // entries, so that the stacks of a CTA fit its shared memory; an entry packs
        // (child, nprims) into one word when every node allows it.
        int depth = 1;
        bool packable = true;
        if (scene->nnodes > 0 && scene->bvh_nodes) {
            std::vector<std::pair<unsigned, int>> st;
            st.push_back({ 0u, 1 });
            while (!st.empty()) {
                auto [node, d] = st.back();
                st.pop_back();
                if (node >= (unsigned)scene->nnodes)
                    throw std::runtime_error("b200_render_create: BVH child index out of range");
                depth = std::max(depth, d);
                unsigned child, nprims;
                memcpy(&child, scene->bvh_nodes + 8 * (size_t)node + 6, 4);
                memcpy(&nprims, scene->bvh_nodes + 8 * (size_t)node + 7, 4);
                packable &= child < (1u << 26) && nprims < 64u;
                if (!nprims) {
                    if (d > 4096)
                        throw std::runtime_error("b200_render_create: BVH is not a tree");
                    st.push_back({ child, d + 1 });
                    st.push_back({ child + 1, d + 1 });
                }
            }
        }
        R->bvh_stack = depth + 2;
        R->stk_words = packable ? 2 : 3;
        std::ostringstream pre;
        pre << "#define OSLD_BVH_STACK " << R->bvh_stack << "\n";
        if (!packable)
            pre << "#define OSLD_BVH_UNPACKED 1\n";
        // tuning knobs of the integrator kernels (defaults in device/osl_b200_render.cuh)
        if (opt.count("chunk"))
            pre << "#define OSLD_TRACE_CHUNK " << atoi(opt["chunk"].c_str()) << "\n";
        if (opt.count("refill"))
            pre << "#define OSLD_TRACE_REFILL " << atoi(opt["refill"].c_str()) << "\n";
        if (opt.count("shade_blocks"))
            pre << "#define OSLD_SHADE_MINBLOCKS " << atoi(opt["shade_blocks"].c_str()) << "\n";
        if (opt.count("inline"))
            pre << "#define OSLD_ENTRY_INLINE " << (atoi(opt["inline"].c_str()) ? "__forceinline__" : "__noinline__") << "\n";
        R->source = pre.str() + body;
        // Sort buckets: materials that emit the same set of closures (their closure-type
        // signature) are neighbours in the sorted wavefront; within a signature, by material.
        {
            std::vector<std::pair<std::string, int>> order;
            for (int m = 0; m < nmaterials; ++m) {
                std::string sig;
                for (const std::string& c : gs[m]->closure_names)
                    sig += c + ",";
                order.push_back({ sig, m });
            }
            std::sort(order.begin(), order.end());
            R->shader_key.assign(nmaterials, 0);
            for (int k = 0; k < nmaterials; ++k)
                R->shader_key[order[k].second] = 1 + k;
        }
        for (Group* gp : gs)  // module texture table = the groups' lists in order
            R->textures.insert(R->textures.end(), gp->textures.begin(), gp->textures.end());
        R->texturepath = opt.count("texturepath") ? opt["texturepath"] : std::string();
    } catch (const std::exception& e) {
        return set_error(B200_ERR_COMPILE, e.what());
    }
    // compile=0: generate the module's source only; NVRTC runs at the first render / cubin request
    if (!(opt.count("compile") && atoi(opt["compile"].c_str()) == 0)) {
        int rc = ensure_compiled(R.get());
        if (rc != B200_OK)
            return rc;
    }
    *out = R.release();
    return B200_OK;
}

extern "C" const char*
b200_render_cuda_source(const b200_render* r)
{
    return r ? r->source.c_str() : "";
}

extern "C" const void*
b200_render_cubin(const b200_render* r, long long* size)
{
    if (r && ensure_compiled(const_cast<b200_render*>(r)) != B200_OK)
        r = nullptr;
    if (size)
        *size = r ? (long long)r->cubin.size() : 0;
    return r ? r->cubin.data() : nullptr;
}

static void
free_dev(b200_render::Dev& d)
{
    for (void* p : d.allocs)
        cudaFree(p);
    d.allocs.clear();
    if (d.host_state)
        cudaFreeHost((void*)d.host_state);
    d.host_state = nullptr;
    if (d.mod && jit_driver().ok)
        jit_driver().cuModuleUnload(d.mod);
}

extern "C" void
b200_render_destroy(b200_render* r)
{
    if (!r)
        return;
    for (auto& kv : r->devs) {
        cudaSetDevice(kv.first);
        free_dev(kv.second);
    }
    delete r;
}

template<class T>
static const T*
upload(b200_render::Dev& d, const T* host, size_t n, bool& ok)
{
    if (!host || n == 0)
        return nullptr;
    void* p = nullptr;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
        ok = false;
        return nullptr;
    }
    d.allocs.push_back(p);
    if (cudaMemcpy(p, host, n * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess)
        ok = false;
    return (const T*)p;
}

// openshadinglanguage_b200/data/bsdl_luts.bin, found relative to this shared library
static std::string
load_bsdl_luts(std::vector<float>& out)
{
    Dl_info info;
    if (!dladdr((const void*)&load_bsdl_luts, &info) || !info.dli_fname)
        return "cannot locate libosl_b200.so to find data/bsdl_luts.bin";
    std::string path = info.dli_fname;
    size_t slash     = path.rfind('/');
    const std::string dir = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/data/";
    // energy tables of the MaterialX microfacet closures (tools/bake_bsdl_luts.cpp), then the
    // Zeltner-Burley sheen LTC coefficients (tools/bake_zeltner_ltc.py), then the spi::Thinlayer
    // energy table of the thinlayer closure (same baker): one block, in this order
    const struct { const char* file; size_t words; } parts[] = { { "bsdl_luts.bin", 256 + 3 * 8192 },
                                                                  { "zeltner_ltc.bin", 32 * 32 * 3 },
                                                                  { "thinlayer_lut.bin", 32 * 16 * 16 } };
    out.clear();
    for (const auto& part : parts) {
        const std::string file = dir + part.file;
        FILE* f                = fopen(file.c_str(), "rb");
        if (!f)
            return "cannot open " + file + " (tables of the MaterialX closures)";
        const size_t at = out.size();
        out.resize(at + part.words);
        size_t n = fread(out.data() + at, sizeof(float), part.words, f);
        fclose(f);
        if (n != part.words)
            return file + " is truncated";
    }
    return "";
}

// dynamic shared memory (bytes) and CTA size of kernel k; *block = 0 for kernels without any
static size_t
kernel_smem(const b200_render* r, int k, unsigned* block)
{
    const size_t stack = (size_t)r->stk_words * r->bvh_stack * 4;              // per thread
    const size_t pool  = r->info.pool_in_smem ? (size_t)(r->info.pool_words + 4) * 4 : 0;  // per thread (OSLD_POOL_STORE)
    switch (k) {
    case K_TRACE: *block = TRACE_BLOCK; return stack * TRACE_BLOCK;
    case K_SHADE:
    case K_LIGHT: *block = SHADE_BLOCK; return pool * SHADE_BLOCK;
    case K_BG_EVAL: *block = 128; return pool * 128;
    case K_TAIL: *block = 32; return (stack + pool) * 32;
    default: *block = 0; return 0;
    }
}

static int
ensure_device(b200_render* r, int device, b200_render::Dev** out)
{
    auto it = r->devs.find(device);
    if (it != r->devs.end()) {
        *out = &it->second;
        return B200_OK;
    }
    Driver& drv = jit_driver();
    if (!drv.ok)
        return set_error(B200_ERR_CUDA, "CUDA driver unavailable: " + drv.why);
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess)
        return set_error(B200_ERR_CUDA, "cudaSetDevice failed");
    {
        int rc = ensure_compiled(r);
        if (rc != B200_OK)
            return rc;
    }
    b200_render::Dev d;
    CUresult_ cr = drv.cuModuleLoadData(&d.mod, r->cubin.data());
    if (cr != 0)
        return set_error(B200_ERR_CUDA, "cuModuleLoadData(render): " + drv.err(cr));
    const bool has_bg = r->host.background_shader >= 0;
    for (int k = 0; k < (has_bg ? (int)K_N : (int)K_FIRST_BG); ++k) {
        cr = drv.cuModuleGetFunction(&d.fn[k], d.mod, KERNELS[k]);
        if (cr != 0)
            return set_error(B200_ERR_CUDA, std::string("cuModuleGetFunction(") + KERNELS[k] + "): " + drv.err(cr));
    }
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, device);
    // dynamic shared memory of the kernels that stage traversal stacks / closure arenas there
    for (int k = 0; k < (has_bg ? (int)K_N : (int)K_FIRST_BG); ++k) {
        d.resident[k] = 8;
        unsigned block = 0;
        size_t smem    = kernel_smem(r, k, &block);
        if (!block)
            continue;
        if (smem > 48 * 1024) {
            cr = drv.cuFuncSetAttribute(d.fn[k], 8 /* MAX_DYNAMIC_SHARED_SIZE_BYTES */, (int)smem);
            if (cr != 0)
                return set_error(B200_ERR_CUDA, std::string("cuFuncSetAttribute(") + KERNELS[k] + ", "
                                                    + std::to_string(smem) + " B of shared memory): " + drv.err(cr));
        }
        int nb = 0;
        if (drv.cuOccupancyMaxActiveBlocksPerMultiprocessor(&nb, d.fn[k], (int)block, smem) == 0 && nb > 0)
            d.resident[k] = nb;
    }
    std::string terr = bind_module_textures(d.mod, r->textures, r->texturepath, d.allocs);
    if (!terr.empty())
        return set_error(B200_ERR_INVALID, terr);
    const b200_render_scene& h = r->host;
    bool ok                    = true;
    DevScene& S                = d.S;
    memset(&S, 0, sizeof S);
    S.nverts = h.nverts; S.ntris = h.ntris; S.nnodes = h.nnodes; S.nlightprims = h.nlightprims;
    S.nshaders = h.nshaders; S.nmeshes = h.nmeshes;
    // vertex-like arrays are indexed up to the largest referenced index: the caller passes exact sizes
    S.verts   = upload(d, h.verts, (size_t)3 * h.nverts, ok);
    // normals / uvs: sizes are not in the description; derive from the index arrays
    int maxn = -1, maxuv = -1;
    for (int i = 0; i < 3 * h.ntris; ++i) {
        if (h.n_triangles[i] > maxn) maxn = h.n_triangles[i];
        if (h.uv_triangles[i] > maxuv) maxuv = h.uv_triangles[i];
    }
    S.normals          = upload(d, h.normals, (size_t)3 * (maxn + 1), ok);
    S.uvs              = upload(d, h.uvs, (size_t)2 * (maxuv + 1), ok);
    S.triangles        = upload(d, h.triangles, (size_t)3 * h.ntris, ok);
    S.n_triangles      = upload(d, h.n_triangles, (size_t)3 * h.ntris, ok);
    S.uv_triangles     = upload(d, h.uv_triangles, (size_t)3 * h.ntris, ok);
    S.shaderids        = upload(d, h.shaderids, (size_t)h.ntris, ok);
    S.meshids          = upload(d, h.meshids, (size_t)h.ntris, ok);
    S.mesh_surfacearea = upload(d, h.mesh_surfacearea, (size_t)h.nmeshes, ok);
    S.bvh_nodes        = upload(d, h.bvh_nodes, (size_t)8 * h.nnodes, ok);
    S.bvh_indices      = upload(d, h.bvh_indices, (size_t)h.ntris, ok);
    S.lightprims       = upload(d, h.lightprims, (size_t)h.nlightprims, ok);
    S.shader_is_light  = upload(d, h.shader_is_light, (size_t)h.nshaders, ok);
    d.shader_key       = upload(d, r->shader_key.data(), r->shader_key.size(), ok);
    {
        int* hs = nullptr;
        if (cudaHostAlloc((void**)&hs, 4 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
            ok = false;
        d.host_state = hs;
    }
    {
        // triangles in BVH leaf order as 3 x float4 (vertex a with the primitive id in w, b, c)
        std::vector<float> lt((size_t)12 * h.ntris, 0.0f);
        for (int p = 0; p < h.ntris; ++p) {
            unsigned id = h.bvh_indices[p];
            for (int v = 0; v < 3; ++v) {
                int vi = h.triangles[3 * id + v];
                memcpy(&lt[(size_t)12 * p + 4 * v], h.verts + 3 * (size_t)vi, 12);
            }
            memcpy(&lt[(size_t)12 * p + 3], &id, 4);
        }
        S.leaf_tris = upload(d, lt.data(), lt.size(), ok);
    }
    if (r->info.uses_luts) {
        // energy-compensation tables of the MaterialX microfacet closures: product data next to the library
        std::vector<float> luts;
        std::string lerr = load_bsdl_luts(luts);
        if (!lerr.empty()) {
            free_dev(d);
            return set_error(B200_ERR_INVALID, lerr);
        }
        S.bsdl_luts = upload(d, luts.data(), luts.size(), ok);
    }
    if (!ok) {
        free_dev(d);
        return set_error(B200_ERR_CUDA, "scene upload failed");
    }
    memcpy(S.eye, h.eye, sizeof S.eye);
    memcpy(S.dir, h.dir, sizeof S.dir);
    memcpy(S.up, h.up, sizeof S.up);
    S.fov = h.fov;
    S.xres = h.xres; S.yres = h.yres;
    S.invw = 1.0f / h.xres;
    S.invh = 1.0f / h.yres;
    S.aa = h.aa < 1 ? 1 : h.aa;
    S.max_bounces = h.max_bounces; S.rr_depth = h.rr_depth; S.no_jitter = h.no_jitter;
    S.show_globals = h.show_globals;
    S.background_shader = h.background_shader; S.background_resolution = h.background_resolution;
    // camera: evaluate Camera::finalize on the device
    float* dcam = nullptr;
    if (cudaMalloc(&dcam, 9 * sizeof(float)) != cudaSuccess) {
        free_dev(d);
        return set_error(B200_ERR_CUDA, "cudaMalloc failed");
    }
    DevLaunch L;
    memset(&L, 0, sizeof L);
    L.S          = S;
    void* args[] = { &L, &dcam };
    cr           = drv.cuLaunchKernel(d.fn[K_CAMERA], 1, 1, 1, 32, 1, 1, 0, nullptr, args, nullptr);
    float cam[9];
    cudaError_t ce = cudaMemcpy(cam, dcam, sizeof cam, cudaMemcpyDeviceToHost);
    cudaFree(dcam);
    if (cr != 0 || ce != cudaSuccess) {
        free_dev(d);
        return set_error(B200_ERR_CUDA, "camera setup kernel failed: " + (cr ? drv.err(cr) : std::string(cudaGetErrorString(ce))));
    }
    memcpy(S.dir, cam, 12);
    memcpy(S.cx, cam + 3, 12);
    memcpy(S.cy, cam + 6, 12);
    count_launches(1);
    // background importance table (SimpleRaytracer::prepare_render, simpleraytracer.cpp:1232-1249):
    // the background shader runs on the device at every texel, the CDFs are built there too
    if (has_bg && h.background_resolution > 0) {
        int res = h.background_resolution < 32 ? 32 : h.background_resolution;
        size_t n = (size_t)res * res;
        bool okb = cudaMalloc(&S.bg_values, 3 * n * sizeof(float)) == cudaSuccess
                   && cudaMalloc(&S.bg_cols, n * sizeof(float)) == cudaSuccess
                   && cudaMalloc(&S.bg_rows, res * sizeof(float)) == cudaSuccess;
        if (S.bg_values) d.allocs.push_back(S.bg_values);
        if (S.bg_cols) d.allocs.push_back(S.bg_cols);
        if (S.bg_rows) d.allocs.push_back(S.bg_rows);
        if (!okb) {
            free_dev(d);
            return set_error(B200_ERR_CUDA, "cudaMalloc(background table) failed");
        }
        S.bg_res         = res;
        S.bg_invres      = 1.0f / res;
        S.bg_invjacobian = res * res / float(4 * M_PI);
        L.S              = S;
        void* bargs[]    = { &L };
        const int ks[4]  = { K_BG_EVAL, K_BG_ROWS, K_BG_FINISH, K_BG_SCALE };
        for (int i = 0; i < 4 && cr == 0; ++i) {
            int block = (ks[i] == K_BG_SCALE) ? 256 : (ks[i] == K_BG_FINISH ? 32 : 128);
            long long work = ks[i] == K_BG_ROWS ? res : (ks[i] == K_BG_FINISH ? 1 : (long long)n);
            int grid = (int)std::min<long long>((work + block - 1) / block, (long long)d.sms * 16);
            unsigned kb = 0;
            size_t smem = kernel_smem(r, ks[i], &kb);
            cr = drv.cuLaunchKernel(d.fn[ks[i]], grid < 1 ? 1 : grid, 1, 1, block, 1, 1, (unsigned)smem, nullptr, bargs,
                                    nullptr);
            count_launches(1);
        }
        if (cr != 0 || cudaDeviceSynchronize() != cudaSuccess) {
            free_dev(d);
            return set_error(B200_ERR_CUDA, "background table kernels failed: "
                                                + (cr ? drv.err(cr) : std::string(cudaGetErrorString(cudaGetLastError()))));
        }
    }
    r->devs[device] = d;
    *out            = &r->devs[device];
    return B200_OK;
}

template<class T>
static bool
grow(b200_render::Dev& d, T*& p, long long& cap, long long want, size_t elem)
{
    if (cap >= want && p)
        return true;
    if (p) {
        cudaFree((void*)p);
        for (auto& a : d.allocs)
            if (a == (void*)p) a = nullptr;
        p = nullptr;
    }
    void* q = nullptr;
    if (cudaMalloc(&q, (size_t)want * elem) != cudaSuccess) {
        cap = 0;
        return false;
    }
    d.allocs.push_back(q);
    p   = (T*)q;
    cap = want;
    return true;
}

// The work set of one call: ntiles rectangles (x0, y0, w, h) of the image.  Pixels are numbered
// tile after tile, row-major inside a tile; out_rgb receives 3 floats per pixel in that order.
extern "C" int
b200_render_tiles(b200_render* r, int device, int ntiles, const int* tiles, void* out_rgb, int out_on_device,
                  b200_render_stats* stats)
{
    if (!r || !out_rgb || ntiles <= 0 || !tiles)
        return set_error(B200_ERR_INVALID, "b200_render_tiles: bad arguments");
    const int xres = r->host.xres, yres = r->host.yres;
    if (xres > 65535 || yres > 65535)
        return set_error(B200_ERR_UNSUPPORTED, "b200_render_tiles: images beyond 65535 pixels a side are not supported");
    long long npix = 0;
    for (int t = 0; t < ntiles; ++t) {
        const int *q = tiles + 4 * t;
        if (q[0] < 0 || q[1] < 0 || q[2] <= 0 || q[3] <= 0 || q[0] + q[2] > xres || q[1] + q[3] > yres)
            return set_error(B200_ERR_INVALID, "b200_render_tiles: tile outside the image");
        npix += (long long)q[2] * q[3];
    }
    b200_render::Dev* dp = nullptr;
    int rc               = ensure_device(r, device, &dp);
    if (rc != B200_OK)
        return rc;
    b200_render::Dev& d = *dp;
    Driver& drv         = jit_driver();
    cudaSetDevice(device);
    const int nsamp = d.S.aa * d.S.aa;
    // ---- rounds: as many whole sample planes as fit the per-sample result buffer ----------------
    // Every sample of a round owns a 12-byte result slot (rt_resolve folds them in sample order
    // afterwards); sample ids are 32-bit.  One round covers the whole call unless the work set
    // times spp is beyond 2^30 samples or a third of the free device memory.
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    free_b += (size_t)d.result_cap * 12;   // our own buffer is reusable
    long long max_samples = std::min<long long>(1LL << 30, (long long)(free_b / 3 / 12));
    long long SB          = std::min<long long>(nsamp, max_samples / npix);
    if (SB < 1)
        return set_error(B200_ERR_UNSUPPORTED, "b200_render_tiles: work set too large for one call; pass fewer tiles");
    // ---- the path pool ----------------------------------------------------------------------------
    // With regeneration the pool only has to keep the machine busy, not hold a frame: the
    // default is 2 Mi slots (13 waves of the 148 x 2048 resident threads; 400 MB of state).
    long long nslots = r->slots_target > 0 ? r->slots_target : (2LL << 20);
    nslots           = std::min<long long>(nslots, SB * npix);
    nslots           = std::max<long long>(nslots, 1);
    if (d.nslots_cap < nslots) {
        long long c1 = d.nslots_cap, c2 = d.nslots_cap, c3 = d.nslots_cap, c4 = d.nslots_cap;
        bool ok = grow(d, d.rec, c1, nslots, (size_t)16 * PATH_QUADS) && grow(d, d.shrec, c2, nslots, (size_t)16 * SHADOW_QUADS)
                  && grow(d, d.queues, c3, nslots, sizeof(int) * 4) && grow(d, d.sort_keys, c4, nslots, sizeof(int));
        if (ok && r->info.uses_media) {
            long long c5 = d.nslots_cap;
            ok           = grow(d, d.medium, c5, nslots, sizeof(float) * MEDIUM_WORDS);
        }
        if (!ok) {
            d.nslots_cap = 0;
            return set_error(B200_ERR_CUDA, "cudaMalloc(path pool) failed");
        }
        d.nslots_cap = nslots;
    }
    if (!d.counters) {
        if (cudaMalloc(&d.counters, sizeof(int) * C_WORDS) != cudaSuccess)
            return set_error(B200_ERR_CUDA, "cudaMalloc failed");
        d.allocs.push_back(d.counters);
    }
    {
        long long c1 = d.pix_cap, c2 = d.pix_cap;
        if (!grow(d, d.accum, c1, npix, sizeof(float) * 3) || !grow(d, d.pixmap, c2, npix, sizeof(int))) {
            d.pix_cap = 0;
            return set_error(B200_ERR_CUDA, "cudaMalloc(work set) failed");
        }
        d.pix_cap = std::min(c1, c2);
        if (!grow(d, d.result, d.result_cap, SB * npix, sizeof(float) * 3))
            return set_error(B200_ERR_CUDA, "cudaMalloc(sample results) failed");
    }
    {
        std::vector<int> pm((size_t)npix);
        size_t k = 0;
        for (int t = 0; t < ntiles; ++t) {
            const int* q = tiles + 4 * t;
            for (int y = q[1]; y < q[1] + q[3]; ++y)
                for (int x = q[0]; x < q[0] + q[2]; ++x)
                    pm[k++] = (int)(((unsigned)y << 16) | (unsigned)x);
        }
        if (cudaMemcpy(d.pixmap, pm.data(), sizeof(int) * (size_t)npix, cudaMemcpyHostToDevice) != cudaSuccess)
            return set_error(B200_ERR_CUDA, "work set upload failed");
    }
    DevLaunch L;
    memset(&L, 0, sizeof L);
    L.S = d.S;
    L.rec = d.rec;
    L.shrec = d.shrec;
    int* qbuf[3]  = { d.queues, d.queues + nslots, d.queues + 2 * nslots };
    L.queue_sh    = d.queues + 3 * nslots;
    L.counters    = d.counters;
    const bool sorting = r->sort && d.S.nshaders > 1;
    L.sort_keys   = sorting ? d.sort_keys : nullptr;
    L.shader_key  = d.shader_key;
    L.pixmap      = d.pixmap;
    L.host_state  = d.host_state;
    L.nslots      = (int)nslots;
    L.npix        = (int)npix;
    L.result      = d.result;
    L.accum       = out_on_device ? (float*)out_rgb : d.accum;
    L.medium      = d.medium;
    long long launches = 0, iters = 0;
    cudaEvent_t e0, e1, t0, t1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    cudaEventRecord(e0, 0);
    float tail_ms = 0;
    auto launch = [&](int k, long long work, bool persistent) -> int {
        unsigned block = 0;
        size_t smem    = kernel_smem(r, k, &block);
        if (!block)
            block = (k == K_SWAP) ? 128 : 256;
        long long want = (work + block - 1) / block;
        long long cap  = (long long)d.sms * (persistent ? d.resident[k] : 16);
        if (k == K_TAIL)
            cap = want;
        unsigned grid  = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        void* args[]   = { &L };
        CUresult_ cr   = drv.cuLaunchKernel(d.fn[k], grid, 1, 1, block, 1, 1, (unsigned)smem, nullptr, args, nullptr);
        ++launches;
        if (cr != 0)
            return set_error(B200_ERR_CUDA, std::string("cuLaunchKernel(") + KERNELS[k] + "): " + drv.err(cr));
        return B200_OK;
    };
    auto fail = [&](const char* what) {
        cudaError_t ce = cudaGetLastError();
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(t0); cudaEventDestroy(t1);
        return set_error(B200_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(ce));
    };
    volatile int* hs = d.host_state;
    for (int s0 = 0; s0 < nsamp; s0 += (int)SB) {
        const int nb = (int)std::min<long long>(nsamp - s0, SB);
        L.s0       = s0;
        L.nsamples = nb;
        L.total    = (int)(nb * npix);
        const long long fill = std::min<long long>(nslots, L.total);
        hs[0] = hs[1] = hs[2] = hs[3] = 0;
        if (cudaMemsetAsync(d.counters, 0, sizeof(int) * C_WORDS, 0) != cudaSuccess)
            return fail("counter reset");
        int cur    = 0;  // qbuf index of the live queue
        L.queue_in = qbuf[cur];
        if ((rc = launch(K_GENERATE, fill, false)) != B200_OK)
            return rc;
        // One step = trace, light, [scatter], shade, swap.  The kernels take their counts from
        // device memory; the host only needs upper bounds for the grids and the moment to stop.
        // rt_swap publishes {step, live, shadow, next} in mapped pinned memory, so the host
        // keeps up to RUNAHEAD steps enqueued and never synchronises inside the loop.
        const int RUNAHEAD = 12;
        long long enq = 0;
        bool tail = false;
        for (;;) {
            const int it = hs[0];
            const long long live = hs[1], shadow = hs[2], next = hs[3];
            const bool exhausted = it > 0 && next >= L.total;
            if (it > 0 && live == 0 && shadow == 0)
                break;
            if (exhausted && r->tail_paths > 0 && live > 0 && live <= r->tail_paths) {
                tail = true;
                break;
            }
            if (enq - it >= RUNAHEAD) {
                cudaError_t q = cudaStreamQuery(0);
                if (q != cudaSuccess && q != cudaErrorNotReady)
                    return fail("render step failed");
                continue;
            }
            const long long bound = exhausted ? std::max(live, shadow) : fill;
            ++iters;
            ++enq;
            L.queue_in  = qbuf[cur];
            L.queue_out = qbuf[(cur + 1) % 3];   // next live queue: filled by light (regen) and shade
            if ((rc = launch(K_TRACE, 2 * bound, true)) != B200_OK) return rc;
            if ((rc = launch(K_LIGHT, bound, false)) != B200_OK) return rc;
            if (sorting) {
                L.queue_out = qbuf[(cur + 2) % 3];   // the sorted copy of the live queue
                if ((rc = launch(K_SORT_SCATTER, bound, false)) != B200_OK) return rc;
                L.queue_in  = qbuf[(cur + 2) % 3];
                L.queue_out = qbuf[(cur + 1) % 3];
            }
            if ((rc = launch(K_SHADE, bound, false)) != B200_OK) return rc;
            if ((rc = launch(K_SWAP, 1, false)) != B200_OK) return rc;
            cur = (cur + 1) % 3;
        }
        if (tail) {
            // the stragglers: finish the pending shadow rays, find the closest hits, then one
            // launch runs every remaining path to its end (rt_tail)
            L.queue_in  = qbuf[cur];
            L.queue_out = qbuf[(cur + 1) % 3];
            const long long bound = std::max<long long>(hs[1], hs[2]);
            if ((rc = launch(K_TRACE, 2 * std::max<long long>(bound, r->tail_paths), true)) != B200_OK) return rc;
            if ((rc = launch(K_LIGHT, std::max<long long>(bound, r->tail_paths), false)) != B200_OK) return rc;
            cudaEventRecord(t0, 0);
            if ((rc = launch(K_TAIL, r->tail_paths * 32, false)) != B200_OK) return rc;   // one CTA (one warp) per path
            cudaEventRecord(t1, 0);
        }
        if (cudaStreamSynchronize(0) != cudaSuccess)
            return fail("render round failed");
        if (tail) {
            float ms = 0;
            cudaEventElapsedTime(&ms, t0, t1);
            tail_ms += ms;
        }
        if ((rc = launch(K_RESOLVE, npix, false)) != B200_OK)
            return rc;
    }
    cudaEventRecord(e1, 0);
    cudaError_t ce = cudaSuccess;
    if (!out_on_device)
        ce = cudaMemcpy(out_rgb, d.accum, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost);
    else
        ce = cudaStreamSynchronize(0);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(t0); cudaEventDestroy(t1);
    count_launches(launches);
    if (ce != cudaSuccess)
        return set_error(B200_ERR_CUDA, std::string("render readback: ") + cudaGetErrorString(ce));
    if (stats) {
        stats->paths             = npix * nsamp;
        stats->launches          = launches;
        stats->bounce_iterations = iters;
        stats->device_ms         = ms;
        stats->tail_ms           = tail_ms;
        stats->slots             = nslots;
        stats->rounds            = (nsamp + SB - 1) / SB;
    }
    return B200_OK;
}

extern "C" int
b200_render_rows(b200_render* r, int device, int y0, int y1, float* host_rgb, b200_render_stats* stats)
{
    if (!r || !host_rgb || y0 < 0 || y1 > r->host.yres || y0 >= y1)
        return set_error(B200_ERR_INVALID, "b200_render_rows: bad arguments");
    const int tile[4] = { 0, y0, r->host.xres, y1 - y0 };
    return b200_render_tiles(r, device, 1, tile, host_rgb, 0, stats);
}
