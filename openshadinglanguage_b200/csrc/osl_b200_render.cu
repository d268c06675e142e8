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

#include <cstring>
#include <map>
#include <memory>
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
};
enum { PATH_QUADS = 8 };  // 8 x float4 = 128 B of state per path (OSLD_PATH_QUADS)
struct DevLaunch {
    DevScene S;
    void* rec;
    int* queue_in;
    int* queue_out;
    int* counters;
    int* sort_keys;
    int nslots, npix, y0, s0, nsamples;
    float* accum;
};

const char* KERNELS[] = { "rt_camera", "rt_generate", "rt_intersect", "rt_sort_count", "rt_sort_scan",
                          "rt_sort_scatter", "rt_shade", "rt_swap", "rt_resolve", "rt_tail",
                          // only in modules generated for a scene with a background
                          "rt_bg_eval", "rt_bg_rows", "rt_bg_finish", "rt_bg_scale" };
enum { K_CAMERA, K_GENERATE, K_INTERSECT, K_SORT_COUNT, K_SORT_SCAN, K_SORT_SCATTER, K_SHADE, K_SWAP, K_RESOLVE,
       K_TAIL, K_BG_EVAL, K_BG_ROWS, K_BG_FINISH, K_BG_SCALE, K_N };
enum { K_FIRST_BG = K_BG_EVAL };

}  // namespace

struct b200_render {
    std::vector<std::unique_ptr<Group>> groups;
    std::string source;
    std::vector<char> cubin;
    b200_render_scene host;  // host-pointer copy of the description
    bool fma = true, sort = true;
    long long slots_target = 0;         // option slots=N; 0 = sized from the free device memory
    long long tail_paths   = 2048;      // option tail=N: at most N live paths -> rt_tail (0 = never)
    std::vector<std::string> textures;  // the module's texture table, in slot order
    std::string texturepath;            // option texturepath=dir[:dir...]
    // per-device state
    struct Dev {
        CUmodule_ mod = nullptr;
        CUfunction_ fn[K_N];
        DevScene S;
        std::vector<void*> allocs;
        int sms = 148;
        // path state
        long long nslots_cap = 0;
        void* rec            = nullptr;  // nslots x 128 B path records
        int* queues          = nullptr;  // 3 x nslots
        int* sort_keys       = nullptr;
        int* counters        = nullptr;
        float* accum         = nullptr;
        long long accum_cap  = 0;
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
        R->source = generate_cuda_render(gs, scene->background_shader >= 0);
        for (Group* gp : gs)  // module texture table = the groups' lists in order
            R->textures.insert(R->textures.end(), gp->textures.begin(), gp->textures.end());
        R->texturepath = opt.count("texturepath") ? opt["texturepath"] : std::string();
    } catch (const std::exception& e) {
        return set_error(B200_ERR_COMPILE, e.what());
    }
    std::string err = jit_compile(R->source, "osl_b200_render.cu", R->fma, R->cubin);
    if (!err.empty())
        return set_error(B200_ERR_COMPILE, err);
    *out = R.release();
    return B200_OK;
}

extern "C" const char*
b200_render_cuda_source(const b200_render* r)
{
    return r ? r->source.c_str() : "";
}

static void
free_dev(b200_render::Dev& d)
{
    for (void* p : d.allocs)
        cudaFree(p);
    d.allocs.clear();
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
            cr       = drv.cuLaunchKernel(d.fn[ks[i]], grid < 1 ? 1 : grid, 1, 1, block, 1, 1, 0, nullptr, bargs, nullptr);
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

extern "C" int
b200_render_rows(b200_render* r, int device, int y0, int y1, float* host_rgb, b200_render_stats* stats)
{
    if (!r || !host_rgb || y0 < 0 || y1 > r->host.yres || y0 >= y1)
        return set_error(B200_ERR_INVALID, "b200_render_rows: bad arguments");
    b200_render::Dev* dp = nullptr;
    int rc               = ensure_device(r, device, &dp);
    if (rc != B200_OK)
        return rc;
    b200_render::Dev& d = *dp;
    Driver& drv         = jit_driver();
    cudaSetDevice(device);
    const int xres      = r->host.xres;
    const long long npix = (long long)(y1 - y0) * xres;
    const int nsamp     = d.S.aa * d.S.aa;
    // Path slots per batch.  Every batch ends in a tail of a few long paths during which the
    // GPU is nearly idle, so batches are made as large as memory allows (HBM is there to be
    // used): by default a quarter of the free device memory at ~148 B of state per slot
    // (render-microfacet 2048^2 x 64 spp: 37 s with 32 Mi slots, 152 s with 4 Mi).
    long long target = r->slots_target;
    if (target <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        target = (long long)(free_b / 4 / 148);
        if (target < (4 << 20)) target = 4 << 20;
    }
    const long long max_slots = 0x7fffffffLL / 4;
    if (target > max_slots) target = max_slots;
    long long SB = target / npix;
    if (SB < 1) SB = 1;
    if (SB > nsamp) SB = nsamp;
    const long long nslots = SB * npix;
    if (nslots > 0x7fffffffLL / 4)
        return set_error(B200_ERR_UNSUPPORTED, "b200_render_rows: band too large; render fewer rows per call");
    // (re)allocate path state
    if (d.nslots_cap < nslots) {
        auto drop = [&](void* p) {
            if (!p) return;
            cudaFree(p);
            for (auto& a : d.allocs)
                if (a == p) a = nullptr;
        };
        drop(d.rec); drop(d.queues); drop(d.sort_keys);
        bool ok = cudaMalloc(&d.rec, (size_t)16 * PATH_QUADS * nslots) == cudaSuccess
                  && cudaMalloc(&d.queues, sizeof(int) * 3 * nslots) == cudaSuccess
                  && cudaMalloc(&d.sort_keys, sizeof(int) * nslots) == cudaSuccess;
        if (!ok)
            return set_error(B200_ERR_CUDA, "cudaMalloc(path state) failed");
        d.allocs.push_back(d.rec); d.allocs.push_back(d.queues); d.allocs.push_back(d.sort_keys);
        d.nslots_cap = nslots;
    }
    if (!d.counters) {
        if (cudaMalloc(&d.counters, sizeof(int) * 256) != cudaSuccess)
            return set_error(B200_ERR_CUDA, "cudaMalloc failed");
        d.allocs.push_back(d.counters);
    }
    if (d.accum_cap < npix) {
        if (d.accum) {
            cudaFree(d.accum);
            for (auto& a : d.allocs)
                if (a == d.accum) a = nullptr;
        }
        if (cudaMalloc(&d.accum, sizeof(float) * 3 * npix) != cudaSuccess)
            return set_error(B200_ERR_CUDA, "cudaMalloc failed");
        d.allocs.push_back(d.accum);
        d.accum_cap = npix;
    }
    cudaMemset(d.counters, 0, sizeof(int) * 256);
    DevLaunch L;
    memset(&L, 0, sizeof L);
    L.S = d.S;
    L.rec = d.rec;
    int* qbuf[3]  = { d.queues, d.queues + nslots, d.queues + 2 * nslots };
    L.counters    = d.counters;
    L.sort_keys   = r->sort ? d.sort_keys : nullptr;
    L.npix        = (int)npix;
    L.y0          = y0;
    L.accum       = d.accum;
    long long launches = 0, iters = 0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, 0);
    auto launch = [&](int k, long long work, unsigned block) -> int {
        long long want = (work + block - 1) / block;
        long long cap  = (long long)d.sms * 16;
        unsigned grid  = (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
        void* args[]   = { &L };
        CUresult_ cr   = drv.cuLaunchKernel(d.fn[k], grid, 1, 1, block, 1, 1, 0, nullptr, args, nullptr);
        ++launches;
        if (cr != 0)
            return set_error(B200_ERR_CUDA, std::string("cuLaunchKernel(") + KERNELS[k] + "): " + drv.err(cr));
        return B200_OK;
    };
    int* hcount = nullptr;
    cudaMallocHost(&hcount, sizeof(int));
    for (int s0 = 0; s0 < nsamp; s0 += (int)SB) {
        int nb     = (int)((nsamp - s0) < SB ? (nsamp - s0) : SB);
        L.s0       = s0;
        L.nsamples = nb;
        L.nslots   = (int)(nb * npix);
        int cur    = 0;  // qbuf index of the live queue
        L.queue_in = qbuf[cur];
        if ((rc = launch(K_GENERATE, L.nslots, 256)) != B200_OK)
            return rc;
        long long live = L.nslots;
        // The kernels take the live count from device memory; the host only needs it to size
        // grids (an upper bound is enough: the count never grows) and to stop.  So several
        // bounces are enqueued back to back and the count is read once per chunk: no host
        // round trip per bounce, and the launches of a chunk overlap the execution of the
        // previous kernels.  Empty trailing bounces cost a few microseconds each.
        int since_sync = 0, chunk = 1;
        while (live > 0) {
            ++iters;
            L.queue_in  = qbuf[cur];
            if (live <= r->tail_paths) {
                // the stragglers: one launch runs each of them to its end (rt_tail)
                if ((rc = launch(K_TAIL, live * 32, 32)) != B200_OK)   // one CTA (one warp) per path
                    return rc;
                if ((rc = launch(K_SWAP, 1, 32)) != B200_OK)   // counters[1] is 0: nothing is queued
                    return rc;
                if (cudaStreamSynchronize(0) != cudaSuccess) {
                    cudaError_t ce = cudaGetLastError();
                    return set_error(B200_ERR_CUDA, std::string("render tail failed: ") + cudaGetErrorString(ce));
                }
                break;
            }
            L.queue_out = qbuf[(cur + 1) % 3];
            if ((rc = launch(K_INTERSECT, live, 256)) != B200_OK)
                return rc;
            if (r->sort && d.S.nshaders > 1) {
                if ((rc = launch(K_SORT_COUNT, live, 256)) != B200_OK) return rc;
                if ((rc = launch(K_SORT_SCAN, 1, 32)) != B200_OK) return rc;
                if ((rc = launch(K_SORT_SCATTER, live, 256)) != B200_OK) return rc;
                cur        = (cur + 1) % 3;  // sorted queue
                L.queue_in = qbuf[cur];
            }
            L.queue_out = qbuf[(cur + 1) % 3];
            if ((rc = launch(K_SHADE, live, 128)) != B200_OK)
                return rc;
            if ((rc = launch(K_SWAP, 1, 32)) != B200_OK)
                return rc;
            cur = (cur + 1) % 3;
            if (++since_sync < chunk)
                continue;
            since_sync = 0;
            chunk      = chunk < 8 ? chunk * 2 : 8;   // 1, 2, 4, 8, 8, ... bounces per read-back
            cudaMemcpyAsync(hcount, d.counters, sizeof(int), cudaMemcpyDeviceToHost, 0);
            if (cudaStreamSynchronize(0) != cudaSuccess) {
                cudaError_t ce = cudaGetLastError();
                return set_error(B200_ERR_CUDA, std::string("render bounce failed: ") + cudaGetErrorString(ce));
            }
            live = *hcount;
        }
        if ((rc = launch(K_RESOLVE, npix, 256)) != B200_OK)
            return rc;
    }
    cudaEventRecord(e1, 0);
    cudaError_t ce = cudaMemcpy(host_rgb, d.accum, sizeof(float) * 3 * npix, cudaMemcpyDeviceToHost);
    cudaFreeHost(hcount);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    count_launches(launches);
    if (ce != cudaSuccess)
        return set_error(B200_ERR_CUDA, std::string("render readback: ") + cudaGetErrorString(ce));
    if (stats) {
        stats->paths             = npix * nsamp;
        stats->launches          = launches;
        stats->bounce_iterations = iters;
        stats->device_ms         = ms;
    }
    return B200_OK;
}
