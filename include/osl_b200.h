/* osl_b200.h — C ABI of libosl_b200.so, the B200-native execution back end
 * for Open Shading Language's data-parallel hot path.
 *
 * Plain C: pointers and sizes only, status codes, no exceptions cross this
 * boundary.  Every entry point names the reference interface it stands in for
 * (paths relative to the OSL source tree).  The C++ mirror of the reference
 * API (OSL::ShadingSystem / BatchedExecutor) in include/OSL/oslexec.h is
 * a thin layer over these calls; INTEGRATION.md shows the binding a reference
 * maintainer would add.
 */
#ifndef OSL_B200_H
#define OSL_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define OSL_B200_ABI_VERSION 3
#define B200_MAX_OUTPUTS 16  /* renderer outputs per group */

/* status codes */
#define B200_OK 0
#define B200_ERR_INVALID 1   /* bad argument / malformed group description   */
#define B200_ERR_COMPILE 2   /* .oso parse, code generation or NVRTC failure */
#define B200_ERR_CUDA 3      /* CUDA driver/runtime error                     */
#define B200_ERR_UNSUPPORTED 4

/* Per-point shader globals, SoA.  Field order mirrors ShaderGlobals
 * (src/include/OSL/shaderglobals.h:55-146); layout mirrors the batched
 * reference's Block<Vec3> = x[W],y[W],z[W] (src/include/OSL/wide.h:439-444)
 * with W = the whole batch. */
typedef enum b200_sg_field {
    B200_SG_P = 0, B200_SG_dPdx, B200_SG_dPdy, B200_SG_dPdz,
    B200_SG_I, B200_SG_dIdx, B200_SG_dIdy,
    B200_SG_N, B200_SG_Ng,
    B200_SG_u, B200_SG_dudx, B200_SG_dudy,
    B200_SG_v, B200_SG_dvdx, B200_SG_dvdy,
    B200_SG_dPdu, B200_SG_dPdv,
    B200_SG_time, B200_SG_dtime, B200_SG_dPdtime,
    B200_SG_Ps, B200_SG_dPsdx, B200_SG_dPsdy,
    B200_SG_surfacearea,
    B200_SG_raytype,        /* int32 */
    B200_SG_flipHandedness, /* int32 */
    B200_SG_backfacing,     /* int32 */
    B200_SG_NFIELDS
} b200_sg_field;

/* varying[f] != NULL : per-point data.  Triples are three planes
 * (x at +0, y at +plane_stride, z at +2*plane_stride elements); scalars one
 * plane; int fields are int32 planes.  varying[f] == NULL : the field is
 * uniform over the batch and uniform[f][0..2] is used (ints bit-cast in [0]),
 * like BatchedShaderGlobals' UniformShaderGlobals
 * (src/include/OSL/batched_shaderglobals.h:21-195). */
/* A named coordinate system, <name> -> "common", row-major like Imath::M44f.
 * Replaces RendererServices::get_matrix(sg, result, from, time)
 * (src/include/OSL/rendererservices.h) and, under the names "shader" and
 * "object", ShaderGlobals::shader2common / object2common
 * (src/include/OSL/shaderglobals.h:118-123): uniform over the batch. */
typedef struct b200_transform {
    const char* name;
    float m[16];
} b200_transform;
#define B200_MAX_SPACES 8 /* distinct named spaces one group may reference */

typedef struct b200_globals {
    const float* varying[B200_SG_NFIELDS];
    float uniform[B200_SG_NFIELDS][4];
    long long plane_stride;
    int ntransforms;                  /* may be 0 */
    const b200_transform* transforms; /* HOST pointer, read at launch */
} b200_globals;

/* ShadingSystem::Parameter (src/include/OSL/oslexec.h:656) */
typedef struct b200_param {
    const char* name;
    int type;              /* 0 = int, 1 = float-based, 2 = string */
    int nvalues;
    const void* values;    /* int* / float* / const char** */
} b200_param;

/* ShadingSystem::Shader (oslexec.h:723): one layer = .oso text + instance values */
typedef struct b200_layer {
    const char* oso_text;
    const char* layername;
    int nparams;
    const b200_param* params;
} b200_layer;

/* ShadingSystem::ConnectShaders (oslexec.h:740) */
typedef struct b200_connection {
    const char* srclayer;
    const char* srcparam;
    const char* dstlayer;
    const char* dstparam;
} b200_connection;

/* SymLocationDesc with arena = Outputs (oslexec.h:69-105): where a renderer
 * output lands: output_base + offset + stride*shadeindex; derivs => val,dx,dy */
typedef struct b200_symloc {
    const char* name;  /* "layer.param" or "param"; also the name of a ShaderGlobals field the entry layer
                        * uses ("P", "N", "u" ...): execute() hands the globals back as the shaders left them
                        * (ShaderGlobals is in / out of ShadingSystem::execute, oslexec.h:833) - how a
                        * displacement shader's P reaches the renderer (simpleraytracer.cpp:1365-1384) */
    long long offset;
    long long stride;
    int derivs;
} b200_symloc;

/* Userdata the renderer supplies per point for interpolated ([[ int lockgeom = 0 ]])
 * parameters: SymLocationDesc with arena = UserData (oslexec.h:69-105) and the "found" result of
 * RendererServices::get_userdata (rendererservices.h; osl_bind_interpolated_param,
 * src/liboslexec/llvm_instance.cpp:805-970).  The value of point i lies at
 *   userdata_base + offset + stride * shadeindex      (val[,dx,dy] contiguous when derivs)
 * so one dense array per name (stride = element size) gives coalesced loads.  A parameter binds
 * to the entry of the same name and type; points whose int32 at
 *   userdata_base + valid_offset + valid_stride * shadeindex
 * is zero (valid_offset >= 0) do not have the value and run the parameter's default / init ops,
 * like a get_userdata() that returns false for them. */
typedef struct b200_userdata {
    const char* name;
    int ncomp;               /* 1 (float, int) or 3 (color, point, vector, normal) */
    int is_int;
    long long offset, stride;
    int derivs;
    long long valid_offset;  /* < 0: every point has the value */
    long long valid_stride;
} b200_userdata;

/* RendererServices::get_attribute (rendererservices.h:232-262) for values that do not vary over the
 * batch - "camera:fov", "camera:resolution", object-independent renderer attributes: typed constants
 * that getattribute() returns when the requested type matches (scalars, triples and arrays of them;
 * strings).  A name known when the group is compiled folds to the value, a name computed by the
 * shader is compared against the table at run time. */
typedef struct b200_attribute {
    const char* name;
    int type;              /* 0 = int, 1 = float-based, 2 = string */
    int nvalues;           /* ints / floats: number of scalars (a float[4] or a color = 4 / 3) */
    const void* values;    /* int* / float* / const char** */
} b200_attribute;

/* ShaderGroupBegin ... ShaderGroupEnd (oslexec.h:634-650) */
typedef struct b200_group_desc {
    const char* name;
    int nlayers;
    const b200_layer* layers;
    int nconnections;
    const b200_connection* connections;
    int noutputs;
    const b200_symloc* outputs;
    /* comma separated k=v, the analogue of ShadingSystem::attribute("options"):
     *   fma=0|1   allow FMA contraction (reference: llvm_jit_fma, default 0 scalar / 1 batched)
     *   block=N   CTA size (default 256) */
    const char* options;
    int nuserdata;                 /* may be 0 */
    const b200_userdata* userdata;
    int nattributes;               /* may be 0 (ABI 3) */
    const b200_attribute* attributes;
} b200_group_desc;

typedef struct b200_group b200_group;

/* Group build + JIT: replaces BackendLLVM::run / BatchedExecutor::jit_group
 * (src/liboslexec/llvm_instance.cpp:2083, oslexec.h:982-1003).  Does not need a
 * GPU: generation + NVRTC compile to an sm_100a cubin happen here; the cubin
 * is loaded lazily on first execute. */
int b200_group_compile(const b200_group_desc* desc, b200_group** out);
void b200_group_destroy(b200_group* g);

/* Introspection (ShadingSystem::getattribute(group, ...), oslexec.h:324-596) */
const char* b200_group_cuda_source(const b200_group* g);   /* generated CUDA C++ */
const void* b200_group_cubin(const b200_group* g, long long* size);
int b200_group_num_warnings(const b200_group* g);
const char* b200_group_warning(const b200_group* g, int i);
/* 1 if the kernel reads globals field f (the reference's "globals_read" bits) */
int b200_group_reads_global(const b200_group* g, int field);

/* Execute over npoints shading points with DEVICE pointers: replaces
 * BatchedExecutor<W>::execute(ctx, group, batch_size, wide_shadeindex, bsg,
 * userdata_base, output_base) (oslexec.h:1005-1033) with batch = the whole
 * range.  shadeindex == NULL means iota.  stream is a cudaStream_t (may be 0).
 * Asynchronous with respect to the host. */
int b200_group_execute(b200_group* g, int device, void* stream, long long npoints,
                       const b200_globals* sg, const int* shadeindex,
                       const void* userdata_base, void* output_base);

/* Same call with HOST pointers: uploads only the planes the group reads, runs,
 * and downloads the renderer outputs into the host arena at
 * output_base + offset + stride*shadeindex (shadeindex = 0..npoints-1);
 * chunked and pipelined over streams so H2D, kernel and D2H overlap.  Outputs
 * that interleave in one record (same stride, offsets within a stride) move
 * as whole records.  This is the renderer-facing path (ShadingSystem::execute
 * with host ShaderGlobals, oslexec.h:833).  Synchronous.  Pinned host memory
 * makes the copies asynchronous; pageable memory works but serialises. */
int b200_group_execute_host(b200_group* g, int device, long long npoints,
                            const b200_globals* sg, void* output_base);

/* execute_host for a group with userdata: the userdata arena (HOST memory, userdata_bytes
 * long) is uploaded once per call, then as above. */
int b200_group_execute_host_userdata(b200_group* g, int device, long long npoints, const b200_globals* sg,
                                     const void* userdata_base, long long userdata_bytes, void* output_base);

/* The general host-pointer call: the batch's points carry the shade indices
 * first_shadeindex .. first_shadeindex + npoints - 1 (what a renderer shading a tile of a larger
 * arena passes as wide_shadeindex, src/testshade/testshade.cpp:1868): outputs land at
 * output_base + offset + stride * shadeindex and userdata is read at the same index.
 * userdata_base may be NULL. */
int b200_group_execute_host_at(b200_group* g, int device, long long npoints, const b200_globals* sg,
                               long long first_shadeindex, const void* userdata_base, long long userdata_bytes,
                               void* output_base);

/* Group option userdata=record: instead of the caller describing its arrays (b200_userdata in
 * the group description) the library lays out one record per point holding every interpolated
 * parameter of the group (validity word + value with derivatives).  These calls say how many
 * fields there are, how long a record is and where each field went; pass the records as
 * userdata_base (to the host calls: the records of the batch only, record 0 = its first point).  Also: the named coordinate systems the group references, which the caller
 * must supply in b200_globals.transforms (RendererServices::get_matrix). */
int b200_group_userdata_fields(const b200_group* g, long long* record_bytes);
int b200_group_userdata_field(const b200_group* g, int i, b200_userdata* out);
int b200_group_num_spaces(const b200_group* g);
const char* b200_group_space_name(const b200_group* g, int i);

/* Text written by the group's printf() ops on `device` since the previous call, ordered by
 * shade index and, within a point, by execution order (what single-threaded testshade
 * prints).  The device only records (format id, argument words); formatting happens here,
 * like the reference's journal (src/include/OSL/journal.h, rs_printfmt in
 * rs_free_function.h).  Synchronises the device.  The pointer is owned by the group and
 * valid until the next call.  Recording is opt-in: group option journal=1 (default buffer,
 * 4 Mi words) or journal=WORDS; without it printf ops are dropped and reported in the
 * group's warnings, so production launches never pay for a debugging aid. */
const char* b200_group_journal(b200_group* g, int device);

/* Image data for texture() (src/liboslexec/optexture.cpp:235-310 osl_texture ->
 * RendererServices::texture, rendservices.cpp:166-232, which the reference forwards to the
 * renderer's OIIO TextureSystem, include/OSL/oslexec.h:172).  The renderer registers a
 * decoded image under the file name its shaders use: `pixels` is [height][width][nchannels]
 * float32 in host memory, top scanline first, nchannels 1..4; it is copied.  Names that
 * were not registered are looked up as Radiance .hdr files (as written, then under the
 * ':'-separated directories of the group / render option texturepath=...) when a group
 * that reads them is first launched on a device.  File names must be constant once the
 * group is compiled (instance parameter values are).  Filtering: level-0 B-spline bicubic
 * with anisotropic probes, options wrap/swrap/twrap, width, blur, fill, interp. */
int b200_texture_add(const char* name, int width, int height, int nchannels, const float* pixels);

/* Number of kernel launches issued by this library so far (bench accounting) */
long long b200_launch_count(void);

/* ---- device shadeop library, batch entry points (device pointers) --------
 * Stand-alone equivalents of the osl_<op>_<codes> runtime
 * (src/liboslexec/builtindecl.h:19-83, opnoise.cpp:263-273, 468-473) over
 * SoA arrays.  kind: 0 noise 1 snoise 2 cellnoise 3 hashnoise 4 simplex 5 usimplex.
 * in:  indim planes of n floats (+ 2*indim derivative planes when derivs)
 * out: outdim planes of n floats (x3 when derivs: val planes, dx planes, dy planes)
 * period (pnoise family): indim floats, or NULL. */
int b200_shadeop_noise(int kind, int outdim, int indim, int derivs, int fma,
                       long long n, const float* in, const float* period,
                       float* out, void* stream);
/* osl_hash_* (builtindecl.h:169-173): indim floats per point -> int32 */
int b200_shadeop_hash(int indim, long long n, const float* in, int* out, void* stream);

/* ---- wavefront path tracer (the testrender hot path) ------------------------
 * Replaces SimpleRaytracer::render / antialias_pixel / subpixel_radiance
 * (src/testrender/simpleraytracer.cpp:957-1216, 1424-1456) and, for the OptiX
 * variant, optixLaunch of __raygen__deferred (src/testrender/cuda/optix_raytracer.cu:291).
 * The renderer keeps owning scene loading, tessellation and the BVH build
 * (bvh.cpp:43-219); it hands over the prepared arrays (HOST pointers, copied
 * to the device once) plus one group description per material. */
typedef struct b200_render_scene {
    int nverts, ntris, nnodes, nlightprims, nshaders, nmeshes;
    const float* verts;            /* 3 per vertex (Scene::verts)              */
    const float* normals;          /* 3 per normal                             */
    const float* uvs;              /* 2 per uv                                 */
    const int* triangles;          /* 3 per triangle (TriangleIndices)         */
    const int* n_triangles;        /* 3 per triangle, -1 = none                */
    const int* uv_triangles;       /* 3 per triangle, -1 = none                */
    const int* shaderids;          /* per triangle                             */
    const int* meshids;            /* per triangle                             */
    const float* mesh_surfacearea; /* per mesh (m_mesh_surfacearea)            */
    const float* bvh_nodes;        /* 8 words per BVHNode: bounds[6],child,nprims (bvh.h) */
    const unsigned* bvh_indices;
    const unsigned* lightprims;    /* m_lightprims                             */
    const int* shader_is_light;    /* per material                             */
    float eye[3], dir[3], up[3], fov;   /* Camera::lookat arguments            */
    float cx[3], cy[3], invw, invh;     /* derived; filled by the library      */
    int xres, yres;
    int aa, max_bounces, rr_depth, no_jitter, show_globals; /* testrender options */
    int background_shader, background_resolution;
} b200_render_scene;

typedef struct b200_render_stats {
    long long paths;        /* camera samples traced                           */
    long long launches;     /* kernels launched                                */
    long long bounce_iterations;
    double device_ms;       /* CUDA-event time of the whole call               */
    double tail_ms;         /* of which: the final rt_tail launches            */
    long long slots;        /* path slots in the pool                          */
    long long rounds;       /* regeneration rounds (1 unless spp x pixels > 2^30) */
} b200_render_stats;

typedef struct b200_render b200_render;

/* Scene BVH of the testrender path: binned SAH over the triangles, the tree of the reference's
 * build_bvh (src/testrender/bvh.cpp:42-237: 16 bins, depth <= 64, in-place partition, children
 * numbered depth first), float32 operation for operation, so that traversal order - and with it
 * every render - matches.  nodes: room for max_nodes x 8 words (2 * ntriangles - 1 always
 * suffices), layout as b200_render_scene::bvh_nodes; indices: ntriangles words. */
int b200_build_bvh(const float* verts, int nverts, const int* triangles, int ntriangles, float* nodes, int max_nodes,
                   unsigned* indices, int* nnodes);

/* options: fma=0|1, sort=0|1 (order live paths by closure signature / material),
 * slots=N (path slots in the regenerating pool), tail=N (run the last N paths in one launch),
 * compile=0 (generate the CUDA source only; NVRTC runs at the first render or cubin request) */
int b200_render_create(const b200_render_scene* scene, int nmaterials, const b200_group_desc* materials,
                       const char* options, b200_render** out);
void b200_render_destroy(b200_render* r);
const char* b200_render_cuda_source(const b200_render* r);
const void* b200_render_cubin(const b200_render* r, long long* size);   /* the sm_100a module */
/* Render image rows [y0, y1) into host_rgb ((y1-y0)*xres*3 floats).  Synchronous. */
int b200_render_rows(b200_render* r, int device, int y0, int y1, float* host_rgb, b200_render_stats* stats);
/* Render a work set of image tiles: tiles = ntiles x {x0, y0, width, height}.  This is how a
 * frame is shared between GPUs (interleaved tiles per GPU; the reference spreads
 * scanline chunks over host threads, simpleraytracer.cpp:1428).  Pixels are numbered tile
 * after tile, row-major inside a tile; out_rgb receives 3 floats per pixel in that order, in
 * host memory or, with out_on_device != 0, in device memory of `device` (for a framebuffer
 * gather without a host round trip).  Synchronous. */
int b200_render_tiles(b200_render* r, int device, int ntiles, const int* tiles, void* out_rgb,
                      int out_on_device, b200_render_stats* stats);

const char* b200_last_error(void);
int b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* OSL_B200_H */
