// osl_b200_render.cuh — wavefront path tracer for sm_100a (product code).
//
// B200-native replacement of the reference's recursive per-pixel integrator
//   SimpleRaytracer::subpixel_radiance / antialias_pixel / globals_from_hit
//       src/testrender/simpleraytracer.cpp:889-1216
//   Scene::intersect (BVH + watertight triangle test)   src/testrender/bvh.cpp:221-356
//   Scene::sample/shapepdf/normal/project/uv, Ray, Camera  src/testrender/raytracer.h:40-344
//   Sampler, MIS, TangentFrame, Sampling               src/testrender/sampling.h:17-295
//   CompositeBSDF + Diffuse/Reflection/Refraction/Transparent lobes
//       src/testrender/shading.h:319-437, shading.cpp:301-322, 1081-1150
//   process_closure                                    src/testrender/shading.cpp:1448-1706
//   fresnel_dielectric / fresnel_refraction            src/testrender/optics.h:13-60
//
// Structure: a pool of path slots lives in HBM, one 128-byte record per slot.  A render is a
// loop of "bounce steps" over the pool; every step is
//   rt_trace         one persistent kernel walks the BVH for ALL rays of the step: the closest
//                    hit of every live path and the NEE shadow rays (any-hit) the previous
//                    shade emitted.  Warps refill finished lanes from the queues (dynamic
//                    fetch), traversal stacks are staged in shared memory.
//   rt_light         NEE of the previous bounce: unoccluded shadow rays add their contribution
//                    (light shader executed here), finished paths are written out and their
//                    slot is REGENERATED with the next camera sample.
//   rt_sort_scatter  counting sort of the live queue by closure-type signature / material
//                    (histogram accumulated by rt_trace in shared memory)
//   rt_shade         globals from hit -> material dispatch -> closure arena (shared memory)
//                    -> lobes -> emission -> NEE set-up -> BSDF sample; dead paths regenerate.
// so a long path (glass interiors live for 10^5 bounces) never drains the machine: the pool
// stays full until the samples of the work set run out.  Every finished sample stores its
// radiance in a per-sample slot and rt_resolve folds them in the reference's order (running
// lerp over the sample index), so in strict mode the image is bit-identical to the scalar
// CPU oracle whatever the scheduling was.
//
// This header is included by the generated render module AFTER the material
// namespaces and `osl_execute_shader(int shaderID, SG&)` have been emitted.
#pragma once

namespace osld {

// ---- scene + state ------------------------------------------------------------
struct RenderScene {
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
    const float4* bvh_nodes;  // 2 x float4 per node: bounds[6], child, nprims
    const unsigned* bvh_indices;
    const unsigned* lightprims;
    const int* shader_is_light;
    float eye[3], dir[3], up[3], fov;
    float cx[3], cy[3], invw, invh;
    int xres, yres;
    int aa, max_bounces, rr_depth, no_jitter, show_globals;
    int background_shader, background_resolution;
    // background importance table (background.h:38-276), built on the device by the
    // rt_bg_* kernels; null when the scene has no importance-sampled background
    float* bg_values;  // res*res RGB, already divided by the texel pdf
    float* bg_rows;    // res: CDF over rows
    float* bg_cols;    // res*res: per-row CDF over columns
    int bg_res;
    float bg_invres, bg_invjacobian;
    // triangles in BVH leaf order, 3 x float4 each: vertex a (w = primitive id bits), b, c.
    // The leaf loop reads them with three contiguous 16-byte loads instead of the reference's
    // index -> triangle -> vertex gather (same values, one dependent load level instead of three).
    const float4* leaf_tris;
    // energy-compensation tables of the MaterialX microfacet closures (osl_b200_mxlobes.cuh), or null
    const float* bsdl_luts;
};

// Path state: one 128-byte record (8 x float4) per path slot.  After the live
// queue has been sorted the slots a warp touches are scattered, so
// the state is laid out per path (every 32-byte sector fetched is fully used
// and moved with 16-byte LDG/STG) rather than as per-field planes.
//   q0 = origin.xyz, radius      q1 = direction.xyz, spread
//   q2 = path_weight.rgb, bsdf_pdf   q3 = path_radiance.rgb, roughness
//   q4 = hit t,u,v, hit id       q5 = raytype, prev_id, bounce, sampler seed
//   q6 = sampler index, sample id, -, -  q7 = light sample u, v, -, -
// Pending NEE shadow rays of a slot: 4 x float4 in a second array
//   s0 = background dir.xyz, flags   s1 = background contribution.rgb, visibility bits
//   s2 = light dir.xyz, distance     s3 = light contribution.rgb, light primitive id
#define OSLD_PATH_QUADS 8
#define OSLD_SHADOW_QUADS 4
OSLD float4 mkf4(V3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
OSLD V3 xyz(float4 q) { return mkv(q.x, q.y, q.z); }
OSLD float4 mki4(int a, int b, int c, int d)
{
    return make_float4(__int_as_float(a), __int_as_float(b), __int_as_float(c), __int_as_float(d));
}
enum { SH_BG = 1, SH_LIGHT = 2, SH_DEAD = 4 };

// counters[]: device-side loop state of a render call
enum { C_LIVE = 0,     // paths in queue_in
       C_OUT,          // paths appended to queue_out by this step
       C_SHADOW,       // entries of queue_sh to trace + light in this step
       C_SHADOW_OUT,   // entries appended to queue_sh by this step's shade
       C_NEXT,         // next sample id of the round to start
       C_FETCH,        // rt_trace's dynamic-fetch cursor
       C_ITER,         // bounce steps done
       C_FINISHED,     // samples written out
       C_HIST   = 8,   // 64 sort buckets: histogram
       C_CURSOR = 8 + 64,
       C_WORDS  = 8 + 128 };

struct RenderLaunch {
    RenderScene S;
    float4* rec;          // nslots x OSLD_PATH_QUADS
    float4* shrec;        // nslots x OSLD_SHADOW_QUADS
    int* queue_in;        // live path slots
    int* queue_out;
    int* queue_sh;        // slots with pending shadow rays
    int* counters;
    int* sort_keys;       // bucket per queue_in entry (null: no sort)
    const int* shader_key;  // bucket of a material: closure-type signature rank, then material
    const int* pixmap;    // the work set: packed (y << 16) | x per pixel, tile after tile
    volatile int* host_state;  // mapped pinned host memory: {iter, live, shadow, next} after every step
    int nslots;           // path slots in the pool
    int npix;             // pixels in the work set
    int s0;               // first sample plane of the round
    int nsamples;         // sample planes in the round
    int total;            // samples in the round = nsamples * npix
    float* result;        // total x 3: radiance of every sample of the round
    float* accum;         // running per-pixel result, 3 floats per work-set pixel
    float* medium;        // nslots x OSLD_MEDIUM_WORDS: the medium stack of every path (null: no media)
};

OSLD V3 ld3(const float* p, int i) { return mkv(__ldg(p + 3 * i), __ldg(p + 3 * i + 1), __ldg(p + 3 * i + 2)); }
OSLD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OSLD float len2(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
OSLD V3 vdiv(V3 a, float l) { return mkv(a.x / l, a.y / l, a.z / l); }
OSLD V3 vnormalized(V3 v)
{
    float l = imath_length(v);
    return l != 0.0f ? vdiv(v, l) : v;
}
OSLD float vcomp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
// (v[kx], v[ky], v[kz]) for the watertight test's axis renaming, kx = (kz + 1) % 3, ky = (kx + 1) % 3:
// two predicates per ray and selects per component, no branches in the triangle loop
OSLD V3 vperm_kz(V3 v, bool z0, bool z1)
{
    return mkv(z0 ? v.y : (z1 ? v.z : v.x), z0 ? v.z : (z1 ? v.x : v.y), z0 ? v.x : (z1 ? v.y : v.z));
}

// ---- Ray ---------------------------------------------------------------------------
enum { RAY_CAMERA = 1, RAY_SHADOW = 2, RAY_DIFFUSE = 16 };
struct Ray {
    V3 origin, direction;
    float radius, spread, roughness;
    int raytype;
};
OSLD void ortho(V3 n, V3& x, V3& y)
{
    x = vnormalized(fabsf(n.x) > .01f ? mkv(n.z, 0.0f, -n.x) : mkv(0.0f, -n.z, n.y));
    y = cross3(n, x);
}
OSLD V3 ray_point(const Ray& r, float t) { return r.origin + r.direction * t; }

// ---- Sampler (integer only: bit-exact) -------------------------------------------------
struct Sampler {
    u32 seed, index;
    OSLD static u32 hash(u32 s)
    {
        s ^= s >> 16; s *= 0x21f0aaadu; s ^= s >> 15; s *= 0xd35a2d97u; s ^= s >> 15;
        return s;
    }
    OSLD static u32 owen(u32 p, u32 s)
    {
        p ^= p * 0x3d20adeau; p += s; p *= (s >> 16) | 1u; p ^= p * 0x05526c56u; p ^= p * 0x53a22864u;
        return __brev(p);
    }
    OSLD void init(int px, int py, int si)
    {
        seed  = (u32)(((px & 2047) << 22) | ((py & 2047) << 11));
        index = __brev((u32)si);
    }
    OSLD V3 get()
    {
        const u32 zmatrix[24] = { 0x000001u, 0x000003u, 0x000006u, 0x000009u, 0x000017u, 0x00003au, 0x000071u, 0x0000a3u,
                                  0x000116u, 0x000339u, 0x000677u, 0x0009aau, 0x001601u, 0x003903u, 0x007706u, 0x00aa09u,
                                  0x010117u, 0x03033au, 0x060671u, 0x0909a3u, 0x171616u, 0x3a3939u, 0x717777u, 0xa3aaaau };
        seed += 4;
        u32 si = owen(index, hash(seed - 4)) & 0xFFFFFFu;
        u32 rx = si, ry = 0, rz = 0, ym = 1;
#pragma unroll
        for (int c = 0; c < 24; c++) {
            u32 bit = (si >> c) & 1u;
            ry ^= bit * ym;
            rz ^= bit * zmatrix[c];
            ym ^= ym << 1;
        }
        return mkv((float)(owen(rx, hash(seed - 3)) >> 8) * 5.96046448e-8f,
                   (float)(owen(ry, hash(seed - 2)) >> 8) * 5.96046448e-8f,
                   (float)(owen(rz, hash(seed - 1)) >> 8) * 5.96046448e-8f);
    }
};

// ---- sampling helpers ---------------------------------------------------------------
struct TangentFrame {
    V3 u, v, w;
};
OSLD TangentFrame frame_from_normal(V3 n)
{
    const float sign = copysignf(1.0f, n.z);
    const float a    = -1 / (sign + n.z);
    const float b    = n.x * n.y * a;
    TangentFrame f;
    f.u = mkv(1 + sign * n.x * n.x * a, sign * b, -sign * n.x);
    f.v = mkv(b, sign + n.y * n.y * a, -n.y);
    f.w = n;
    return f;
}
OSLD V3 frame_get(const TangentFrame& f, float x, float y, float z) { return x * f.u + y * f.v + z * f.w; }
OSLD void to_unit_disk(float& x, float& y)
{
    const float PI_OVER_4 = (float)(OSLD_PI / 4), PI_OVER_2 = (float)(OSLD_PI / 2);
    float phi, r;
    float a = 2 * x - 1, b = 2 * y - 1;
    if (a * a > b * b) {
        r   = a;
        phi = PI_OVER_4 * (b / a);
    } else if (b != 0) {
        r   = b;
        phi = PI_OVER_2 - PI_OVER_4 * (a / b);
    } else {
        r   = 0;
        phi = 0;
    }
    fast_sincos(phi, &x, &y);
    x *= r;
    y *= r;
}
OSLD void sample_cosine_hemisphere(V3 N, float rndx, float rndy, V3& out, float& pdf)
{
    to_unit_disk(rndx, rndy);
    float cos_theta = sqrtf(fmaxf(1 - rndx * rndx - rndy * rndy, 0.0f));
    out             = frame_get(frame_from_normal(N), rndx, rndy, cos_theta);
    pdf             = cos_theta * (float)(1.0 / OSLD_PI);
}
enum { WEIGHT_WEIGHT, WEIGHT_EVAL, EVAL_WEIGHT };
template<int mode> OSLD float power_heuristic(float sampled_pdf, float other_pdf)
{
    float r, mis;
    if (sampled_pdf > other_pdf) {
        r   = other_pdf / sampled_pdf;
        mis = 1 / (1 + r * r);
    } else if (sampled_pdf < other_pdf) {
        r   = sampled_pdf / other_pdf;
        mis = 1 - 1 / (1 + r * r);
    } else {
        r   = 1.0f;
        mis = 0.5f;
    }
    if (mode == WEIGHT_WEIGHT)
        return fminf(other_pdf, OSLD_FLT_MAX) * mis;
    if (mode == WEIGHT_EVAL)
        return mis;
    return mis * ((other_pdf > sampled_pdf) ? fminf(1 / r, OSLD_FLT_MAX) : r);
}
OSLD void update_eval(V3* w, float* pdf, V3 ow, float opdf, float b)
{
    if (b > OSLD_FLT_MIN) {
        opdf *= b;
        ow = ow * (1 / b);
        float mis;
        if (*pdf < opdf)
            mis = 1 / (1 + *pdf / opdf);
        else if (opdf < *pdf)
            mis = 1 - 1 / (1 + opdf / *pdf);
        else
            mis = 0.5f;
        *w = *w * (1 - mis) + ow * mis;
        *pdf += opdf;
    }
}

// ---- fresnel -----------------------------------------------------------------------
OSLD float fresnel_dielectric(float cosi, float eta)
{
    if (eta == 0)
        return 1;
    if (cosi < 0.0f)
        eta = 1.0f / eta;
    float c = fabsf(cosi);
    float g = eta * eta - 1 + c * c;
    if (g > 0) {
        g       = sqrtf(g);
        float A = (g - c) / (g + c);
        float B = (c * (g + c) - 1) / (c * (g - c) + 1);
        return 0.5f * A * A * (1 + B * B);
    }
    return 1.0f;
}
OSLD float fresnel_refraction(V3 I, V3 N, float eta, V3& T)
{
    float cosi = -dot3(I, N);
    V3 Nn;
    float neta;
    if (cosi > 0) {
        neta = 1 / eta;
        Nn   = N;
    } else {
        cosi = -cosi;
        neta = eta;
        Nn   = -N;
    }
    float arg = 1.0f - (neta * neta * (1.0f - cosi * cosi));
    if (arg >= 0) {
        float dnp = sqrtf(arg);
        float nK  = (neta * cosi) - dnp;
        T         = I * neta + Nn * nK;
        return 1 - fresnel_dielectric(cosi, eta);
    }
    T = mkv(0.0f);
    return 0;
}

// ---- lobes ---------------------------------------------------------------------------
struct BSample {
    V3 wi, weight;
    float pdf, roughness;
};
OSLD BSample bs_null()
{
    BSample s;
    s.wi = mkv(0.0f);
    s.weight = mkv(0.0f);
    s.pdf = 0.0f;
    s.roughness = 0.0f;
    return s;
}
OSLD BSample bs_make(V3 wi, V3 w, float pdf, float r)
{
    BSample s;
    s.wi = wi;
    s.weight = w;
    s.pdf = pdf;
    s.roughness = r;
    return s;
}
#define OSLD_INF __int_as_float(0x7f800000)
enum { LOBE_DIFFUSE, LOBE_TRANSLUCENT, LOBE_REFLECTION, LOBE_REFRACTION, LOBE_TRANSPARENT,
       LOBE_PHONG, LOBE_WARD, LOBE_MICROFACET, LOBE_BSDL_OREN_NAYAR, LOBE_BSDL_BURLEY, LOBE_BSDL_SHEEN,
       LOBE_MX_SPEC /* conductor / dielectric / generalized schlick */, LOBE_MX_TRANSLUCENT,
       LOBE_SPI_THINLAYER /* spi::ThinLayerLobe */ };
}  // namespace osld
#ifdef OSLD_MX_LOBES
#include "osl_b200_mxlobes.cuh"
#endif
namespace osld {
struct Lobe {
    int type;
    V3 N;
    float eta;
#ifdef OSLD_GLOSSY_LOBES
    // phong: ax = exponent.  ward / microfacet: tangent frame (fu, fv, N) and roughnesses.
    // libbsdl diffuse lobes: frame in (fu, fv, N), roughness in ax, the
    // energy-compensation flag in refract, albedo below.
    V3 fu, fv;
    float ax, ay;
    int refract, ggx;
    V3 albedo;
#ifdef OSLD_SHEEN_LTC
    V3 ltc;      // Zeltner-Burley sheen: the view's (A, B, R) coefficients
#endif
#endif
#ifdef OSLD_MX_LOBES
    MxSpec mx;   // LOBE_MX_SPEC: frame in (fu, fv, N)
#endif
#ifdef OSLD_THINLAYER
    ThinSpec thin;   // LOBE_SPI_THINLAYER: frame in (fu, fv, N)
#endif
};
}  // namespace osld
#ifdef OSLD_GLOSSY_LOBES
#include "osl_b200_lobes.cuh"
#endif
#ifdef OSLD_THINLAYER
#include "osl_b200_thinlayer.cuh"
#endif
namespace osld {
OSLD V3 lobe_albedo(const Lobe& l, V3 wo)
{
#ifdef OSLD_GLOSSY_LOBES
    if (l.type == LOBE_MICROFACET)
        return mf_albedo(l, wo);
    if (l.type == LOBE_BSDL_OREN_NAYAR || l.type == LOBE_BSDL_BURLEY)
        return l.albedo;  // BSDL_WRAP::get_albedo = albedo_impl().toRGB(0)
    if (l.type == LOBE_BSDL_SHEEN)
        return l.albedo * (1 - l.eta);  // tint * (1 - Emiss)
#endif
#ifdef OSLD_MX_LOBES
    if (l.type == LOBE_MX_SPEC)
        return mx_albedo(l.mx);
    if (l.type == LOBE_MX_TRANSLUCENT)
        return l.albedo;
#endif
    if (l.type == LOBE_REFLECTION) {
        float cosNO = dot3(l.N, wo);
        return cosNO > 0 ? mkv(fresnel_dielectric(cosNO, l.eta)) : mkv(1.0f);
    }
    if (l.type == LOBE_REFRACTION)
        return mkv(1 - fresnel_dielectric(dot3(l.N, wo), l.eta));
    return mkv(1.0f);
}
OSLD BSample lobe_eval(const Lobe& l, V3 wo, V3 wi)
{
    if (l.type == LOBE_DIFFUSE || l.type == LOBE_TRANSLUCENT) {
        const float pdf = fmaxf(dot3(l.N, wi), 0.0f) * (float)(1.0 / OSLD_PI);
        return bs_make(wi, mkv(1.0f), pdf, 1.0f);
    }
#ifdef OSLD_GLOSSY_LOBES
    if (l.type == LOBE_PHONG)
        return phong_eval(l, wo, wi);
    if (l.type == LOBE_WARD)
        return ward_eval(l, wo, wi);
    if (l.type == LOBE_MICROFACET)
        return mf_eval(l, wo, wi);
    if (l.type == LOBE_BSDL_OREN_NAYAR || l.type == LOBE_BSDL_BURLEY)
        return bsdl_diffuse_eval(l, wo, wi);
    if (l.type == LOBE_BSDL_SHEEN)
        return sheen_eval(l, wo, wi);
#endif
#ifdef OSLD_MX_LOBES
    if (l.type == LOBE_MX_SPEC) {   // BSDL_WRAP::eval (shading.cpp:88-93)
        BSample s = mx_eval_local(l.mx, frame_tolocal(l, wo), frame_tolocal(l, wi));
        s.wi      = wi;
        return s;
    }
#ifdef OSLD_THINLAYER
    if (l.type == LOBE_SPI_THINLAYER) {   // SpiThinLayer::eval (shading.cpp:138-143)
        BSample s = thin_eval_local(l.thin, frame_tolocal(l, wo), frame_tolocal(l, wi));
        s.wi      = wi;
        return s;
    }
#endif
    if (l.type == LOBE_MX_TRANSLUCENT) {   // mtx::TranslucentLobe (bsdf_translucent_impl.h)
        const float z = dot3(wi, l.N);
        if (z >= 0.0f)
            return bs_make(wi, mkv(0.0f), 0.0f, 0.0f);
        return bs_make(wi, l.albedo, fabsf(z) * (1 / (float)OSLD_PI), 1.0f);
    }
#endif
    return bs_null();
}
OSLD BSample lobe_sample(const Lobe& l, V3 wo, float rx, float ry, float rz)
{
    switch (l.type) {
    case LOBE_DIFFUSE:
    case LOBE_TRANSLUCENT: {
        V3 out;
        float pdf;
        sample_cosine_hemisphere(l.N, rx, ry, out, pdf);
        return bs_make(out, mkv(1.0f), pdf, 1.0f);
    }
    case LOBE_REFLECTION: {
        float cosNO = dot3(l.N, wo);
        if (cosNO > 0) {
            V3 wi = (2 * cosNO) * l.N - wo;
            return bs_make(wi, mkv(fresnel_dielectric(cosNO, l.eta)), OSLD_INF, 0.0f);
        }
        return bs_null();
    }
    case LOBE_REFRACTION: {
        V3 wi;
        float Ft = fresnel_refraction(-wo, l.N, l.eta, wi);
        return bs_make(wi, mkv(Ft), OSLD_INF, 0.0f);
    }
#ifdef OSLD_GLOSSY_LOBES
    case LOBE_PHONG: return phong_sample(l, wo, rx, ry);
    case LOBE_WARD: return ward_sample(l, wo, rx, ry);
    case LOBE_MICROFACET: return mf_sample(l, wo, rx, ry, rz);
    case LOBE_BSDL_OREN_NAYAR:
    case LOBE_BSDL_BURLEY: return bsdl_diffuse_sample(l, wo, rx, ry);
    case LOBE_BSDL_SHEEN: return sheen_sample(l, wo, rx, ry);
#endif
#ifdef OSLD_MX_LOBES
    case LOBE_MX_SPEC: {   // BSDL_WRAP::sample (shading.cpp:94-101)
        BSample s = mx_sample_local(l.mx, frame_tolocal(l, wo), rx, ry, rz);
        s.wi      = frame_toworld(l, s.wi);
        return s;
    }
#ifdef OSLD_THINLAYER
    case LOBE_SPI_THINLAYER: {   // SpiThinLayer::sample (shading.cpp:144-151)
        BSample s = thin_sample_local(l.thin, frame_tolocal(l, wo), mkv(rx, ry, rz));
        s.wi      = frame_toworld(l, s.wi);
        return s;
    }
#endif
    case LOBE_MX_TRANSLUCENT: {
        V3 wi_l = bsdl_sample_cos_hemisphere(rx, ry);
        wi_l.z  = -wi_l.z;
        if (wi_l.z >= 0.0f)
            return bs_make(frame_toworld(l, mkv(0.0f)), mkv(0.0f), 0.0f, 0.0f);
        return bs_make(frame_toworld(l, wi_l), l.albedo, fabsf(wi_l.z) * (1 / (float)OSLD_PI), 1.0f);
    }
#endif
    default: return bs_make(-wo, mkv(1.0f), OSLD_INF, 0.0f);
    }
}

// The reference's CompositeBSDF holds at most 8 lobes (shading.h:319).  The code generator
// knows how many BSDF components the scene's materials can create (OSLD_MAX_LOBES <= 8) and how
// deep their add / layer trees go (OSLD_CLOSURE_STACK <= 16), so the per-thread arrays are
// sized for the scene, not for the worst case.
#ifndef OSLD_MAX_LOBES
#define OSLD_MAX_LOBES 8
#endif
#ifndef OSLD_CLOSURE_STACK
#define OSLD_CLOSURE_STACK 16
#endif
struct CompositeBSDF {
    V3 weights[OSLD_MAX_LOBES];
    float pdfs[OSLD_MAX_LOBES];
    Lobe lobes[OSLD_MAX_LOBES];
    int num;
};
OSLD void bsdf_prepare(CompositeBSDF& B, V3 wo, V3 path_weight, bool absorb)
{
    float total = 0;
    for (int i = 0; i < B.num; i++) {
        B.pdfs[i] = dot3(B.weights[i], path_weight * lobe_albedo(B.lobes[i], wo))
                    / (path_weight.x + path_weight.y + path_weight.z);
        total += B.pdfs[i];
    }
    if ((!absorb && total > 0) || total > 1)
        for (int i = 0; i < B.num; i++)
            B.pdfs[i] /= total;
}
OSLD BSample bsdf_eval(const CompositeBSDF& B, V3 wo, V3 wi)
{
    BSample s = bs_null();
    for (int i = 0; i < B.num; i++) {
        BSample b = lobe_eval(B.lobes[i], wo, wi);
        b.weight  = b.weight * B.weights[i];
        update_eval(&s.weight, &s.pdf, b.weight, b.pdf, B.pdfs[i]);
        s.roughness += b.roughness * B.pdfs[i];
    }
    return s;
}
OSLD BSample bsdf_sample(const CompositeBSDF& B, V3 wo, float rx, float ry, float rz)
{
    float accum = 0;
    for (int i = 0; i < B.num; i++) {
        if (rx < (B.pdfs[i] + accum)) {
            rx        = (rx - accum) / B.pdfs[i];
            rx        = fminf(rx, 0.99999994f);
            BSample s = lobe_sample(B.lobes[i], wo, rx, ry, rz);
            s.weight  = s.weight * (B.weights[i] * (1 / B.pdfs[i]));
            s.pdf *= B.pdfs[i];
            if (s.pdf == 0.0f)
                return bs_null();
            for (int j = 0; j < B.num; j++) {
                if (i != j) {
                    BSample b = lobe_eval(B.lobes[j], wo, s.wi);
                    b.weight  = b.weight * B.weights[j];
                    update_eval(&s.weight, &s.pdf, b.weight, b.pdf, B.pdfs[j]);
                }
            }
            return s;
        }
        accum += B.pdfs[i];
    }
    return bs_null();
}

#ifdef OSLD_GLOSSY_LOBES
// evaluate_layer_opacity (shading.cpp:1198-1282): what the top stack of a layer() takes;
// returns the weight held when the walk ends, as the reference does
OSLD V3 evaluate_layer_opacity(const ClosurePool& pool, int closure, V3 wo, bool backfacing, float path_roughness,
                               const float* luts)
{
    if (!closure)
        return mkv(0.0f);
    int ptr_stack[OSLD_CLOSURE_STACK];
    V3 weight_stack[OSLD_CLOSURE_STACK];
    int sp    = 0;
    V3 weight = mkv(1.0f);
    while (closure) {
        int id = pool.id(closure);
        if (id == CL_MUL) {
            weight  = weight * pool.weight(closure);
            closure = __float_as_int(pool.w[closure + 4]);
        } else if (id == CL_ADD) {
            ptr_stack[sp]      = __float_as_int(pool.w[closure + 2]);
            weight_stack[sp++] = weight;
            closure            = __float_as_int(pool.w[closure + 1]);
        } else {
            const V3 w     = pool.weight(closure);
            const PoolPtr q = pool.w + (closure + 4);
            closure        = 0;
            if (id == MX_LAYER_ID) {
                closure            = __float_as_int(q[0]);
                ptr_stack[sp]      = __float_as_int(q[1]);
                weight_stack[sp++] = weight * w;
            } else if (id == REFLECTION_ID || id == FRESNEL_REFLECTION_ID) {
                Lobe l;
                l.type = LOBE_REFLECTION;
                l.N    = mkv(q[0], q[1], q[2]);
                l.eta  = id == FRESNEL_REFLECTION_ID ? q[3] : 0.0f;
                weight = weight * (w * lobe_albedo(l, wo));
            } else if (id == MX_SHEEN_ID) {
                Lobe l;
                l.N      = mkv(q[0], q[1], q[2]);
                l.albedo = mkv(q[3], q[4], q[5]);
                sheen_setup(l, wo, q[6], backfacing, path_roughness, __float_as_int(q[7]), luts);
                weight = weight * (w * (mkv(1.0f) - mkv(l.eta)));
#ifdef OSLD_MX_LOBES
            } else if (id == MX_DIELECTRIC_ID) {
                Lobe l;
                mx_from_component(luts, l, id, q, wo, backfacing, path_roughness);
                weight = weight * (w * (mkv(1.0f) - mx_filter_o(l.mx, false)));
            } else if (id == MX_GENERALIZED_SCHLICK_ID) {
                // transmissive dielectrics are opaque to the layer below
                if (q[9] == 0 && q[10] == 0 && q[11] == 0) {
                    Lobe l;
                    mx_from_component(luts, l, id, q, wo, backfacing, path_roughness);
                    weight = weight * (w * (mkv(1.0f) - mx_filter_o(l.mx, true)));
                }
#endif
            }  // anything else: opaque
        }
        if (closure == 0 && sp > 0) {
            closure = ptr_stack[--sp];
            weight  = weight_stack[sp];
        }
    }
    return weight;
}
#endif

// closure tree -> emission + lobes (16-deep explicit stack, weights root->leaf)
OSLD void process_closure(const ClosurePool& pool, int closure, V3& Le, CompositeBSDF& B, bool light_only,
                          V3 wo = mkv(0.0f, 0.0f, 1.0f), bool backfacing = false, float path_roughness = 0.0f,
                          const float* luts = nullptr, bool false_intersection = false)
{
    int ptr_stack[OSLD_CLOSURE_STACK];
    V3 weight_stack[OSLD_CLOSURE_STACK];
    int sp    = 0;
    V3 weight = mkv(1.0f);
    while (closure) {
        int id = pool.id(closure);
        if (id == CL_MUL) {
            weight  = weight * pool.weight(closure);
            closure = __float_as_int(pool.w[closure + 4]);
        } else if (id == CL_ADD) {
            ptr_stack[sp]      = __float_as_int(pool.w[closure + 2]);
            weight_stack[sp++] = weight;
            closure            = __float_as_int(pool.w[closure + 1]);
        } else {
            V3 cw          = weight * pool.weight(closure);
            const PoolPtr q = pool.w + (closure + 4);
            closure        = 0;
            if (id == EMISSION_ID)
                Le = Le + cw;
            else if (id == MX_UNIFORM_EDF_ID)
                Le = Le + cw * mkv(q[0], q[1], q[2]);
            else if (!light_only) {
                Lobe l;
                l.N   = mkv(q[0], q[1], q[2]);
                l.eta = 0.0f;
                bool known = true;
                switch (id) {
                case DIFFUSE_ID: l.type = LOBE_DIFFUSE; break;
                case TRANSLUCENT_ID: l.type = LOBE_TRANSLUCENT; l.N = -l.N; break;
                case REFLECTION_ID: l.type = LOBE_REFLECTION; break;
                case FRESNEL_REFLECTION_ID: l.type = LOBE_REFLECTION; l.eta = q[3]; break;
                case REFRACTION_ID: l.type = LOBE_REFRACTION; l.eta = q[3]; break;
                case TRANSPARENT_ID:
                case MX_TRANSPARENT_ID: l.type = LOBE_TRANSPARENT; break;
#ifdef OSLD_GLOSSY_LOBES
                case OREN_NAYAR_ID:
                    // -> MxOrenNayarDiffuse{N, albedo 1, sigma, no energy compensation}
                    l.type    = LOBE_BSDL_OREN_NAYAR;
                    l.ax      = bsdl_clamp(q[3], 0.0f, 1.0f);
                    l.albedo  = mkv(1.0f);
                    l.refract = 0;
                    lobe_set_bsdl_frame(l, wo);
                    break;
                case MX_SHEEN_ID:
                    // params: N, albedo, roughness, mode (1: Zeltner-Burley LTC sheen, else Conty-Kulla)
                    l.albedo = mkv(q[3], q[4], q[5]);
                    l.type   = LOBE_BSDL_SHEEN;
                    sheen_setup(l, wo, q[6], backfacing, path_roughness, __float_as_int(q[7]), luts);
                    break;
#ifdef OSLD_MX_LOBES
                case MX_CONDUCTOR_ID:
                case MX_DIELECTRIC_ID:
                case MX_GENERALIZED_SCHLICK_ID:
                    // a boundary the medium stack rules out (nested dielectrics) is passed straight
                    // through (shading.cpp:1580-1609)
                    if (id != MX_CONDUCTOR_ID && false_intersection)
                        l.type = LOBE_TRANSPARENT;
                    else
                        mx_from_component(luts, l, id, q, wo, backfacing, path_roughness);
                    break;
                case MX_TRANSLUCENT_ID:
                    // params: N, albedo: a cosine lobe on the far side of the visible normal
                    l.type   = LOBE_MX_TRANSLUCENT;
                    l.albedo = mkv(q[3], q[4], q[5]);
                    lobe_set_bsdl_frame(l, wo);
                    break;
                case MX_SUBSURFACE_ID:
                    // no BSSRDF in testrender: a diffuse lobe weighted by the albedo (shading.cpp:1626-1635)
                    l.type = LOBE_DIFFUSE;
                    cw     = cw * mkv(q[3], q[4], q[5]);
                    break;
#ifdef OSLD_THINLAYER
                case SPI_THINLAYER: {
                    // params: N, T, IOR, roughness, anisotropy, thickness, refl_tint, refr_tint, sigma_t
                    // (ThinLayerLobe::Data registration order; shading.cpp:1668-1674)
                    l.type     = LOBE_SPI_THINLAYER;
                    const V3 Z = bsdl_visible_normal(wo, l.N, l.N);
                    mx_set_frame_zx(l, Z, mkv(q[3], q[4], q[5]));
                    l.thin = thin_setup(luts, dot3(wo, Z), q[6], q[7], q[8], q[9], mkv(q[10], q[11], q[12]),
                                        mkv(q[13], q[14], q[15]), mkv(q[16], q[17], q[18]), path_roughness);
                    break;
                }
#endif
#endif
                case MX_LAYER_ID: {
                    // layer(top, base): the base is attenuated by what the top stack takes
                    // (shading.cpp:1645-1661)
                    const int top = __float_as_int(q[0]), base = __float_as_int(q[1]);
                    V3 op     = evaluate_layer_opacity(pool, top, wo, backfacing, path_roughness, luts);
                    op        = mkv(fminf(fmaxf(op.x, 0.f), 1.f), fminf(fmaxf(op.y, 0.f), 1.f), fminf(fmaxf(op.z, 0.f), 1.f));
                    V3 base_w = weight * (mkv(1.0f) - op);
                    closure   = top;
                    weight    = cw;
                    if (!(base_w.x == 0 && base_w.y == 0 && base_w.z == 0)) {
                        ptr_stack[sp]      = base;
                        weight_stack[sp++] = base_w;
                    }
                    known = false;
                    break;
                }
                case MX_OREN_NAYAR_DIFFUSE_ID:
                case MX_BURLEY_DIFFUSE_ID:
                    // params: N, albedo, roughness [, energy_compensation] (libbsdl Data structs)
                    l.type    = id == MX_BURLEY_DIFFUSE_ID ? LOBE_BSDL_BURLEY : LOBE_BSDL_OREN_NAYAR;
                    l.albedo  = mkv(q[3], q[4], q[5]);
                    l.ax      = bsdl_clamp(q[6], 0.0f, 1.0f);
                    l.refract = (id == MX_OREN_NAYAR_DIFFUSE_ID) ? (__float_as_int(q[7]) != 0) : 0;
                    lobe_set_bsdl_frame(l, wo);
                    break;
                case PHONG_ID:
                    l.type = LOBE_PHONG;
                    l.ax   = q[3];
                    break;
                case WARD_ID:
                    l.type = LOBE_WARD;
                    l.ax   = q[6];
                    l.ay   = q[7];
                    lobe_set_frame(l, mkv(q[3], q[4], q[5]));
                    break;
                case MICROFACET_ID: {
                    // params: dist code, N, U, xalpha, yalpha, eta, refract (shading.cpp:222-231)
                    const int dist = __float_as_int(q[0]);
                    l.type    = LOBE_MICROFACET;
                    l.N       = mkv(q[1], q[2], q[3]);
                    l.ax      = q[7];
                    l.ay      = q[8];
                    l.eta     = q[9];
                    l.refract = __float_as_int(q[10]);
                    l.ggx     = dist == 1;
                    lobe_set_frame(l, mkv(q[4], q[5], q[6]));
                    known = (dist >= 1 && dist <= 3) && l.refract >= 0 && l.refract <= 2;
                    break;
                }
#endif
                default: known = false; break;
                }
                if (known && B.num < OSLD_MAX_LOBES) {
                    B.weights[B.num] = cw;
                    B.lobes[B.num]   = l;
                    ++B.num;
                }
            }
        }
        if (closure == 0 && sp > 0) {
            closure = ptr_stack[--sp];
            weight  = weight_stack[sp];
        }
    }
}

// ---- BVH traversal ---------------------------------------------------------------------
struct Hit {
    float t, u, v;
    unsigned id;
};
OSLD float minf_(float a, float b) { return b < a ? b : a; }
OSLD float maxf_(float a, float b) { return b > a ? b : a; }
// The same slab test with the hardware min / max (one FMNMX each instead of compare + select).
// minf_ / maxf_ differ from fminf / fmaxf only when their FIRST operand is a NaN, and a NaN can only
// come from 0 * inf, i.e. from a ray with a zero (or denormal) direction component; trav_init flags those rays
// (Trav::exact) and they take box_intersect.  The sign of a zero result may differ, which no
// comparison below can see.
OSLD bool box_intersect_finite(V3 org, V3 rdir, float tmax, float4 b0, float4 b1, float* dist)
{
    const float tx1 = (b0.x - org.x) * rdir.x, tx2 = (b0.y - org.x) * rdir.x;
    const float ty1 = (b0.z - org.y) * rdir.y, ty2 = (b0.w - org.y) * rdir.y;
    const float tz1 = (b1.x - org.z) * rdir.z, tz2 = (b1.y - org.z) * rdir.z;
    float tmin      = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fminf(tz1, tz2));
    tmax            = fminf(fminf(fminf(tmax, fmaxf(tx1, tx2)), fmaxf(ty1, ty2)), fmaxf(tz1, tz2));
    *dist           = tmin;
    tmin            = fmaxf(0.0f, tmin);
    return tmin <= tmax;
}
OSLD bool box_intersect(V3 org, V3 rdir, float tmax, float4 b0, float4 b1, float* dist)
{
    // bounds = { b0.x, b0.y, b0.z, b0.w, b1.x, b1.y } = minx,maxx,miny,maxy,minz,maxz
    const float tx1 = (b0.x - org.x) * rdir.x, tx2 = (b0.y - org.x) * rdir.x;
    const float ty1 = (b0.z - org.y) * rdir.y, ty2 = (b0.w - org.y) * rdir.y;
    const float tz1 = (b1.x - org.z) * rdir.z, tz2 = (b1.y - org.z) * rdir.z;
    float tmin      = minf_(tx1, tx2);
    tmax            = minf_(tmax, maxf_(tx1, tx2));
    tmin            = maxf_(tmin, minf_(ty1, ty2));
    tmax            = minf_(tmax, maxf_(ty1, ty2));
    tmin            = maxf_(tmin, minf_(tz1, tz2));
    tmax            = minf_(tmax, maxf_(tz1, tz2));
    *dist           = tmin;
    tmin            = maxf_(0.0f, tmin);
    return tmin <= tmax;
}
OSLD float xorf(float a, unsigned b) { return __int_as_float((int)(fbits(a) ^ b)); }

// Traversal stack of a thread: OSLD_BVH_STACK entries (the host sizes it from the depth of the
// scene's BVH when the module is compiled) of OSLD_STK_WORDS words, word j of the thread at
// stk[j * stride] - a slice of the CTA's shared memory interleaved by thread (conflict free),
// where the reference keeps three 64-entry arrays on the CPU stack (bvh.cpp:271-273).
// An entry carries the pending node's (child, nprims) words, already fetched with its bounds
// when the parent was visited, so a pop costs no further node load; packed into one word when
// the host found child < 2^26 and nprims < 64 for every node.
#ifndef OSLD_BVH_STACK
#define OSLD_BVH_STACK 64
#endif
#ifdef OSLD_BVH_UNPACKED
#define OSLD_STK_WORDS 3
#else
#define OSLD_STK_WORDS 2
#endif
struct Trav {
    V3 org, rdir;
    float shx, shy, shz;
    float tmax0;     // any-hit rays: a hit strictly closer than this ends the walk (occluded)
    Hit hit;
    unsigned skip1, skip2;
    int kx, ky, kz, sp, anyhit;
    bool exact;      // an infinite reciprocal direction: 0 * inf NaNs possible, take the compare-and-select slab test
    unsigned lchild, lnprims;   // rt_trace: the part of the current leaf not tested yet (lnprims == 0: none)
    unsigned* stk;
    int stride;
};
OSLD void stk_put(const Trav& T, int i, unsigned child, unsigned nprims, float dist)
{
    unsigned* e = T.stk + (size_t)(OSLD_STK_WORDS * i) * T.stride;
#ifdef OSLD_BVH_UNPACKED
    e[0]            = child;
    e[T.stride]     = fbits(dist);
    e[2 * T.stride] = nprims;
#else
    e[0]        = child | (nprims << 26);
    e[T.stride] = fbits(dist);
#endif
}
OSLD float stk_dist(const Trav& T, int i)
{
    return __int_as_float((int)T.stk[(size_t)(OSLD_STK_WORDS * i + 1) * T.stride]);
}
OSLD void stk_get(const Trav& T, int i, unsigned& child, unsigned& nprims)
{
    const unsigned* e = T.stk + (size_t)(OSLD_STK_WORDS * i) * T.stride;
#ifdef OSLD_BVH_UNPACKED
    child  = e[0];
    nprims = e[2 * T.stride];
#else
    child  = e[0] & 0x3ffffffu;
    nprims = e[0] >> 26;
#endif
}
OSLD void trav_init(const RenderScene& S, Trav& T, unsigned* stk, int stride, V3 org, V3 dir, float tmax,
                    unsigned skip1, unsigned skip2, int anyhit)
{
    T.stk = stk; T.stride = stride;
    T.org = org;
    T.hit.t = tmax; T.hit.u = T.hit.v = 0.0f; T.hit.id = 0;
    T.tmax0 = tmax; T.skip1 = skip1; T.skip2 = skip2; T.anyhit = anyhit;
    T.lchild = 0u; T.lnprims = 0u;
    const float4 r1 = __ldg(S.bvh_nodes + 1);
    stk_put(T, 0, fbits(r1.z), fbits(r1.w), tmax);
    T.sp   = 1;
    T.rdir = mkv(1 / dir.x, 1 / dir.y, 1 / dir.z);
    T.exact = fabsf(T.rdir.x) == OSLD_INF || fabsf(T.rdir.y) == OSLD_INF || fabsf(T.rdir.z) == OSLD_INF;
    int kz = 0;
    if (fabsf(dir.y) > fabsf(vcomp(dir, kz)))
        kz = 1;
    if (fabsf(dir.z) > fabsf(vcomp(dir, kz)))
        kz = 2;
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    T.kx = kx; T.ky = ky; T.kz = kz;
    T.shx = vcomp(dir, kx) / vcomp(dir, kz);
    T.shy = vcomp(dir, ky) / vcomp(dir, kz);
    T.shz = vcomp(T.rdir, kz);
}
// one triangle of a leaf against the ray (watertight test, bvh.cpp:203-263)
OSLD void trav_tri(const RenderScene& S, const Trav& T, unsigned at, bool z0, bool z1, Hit& result)
{
    const V3 org = T.org;
    const float shx = T.shx, shy = T.shy, shz = T.shz;
    const float4* lt = S.leaf_tris + 3 * (size_t)at;
    const float4 ta = __ldg(lt), tb = __ldg(lt + 1), tc = __ldg(lt + 2);
    const unsigned id = fbits(ta.w);
    const V3 A = vperm_kz(xyz(ta) - org, z0, z1);
    const V3 B = vperm_kz(xyz(tb) - org, z0, z1);
    const V3 C = vperm_kz(xyz(tc) - org, z0, z1);
    const float Ax = A.x - shx * A.z, Ay = A.y - shy * A.z;
    const float Bx = B.x - shx * B.z, By = B.y - shy * B.z;
    const float Cx = C.x - shx * C.z, Cy = C.y - shy * C.z;
    const float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
    if ((U < 0 || V < 0 || W < 0) && (U > 0 || V > 0 || W > 0))
        return;
    const float det = U + V + W;
    if (det == 0)
        return;
    const float Tt      = shz * (U * A.z + V * B.z + W * C.z);
    const unsigned mask = fbits(det) & 0x80000000u;
    if (xorf(Tt, mask) < 0)
        return;
    if (xorf(Tt, mask) > result.t * xorf(det, mask))
        return;
    if (id == T.skip1 || id == T.skip2)
        return;
    const float rcpDet = 1 / det;
    result.t  = Tt * rcpDet;
    result.u  = V * rcpDet;
    result.v  = W * rcpDet;
    result.id = id;
}
// the triangles of one leaf, in leaf order
OSLD void trav_leaf(const RenderScene& S, const Trav& T, unsigned child, unsigned nprims, Hit& result)
{
    const bool z0 = T.kz == 0, z1 = T.kz == 1;
    for (unsigned i = 0; i < nprims; i++)
        trav_tri(S, T, child + i, z0, z1, result);
}
// one inner node: both children's boxes, far child pushed first (bvh.cpp:300-340)
OSLD void trav_visit(const RenderScene& S, const Trav& T, unsigned child, float tnear, int& sp)
{
    // the two children are adjacent: 64 contiguous bytes
    const float4* cn = S.bvh_nodes + 2 * (size_t)child;
    const float4 a0 = __ldg(cn), a1 = __ldg(cn + 1), b0 = __ldg(cn + 2), b1 = __ldg(cn + 3);
    float d1 = 0, d2 = 0;
    bool h1, h2;
    if (T.exact) {
        h1 = box_intersect(T.org, T.rdir, tnear, a0, a1, &d1);
        h2 = box_intersect(T.org, T.rdir, tnear, b0, b1, &d2);
    } else {
        h1 = box_intersect_finite(T.org, T.rdir, tnear, a0, a1, &d1);
        h2 = box_intersect_finite(T.org, T.rdir, tnear, b0, b1, &d2);
    }
    unsigned k1 = fbits(a1.z), n1 = fbits(a1.w), k2 = fbits(b1.z), n2 = fbits(b1.w);
    if (d1 > d2) {
        bool th = h1; h1 = h2; h2 = th;
        float td = d1; d1 = d2; d2 = td;
        unsigned tk = k1; k1 = k2; k2 = tk;
        unsigned tn = n1; n1 = n2; n2 = tn;
    }
    stk_put(T, sp, k2, n2, d2);
    sp += h2 ? 1 : 0;
    stk_put(T, sp, k1, n1, d1);
    sp += h1 ? 1 : 0;
}
// Scene::intersect (bvh.cpp:265-356), resumable, one STEP per call: walk at most `max_nodes`
// inner nodes; if that reaches a leaf, test its triangles.  Returns true when the ray is finished.
// "while-while" traversal (Aila & Laine) with a bounded walk: every lane first walks inner
// nodes until it holds a leaf (or has spent its node budget), then the lanes holding a leaf
// test triangles together - a lane that finds its leaf early waits for at most max_nodes
// visits of the others.  Per ray the visiting order is exactly the reference's pop / test /
// push-far-then-near sequence; only the lock-step grouping of the lanes changes, so the hit
// (and any tie between coplanar triangles) is the reference's.
OSLD bool trav_step(const RenderScene& S, Trav& T, int max_nodes)
{
    int sp = T.sp;
    Hit result = T.hit;
    unsigned child = 0, nprims = 0;
    while (sp != 0 && max_nodes > 0) {
        --sp;
        if (result.t < stk_dist(T, sp))
            continue;
        stk_get(T, sp, child, nprims);
        if (nprims)
            break;
        --max_nodes;
        trav_visit(S, T, child, result.t, sp);
    }
    bool finished = false;
    if (nprims) {
        trav_leaf(S, T, child, nprims, result);
        // a shadow ray only asks "is anything strictly closer than tmax0": the closest-hit walk
        // can only move t further down from here, so the answer is already known
        finished = T.anyhit && result.t < T.tmax0;
    } else
        finished = sp == 0;
    T.sp  = sp;
    T.hit = result;
    return finished;
}
// trav_step for a whole warp (rt_trace): every lane calls it, `active` says whether the lane holds a
// ray.  Per ray the visiting order is trav_step's (the reference's); what changes is how the lanes are
// grouped.  Each iteration the warp votes: if at least as many lanes want to visit an inner node as
// hold an untested leaf triangle, the walkers visit one node, otherwise the leaf holders test one
// triangle.  A lane moves between the two roles as its ray demands, the phase with more ready lanes
// always runs, and leaves of different sizes no longer wait for each other.  (With trav_step's fixed
// walk-then-leaf phases ncu's source page showed 7.7 of 32 lanes in the node visit and 14 in the
// triangle test: profiles/ncu_r02_rt_trace_sass.txt.)  The call returns when every ray of the warp is
// finished, after `budget` iterations, or - while the queue can still refill lanes - as soon as
// `min_busy` or fewer lanes are busy.
OSLD bool trav_step_warp(const RenderScene& S, Trav& T, bool active, int budget, int min_busy)
{
    int sp = T.sp;
    Hit result = T.hit;
    unsigned child = T.lchild, nprims = active ? T.lnprims : 0u;
    bool done = !active || (nprims == 0u && sp == 0);
    const bool z0 = T.kz == 0, z1 = T.kz == 1;
    for (int it = 0; it < budget; ++it) {
        const bool inleaf  = nprims != 0u;
        const bool walking = !done && !inleaf;
        const int nw = __popc(__ballot_sync(0xffffffffu, walking)), nl = __popc(__ballot_sync(0xffffffffu, inleaf));
        if (nw + nl <= min_busy)
            break;
        if (nw >= nl) {
            if (walking) {
                // pop to the next entry the current hit has not culled
                bool got = false;
                while (sp != 0) {
                    --sp;
                    if (result.t < stk_dist(T, sp))
                        continue;
                    stk_get(T, sp, child, nprims);
                    got = true;
                    break;
                }
                if (!got)
                    done = true;
                else if (nprims == 0u) {
                    trav_visit(S, T, child, result.t, sp);
                    done = sp == 0;
                }
            }
        } else if (inleaf) {
            // two triangles per vote (the reference's leaves hold 8 on average)
            trav_tri(S, T, child, z0, z1, result);
            if (nprims > 1u)
                trav_tri(S, T, child + 1u, z0, z1, result);
            const unsigned k = nprims > 1u ? 2u : 1u;
            child += k;
            nprims -= k;
            // a shadow ray only asks "is anything strictly closer than tmax0": answered by the first hit
            if (T.anyhit && result.t < T.tmax0) {
                nprims = 0u;
                done   = true;
            } else if (nprims == 0u)
                done = sp == 0;
        }
    }
    T.sp      = sp;
    T.hit     = result;
    T.lchild  = child;
    T.lnprims = nprims;
    return active && done;
}
OSLD Hit scene_intersect(const RenderScene& S, unsigned* stk, int stride, V3 org, V3 dir, float tmax,
                         unsigned skip1, unsigned skip2, int anyhit = 0)
{
    Trav T;
    trav_init(S, T, stk, stride, org, dir, tmax, skip1, skip2, anyhit);
    while (!trav_step(S, T, 1 << 30)) {}
    return T.hit;
}

// ---- scene queries -----------------------------------------------------------------------
struct LightSample {
    V3 dir;
    float dist, pdf, u, v;
};
OSLD void tri_verts(const RenderScene& S, int id, V3& va, V3& vb, V3& vc)
{
    va = ld3(S.verts, __ldg(S.triangles + 3 * id));
    vb = ld3(S.verts, __ldg(S.triangles + 3 * id + 1));
    vc = ld3(S.verts, __ldg(S.triangles + 3 * id + 2));
}
OSLD LightSample scene_sample(const RenderScene& S, int primID, V3 x, float xi, float yi)
{
    if (yi > xi) {
        xi *= 0.5f;
        yi -= xi;
    } else {
        yi *= 0.5f;
        xi -= yi;
    }
    V3 va, vb, vc;
    tri_verts(S, primID, va, vb, vc);
    const V3 n = cross3(va - vb, va - vc);
    V3 l       = ((1 - xi - yi) * va + xi * vb + yi * vc) - x;
    float d2   = len2(l);
    V3 dir     = vnormalized(l);
    LightSample s;
    s.dir  = dir;
    s.dist = sqrtf(d2);
    s.pdf  = d2 / (0.5f * fabsf(dot3(dir, n)));
    s.u    = xi;
    s.v    = yi;
    return s;
}
OSLD float scene_shapepdf(const RenderScene& S, int primID, V3 x, V3 p)
{
    V3 va, vb, vc;
    tri_verts(S, primID, va, vb, vc);
    const V3 n = cross3(va - vb, va - vc);
    V3 l       = p - x;
    float d2   = len2(l);
    V3 dir     = vnormalized(l);
    return d2 / (0.5f * fabsf(dot3(dir, n)));
}

// globals_from_hit (simpleraytracer.cpp:889-932) incl. Scene::normal/project/uv
OSLD void globals_from_hit(const RenderScene& S, SG& sg, const Ray& r, float t, int id, float u, float v)
{
    // r.dual_direction()
    sg.I = r.direction;
    ortho(r.direction, sg.I_dx, sg.I_dy);
    sg.I_dx = sg.I_dx * r.spread;
    sg.I_dy = sg.I_dy * r.spread;
    // r.point(Dual t)
    const float rr = r.radius + r.spread * t;
    V3 P           = ray_point(r, t), Pdx, Pdy;
    ortho(r.direction, Pdx, Pdy);
    Pdx = Pdx * rr;
    Pdy = Pdy * rr;
    V3 va, vb, vc;
    tri_verts(S, id, va, vb, vc);
    V3 Ng = vnormalized(cross3(va - vb, va - vc));
    V3 N  = Ng;
    if (__ldg(S.n_triangles + 3 * id) >= 0) {
        const V3 na = ld3(S.normals, __ldg(S.n_triangles + 3 * id)), nb = ld3(S.normals, __ldg(S.n_triangles + 3 * id + 1)),
                 nc = ld3(S.normals, __ldg(S.n_triangles + 3 * id + 2));
        N = vnormalized((1 - u - v) * na + u * nb + v * nc);
    }
    // project
    {
        V3 nI      = vnormalized(sg.I);
        float cosI = dot3(-nI, N);
        if (fabsf(cosI) > 1e-3f) {
            float deltaX = dot3(Pdx, N) / cosI;
            float deltaY = dot3(Pdy, N) / cosI;
            Pdx = Pdx + nI * deltaX;
            Pdy = Pdy + nI * deltaY;
        }
    }
    sg.P    = P;
    sg.P_dx = Pdx;
    sg.P_dy = Pdy;
    // uv
    sg.dPdu = mkv(0.0f);
    sg.dPdv = mkv(0.0f);
    sg.u = sg.u_dx = sg.u_dy = sg.v = sg.v_dx = sg.v_dy = 0.0f;
    if (__ldg(S.uv_triangles + 3 * id) >= 0) {
        const int ia = __ldg(S.uv_triangles + 3 * id), ib = __ldg(S.uv_triangles + 3 * id + 1), ic = __ldg(S.uv_triangles + 3 * id + 2);
        const float tax = __ldg(S.uvs + 2 * ia), tay = __ldg(S.uvs + 2 * ia + 1);
        const float tbx = __ldg(S.uvs + 2 * ib), tby = __ldg(S.uvs + 2 * ib + 1);
        const float tcx = __ldg(S.uvs + 2 * ic), tcy = __ldg(S.uvs + 2 * ic + 1);
        const float dt02x = tax - tcx, dt02y = tay - tcy, dt12x = tbx - tcx, dt12y = tby - tcy;
        const V3 dp02 = va - vc, dp12 = vb - vc;
        const float det = dt02x * dt12y - dt02y * dt12x;
        if (det != 0) {
            float invdet = 1 / det;
            sg.dPdu      = (dt12y * dp02 - dt02y * dp12) * invdet;
            sg.dPdv      = (-dt12x * dp02 + dt02x * dp12) * invdet;
        }
        V3 La = cross3(N, vc - vb);
        La    = vdiv(La, dot3(va - vb, La));
        V3 Lb = cross3(N, va - vc);
        Lb    = vdiv(Lb, dot3(vb - vc, Lb));
        V3 Lc = cross3(N, vb - va);
        Lc    = vdiv(Lc, dot3(vc - va, Lc));
        float ax = dot3(La, Pdx), bx = dot3(Lb, Pdx), cx = dot3(Lc, Pdx);
        float ay = dot3(La, Pdy), by = dot3(Lb, Pdy), cy = dot3(Lc, Pdy);
        float w0 = 1 - u - v;
        sg.u    = w0 * tax + u * tbx + v * tcx;
        sg.v    = w0 * tay + u * tby + v * tcy;
        sg.u_dx = ax * tax + bx * tbx + cx * tcx;
        sg.v_dx = ax * tay + bx * tby + cx * tcy;
        sg.u_dy = ay * tax + by * tbx + cy * tcx;
        sg.v_dy = ay * tay + by * tby + cy * tcy;
    }
    sg.surfacearea = __ldg(S.mesh_surfacearea + __ldg(S.meshids + id));
    sg.backfacing  = dot3(Ng, sg.I) > 0 ? 1 : 0;
    if (sg.backfacing) {
        N  = -N;
        Ng = -Ng;
    }
    sg.N              = N;
    sg.Ng             = Ng;
    sg.raytype        = r.raytype;
    sg.flipHandedness = dot3(cross3(Pdx, Pdy), N) < 0 ? 1 : 0;
    sg.dPdz = mkv(0.0f);
    sg.time = sg.dtime = 0.0f;
    sg.dPdtime = mkv(0.0f);
    sg.Ps = sg.Ps_dx = sg.Ps_dy = mkv(0.0f);
    sg.shadeindex = 0;
}

OSLD Ray camera_ray(const RenderScene& S, float x, float y)
{
    V3 cx = mkv(S.cx[0], S.cx[1], S.cx[2]), cy = mkv(S.cy[0], S.cy[1], S.cy[2]), dir = mkv(S.dir[0], S.dir[1], S.dir[2]);
    const V3 v        = vnormalized(cx * (x * S.invw - 0.5f) + cy * (0.5f - y * S.invh) + dir);
    const float cos_a = dot3(dir, v);
    Ray r;
    r.origin    = mkv(S.eye[0], S.eye[1], S.eye[2]);
    r.direction = v;
    r.radius    = 0.0f;
    r.spread    = sqrtf(S.invw * S.invh * imath_length(cx) * imath_length(cy) * cos_a) * cos_a;
    r.roughness = 0.0f;
    r.raytype   = RAY_CAMERA;
    return r;
}

// ---- background (background.h, simpleraytracer.cpp:937-954, shading.cpp:1709-1746) ---------
// Compiled in only for scenes with a <Background> (OSLD_HAS_BACKGROUND from the generator).
#ifdef OSLD_HAS_BACKGROUND
// The reference returns the weight held when the tree walk ends (last visited branch).
OSLD V3 process_background_closure(const ClosurePool& pool, int closure)
{
    if (!closure)
        return mkv(0.0f);
    int ptr_stack[OSLD_CLOSURE_STACK];
    V3 weight_stack[OSLD_CLOSURE_STACK];
    int sp    = 0;
    V3 weight = mkv(1.0f);
    while (closure) {
        int id = pool.id(closure);
        if (id == CL_MUL) {
            weight  = weight * pool.weight(closure);
            closure = __float_as_int(pool.w[closure + 4]);
        } else if (id == CL_ADD) {
            ptr_stack[sp]      = __float_as_int(pool.w[closure + 2]);
            weight_stack[sp++] = weight;
            closure            = __float_as_int(pool.w[closure + 1]);
        } else {
            if (id == BACKGROUND_ID)
                weight = weight * pool.weight(closure);
            closure = 0;
        }
        if (closure == 0 && sp > 0) {
            closure = ptr_stack[--sp];
            weight  = weight_stack[sp];
        }
    }
    return weight;
}
OSLD V3 eval_background(const RenderScene& S, V3 dir, V3 ddx, V3 ddy, int bounce, ClosurePool& pool)
{
    SG sg;
    memset(&sg, 0, sizeof(SG));
    sg.I    = dir;
    sg.I_dx = ddx;
    sg.I_dy = ddy;
    if (bounce >= 0)
        sg.raytype = bounce > 0 ? RAY_DIFFUSE : RAY_CAMERA;
    pool.reset();
    sg.pool = &pool;
    sg.Ci   = 0;
    osl_execute_shader(S.background_shader, sg);
    return process_background_closure(pool, sg.Ci);
}
// texel (x, y) of the table -> direction with derivatives (Background::map)
OSLD void bg_map(const RenderScene& S, float x, float y, V3& d, V3& ddx, V3& ddy)
{
    Df u     = mkd(x, 1.0f, 0.0f) * S.bg_invres;
    Df v     = mkd(y, 0.0f, 1.0f) * S.bg_invres;
    Df theta = u * (float)(2 * OSLD_PI);
    float s, c;
    fast_sincos(theta.val, &s, &c);
    Df st = chain(theta, s, c), ct = chain(theta, c, -s);
    Df cos_phi = 1.0f - 2.0f * v;
    Df sin_phi = d_sqrt(1.0f - cos_phi * cos_phi);
    Df X = sin_phi * ct, Y = sin_phi * st;
    d   = mkv(X.val, Y.val, cos_phi.val);
    ddx = mkv(X.dx, Y.dx, cos_phi.dx);
    ddy = mkv(X.dy, Y.dy, cos_phi.dy);
}
OSLD V3 bg_eval(const RenderScene& S, V3 dir, float& pdf)
{
    const int res = S.bg_res;
    float u = fast_atan2(dir.y, dir.x) * (float)(0.31830988618379067154 * 0.5f);
    if (u < 0)
        u++;
    float v = (1 - dir.z) * 0.5f;
    int x   = (int)(u * res);
    x       = x < 0 ? 0 : (x >= res ? res - 1 : x);
    int y   = (int)(v * res);
    y       = y < 0 ? 0 : (y >= res ? res - 1 : y);
    int i   = y * res + x;
    float row_pdf = S.bg_rows[y] - (y > 0 ? S.bg_rows[y - 1] : 0.0f);
    float col_pdf = S.bg_cols[i] - (x > 0 ? S.bg_cols[i - 1] : 0.0f);
    pdf           = fmaxf(0.0f, row_pdf * col_pdf * S.bg_invjacobian);
    return mkv(S.bg_values[3 * i], S.bg_values[3 * i + 1], S.bg_values[3 * i + 2]);
}
OSLD float bg_sample_cdf(const float* data, int n, float x, int* idx, float* pdf)
{
    const float* first = data;
    int len            = n;
    while (len != 0) {  // upper_bound
        int l2         = len / 2;
        const float* m = first + l2;
        if (x < *m)
            len = l2;
        else {
            first = m + 1;
            len -= l2 + 1;
        }
    }
    int i = (int)(first - data);
    *idx  = i;
    float scaled;
    if (i == 0) {
        *pdf   = data[0];
        scaled = x / data[0];
    } else {
        *pdf   = data[i] - data[i - 1];
        scaled = (x - data[i - 1]) / (data[i] - data[i - 1]);
    }
    return fminf(scaled, 0.99999994f);
}
OSLD V3 bg_sample(const RenderScene& S, float rx, float ry, V3& dir, float& pdf)
{
    const int res = S.bg_res;
    float row_pdf, col_pdf;
    int x, y;
    ry = bg_sample_cdf(S.bg_rows, res, ry, &y, &row_pdf);
    rx = bg_sample_cdf(S.bg_cols + (size_t)y * res, res, rx, &x, &col_pdf);
    V3 ddx, ddy;
    bg_map(S, (float)x + rx, (float)y + ry, dir, ddx, ddy);
    pdf   = fmaxf(0.0f, row_pdf * col_pdf * S.bg_invjacobian);
    int i = y * res + x;
    return mkv(S.bg_values[3 * i], S.bg_values[3 * i + 1], S.bg_values[3 * i + 2]);
}
#endif  // OSLD_HAS_BACKGROUND


// warp-aggregated append to a queue
OSLD void queue_push(int* queue, int* counter, int value, bool pred)
{
    unsigned m = __ballot_sync(__activemask(), pred);
    if (!pred)
        return;
    int lane   = threadIdx.x & 31;
    int leader = __ffs(m) - 1;
    int base   = 0;
    if (lane == leader)
        base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    queue[base + __popc(m & ((1u << lane) - 1u))] = value;
}

// The closure arena of a thread: a slice of the CTA's shared memory when the generator could
// bound it (OSLD_POOL_SMEM: <= 64 words per thread), else a private array.
#ifdef OSLD_POOL_SMEM
#define OSLD_POOL_DECL(pool, smem_words)                                   \
    ClosurePool pool;                                                      \
    pool.bind(reinterpret_cast<float*>(smem_words) + threadIdx.x, blockDim.x)
#else
#define OSLD_POOL_DECL(pool, smem_words)     \
    float pool##_store_[OSLD_POOL_STORE];    \
    ClosurePool pool;                        \
    pool.bind(pool##_store_, 1)
#endif

}  // namespace osld

using namespace osld;

// ---- kernels -------------------------------------------------------------------------------
// Camera::lookat + resolution + finalize (raytracer.h:95-126), evaluated once on the
// device so the OIIO fast_tan restatement exists in exactly one place.
// out[0..2] = normalized dir, out[3..5] = cx, out[6..8] = cy
extern "C" __global__ void rt_camera(const __grid_constant__ RenderLaunch L, float* out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    const RenderScene& S = L.S;
    V3 dir   = vnormalized(mkv(S.dir[0], S.dir[1], S.dir[2]));
    V3 up    = mkv(S.up[0], S.up[1], S.up[2]);
    float k  = fast_tan(S.fov * (float)(OSLD_PI / 360));
    V3 right = vnormalized(cross3(dir, up));
    V3 cx    = right * ((float)S.xres * k / (float)S.yres);
    V3 cy    = vnormalized(cross3(cx, dir)) * k;
    out[0] = dir.x; out[1] = dir.y; out[2] = dir.z;
    out[3] = cx.x; out[4] = cx.y; out[5] = cx.z;
    out[6] = cy.x; out[7] = cy.y; out[8] = cy.z;
}

// Background::prepare (background.h:52-95) in four passes that keep the reference's
// floating-point summation order: (1) shade every texel, (2) one thread per row runs the
// sequential column prefix sum and normalises the row, (3) one thread runs the row prefix
// sum, (4) every texel is divided by its pdf.
#ifdef OSLD_HAS_BACKGROUND
extern "C" __global__ void __launch_bounds__(128) rt_bg_eval(const __grid_constant__ RenderLaunch L)
{
    extern __shared__ unsigned smem_[];
    OSLD_POOL_DECL(pool, smem_);
    const RenderScene& S = L.S;
    const int res = S.bg_res, n = res * res;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int y = i / res, x = i - y * res;
        V3 d, ddx, ddy;
        bg_map(S, (float)x + 0.5f, (float)y + 0.5f, d, ddx, ddy);
        V3 c = eval_background(S, d, ddx, ddy, -1, pool);
        S.bg_values[3 * i] = c.x; S.bg_values[3 * i + 1] = c.y; S.bg_values[3 * i + 2] = c.z;
    }
}
extern "C" __global__ void __launch_bounds__(128) rt_bg_rows(const __grid_constant__ RenderLaunch L)
{
    const RenderScene& S = L.S;
    const int res = S.bg_res;
    for (int y = blockIdx.x * blockDim.x + threadIdx.x; y < res; y += gridDim.x * blockDim.x) {
        float* cols = S.bg_cols + (size_t)y * res;
        const float* vals = S.bg_values + (size_t)3 * y * res;
        float run = 0.0f;
        for (int x = 0; x < res; ++x) {
            float m = fmaxf(fmaxf(vals[3 * x], vals[3 * x + 1]), vals[3 * x + 2]);
            run     = m + ((x > 0) ? run : 0.0f);
            cols[x] = run;
        }
        S.bg_rows[y] = run;  // row total; the prefix sum over rows runs in rt_bg_finish
        if (run > 0)
            for (int x = 0; x < res; ++x)
                cols[x] /= run;
    }
}
extern "C" __global__ void rt_bg_finish(const __grid_constant__ RenderLaunch L)
{
    const RenderScene& S = L.S;
    if (blockIdx.x != 0 || threadIdx.x != 0)
        return;
    const int res = S.bg_res;
    for (int y = 1; y < res; ++y)
        S.bg_rows[y] = S.bg_rows[y] + S.bg_rows[y - 1];
    for (int y = 0; y < res; ++y)
        S.bg_rows[y] /= S.bg_rows[res - 1];
}
extern "C" __global__ void __launch_bounds__(256) rt_bg_scale(const __grid_constant__ RenderLaunch L)
{
    const RenderScene& S = L.S;
    const int res = S.bg_res, n = res * res;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int y = i / res, x = i - y * res;
        float row_pdf = S.bg_rows[y] - (y > 0 ? S.bg_rows[y - 1] : 0.0f);
        float col_pdf = S.bg_cols[i] - (x > 0 ? S.bg_cols[i - 1] : 0.0f);
        float dv      = row_pdf * col_pdf * S.bg_invjacobian;
        S.bg_values[3 * i] /= dv; S.bg_values[3 * i + 1] /= dv; S.bg_values[3 * i + 2] /= dv;
    }
}
#endif  // OSLD_HAS_BACKGROUND

#ifdef OSLD_HAS_MEDIA
// ---- participating media -------------------------------------------------------------------------
// MediumParams / MediumStack (shading.h:438-725) per path slot, OSLD_MEDIUM_WORDS words:
//   w0 = depth | pool_size << 8      w1,w2 = mediums[8] (pool indices by descending priority, one byte each)
//   w3,w4 = entry_order[8] (LIFO)    w8 + 8 p .. = pool[p] = sigma_t.rgb, sigma_s.rgb, g, priority
// current_params / cdf / overlapping_medium_indices are functions of the stack and are recomputed
// where integrate() needs them (same additions in the same order, hence the same floats).
// refraction_ior is carried by the reference but never read by a lobe, so it has no slot here.
#define OSLD_MEDIUM_WORDS 72
struct MediumData {
    V3 sigma_t, sigma_s;
    float g;
    int priority;
};
OSLD MediumData medium_vacuum()
{
    MediumData m;
    m.sigma_t = m.sigma_s = mkv(0.0f);
    m.g        = 0.0f;
    m.priority = 0;
    return m;
}
OSLD bool medium_is_vacuum(const MediumData& m) { return m.sigma_t.x <= 0.0f && m.sigma_t.y <= 0.0f && m.sigma_t.z <= 0.0f; }
struct MediumHead {
    int depth, pool_size;
    unsigned long long mediums, order;
};
OSLD int mbyte(unsigned long long v, int i) { return (int)((v >> (8 * i)) & 0xffull); }
OSLD MediumHead medium_head(const float* med)
{
    const float4 h = *reinterpret_cast<const float4*>(med);
    MediumHead H;
    const int w0 = __float_as_int(h.x);
    H.depth      = w0 & 0xff;
    H.pool_size  = (w0 >> 8) & 0xff;
    H.mediums    = (unsigned long long)(unsigned)__float_as_int(h.y)
                | ((unsigned long long)(unsigned)__float_as_int(h.z) << 32);
    H.order = (unsigned long long)(unsigned)__float_as_int(h.w)
              | ((unsigned long long)(unsigned)__float_as_int(med[4]) << 32);
    return H;
}
OSLD void medium_store_head(float* med, const MediumHead& H)
{
    *reinterpret_cast<float4*>(med) = mki4(H.depth | (H.pool_size << 8), (int)(unsigned)H.mediums,
                                           (int)(unsigned)(H.mediums >> 32), (int)(unsigned)H.order);
    med[4] = __int_as_float((int)(unsigned)(H.order >> 32));
}
OSLD MediumData medium_entry(const float* med, int p)
{
    const float4* e = reinterpret_cast<const float4*>(med + 8 + 8 * p);
    const float4 a = e[0], b = e[1];
    MediumData m;
    m.sigma_t  = mkv(a.x, a.y, a.z);
    m.sigma_s  = mkv(a.w, b.x, b.y);
    m.g        = b.z;
    m.priority = __float_as_int(b.w);
    return m;
}
OSLD int medium_entry_priority(const float* med, int p) { return __float_as_int(med[8 + 8 * p + 7]); }
// MediumStack::add_medium / pop_medium (shading.h:615-676)
OSLD void medium_add(float* med, const MediumData& np)
{
    MediumHead H = medium_head(med);
    if (H.depth >= 8 || H.pool_size >= 8)
        return;
    const int p = H.pool_size++;
    float4* e   = reinterpret_cast<float4*>(med + 8 + 8 * p);
    e[0]        = make_float4(np.sigma_t.x, np.sigma_t.y, np.sigma_t.z, np.sigma_s.x);
    e[1]        = make_float4(np.sigma_s.y, np.sigma_s.z, np.g, __int_as_float(np.priority));
    int insert_pos = H.depth;
    for (int i = 0; i < H.depth; ++i)
        if (np.priority > medium_entry_priority(med, mbyte(H.mediums, i))) {
            insert_pos = i;
            break;
        }
    const unsigned long long lowmask = insert_pos ? (~0ull >> (64 - 8 * insert_pos)) : 0ull;
    const unsigned long long high    = insert_pos < 7 ? ((H.mediums & ~lowmask) << 8) : 0ull;
    H.mediums = (H.mediums & lowmask) | ((unsigned long long)p << (8 * insert_pos)) | high;
    H.order   = (H.order & ~(0xffull << (8 * H.depth))) | ((unsigned long long)p << (8 * H.depth));
    H.depth++;
    medium_store_head(med, H);
}
OSLD void medium_pop(float* med)
{
    MediumHead H = medium_head(med);
    if (H.depth <= 0)
        return;
    H.depth--;
    const int p      = mbyte(H.order, H.depth);
    int sorted_index = -1;
    for (int i = 0; i <= H.depth; ++i)
        if (mbyte(H.mediums, i) == p) {
            sorted_index = i;
            break;
        }
    if (sorted_index >= 0) {
        const unsigned long long lowmask = sorted_index ? (~0ull >> (64 - 8 * sorted_index)) : 0ull;
        H.mediums = (H.mediums & lowmask) | ((H.mediums >> 8) & ~lowmask);
        if (p == H.pool_size - 1)
            H.pool_size--;
    }
    medium_store_head(med, H);
}
// MediumStack::false_intersection_with (shading.h:678-695)
OSLD bool medium_false_intersection(int entrant_priority, bool has_current, int current_priority)
{
    if (!has_current)
        return false;
    if (entrant_priority == 0 && current_priority == 0)
        return false;
    if (entrant_priority == current_priority)
        return true;
    return entrant_priority > current_priority;
}
// OIIO::fast_sinpi / fast_cospi (fmath.h): BSDLConfig::Fast::sinpif / cospif of the phase sampler
OSLD float fast_sinpi(float x)
{
    const float z = x - ((x + 25165824.0f) - 25165824.0f);
    const float y = z - z * fabsf(z);
    const float Q = 3.10396624f;
    const float P = 3.584135056f;
    return y * (Q + P * fabsf(y));
}
OSLD float fast_cospi(float x) { return fast_sinpi(x + 0.5f); }
// expf / logf of the host's libm: evaluated in double and rounded once, which is what a
// correctly rounded libm returns (glibc's are within 0.502 / 0.818 ulp: a rare last-bit difference)
OSLD float libm_expf(float x) { return (float)exp((double)x); }
OSLD float libm_logf(float x) { return (float)log((double)x); }
OSLD V3 medium_transmittance(V3 sigma_t, float distance)
{
    return mkv(libm_expf(-sigma_t.x * distance), libm_expf(-sigma_t.y * distance), libm_expf(-sigma_t.z * distance));
}
// spi::VolumeLobe{g, g, blend 0} (BSDL/SPI/bsdf_volume_impl.h) in the frame around -wo:
// MediumParams::sample_phase_func (shading.cpp:1185-1196)
OSLD float hg_phase(float costheta, float g)
{
    if (g == 0)
        return 0.25f * (float)(1.0 / OSLD_PI);
    const float num = 0.25f * (float)(1.0 / OSLD_PI) * (1 - g * g);
    const float den = 1 + g * g + 2.0f * g * costheta;
    return num / sqrtf(den * den * den);
}
OSLD BSample medium_sample_phase(const MediumData& m, V3 wo, float rx, float ry)
{
    BSample s;
    if (medium_is_vacuum(m)) {
        s.wi = mkv(1.0f); s.weight = mkv(1.0f); s.pdf = 0.0f; s.roughness = 0.0f;
        return s;
    }
    const float g1 = fminf(fmaxf(m.g, -0.99f), 0.99f), blend = 0.0f;
    float g, x;
    if (rx < blend) {
        g = g1;
        x = rx / blend;
    } else {
        g = g1;
        x = (rx - blend) / (1 - blend);
    }
    float cosTheta;
    if (fabsf(g) < 1e-3f)
        cosTheta = 1 - 2 * x;
    else {
        float k  = (1 - g * g) / (1 - g + 2 * g * x);
        cosTheta = (1 + g * g - k * k) / (2 * g);
    }
    const float sinTheta = sqrtf(fmaxf(0.0f, 1.0f - cosTheta * cosTheta));
    const float phi      = 2 * ry;
    const V3 wl          = mkv(sinTheta * fast_cospi(phi), sinTheta * fast_sinpi(phi), cosTheta);
    const float OdotI    = fminf(fmaxf(-wl.z, -1.0f), 1.0f);
    const float p1       = hg_phase(OdotI, g1);
    const float pdf      = (1 - blend) * p1 + blend * p1;   // LERP(blend, p1, p2) with g2 = g1
    TangentFrame f       = frame_from_normal(-wo);
    s.wi        = frame_get(f, wl.x, wl.y, wl.z);
    s.weight    = mkv(1.0f);
    s.pdf       = pdf;
    s.roughness = 1.0f;
    return s;
}
// MediumStack::integrate (shading.h:521-613) on the stack stored at `med`.  Returns true when
// the path scattered inside the medium: origin / direction / bsdf_pdf then describe the new ray.
OSLD bool medium_integrate(const float* med, const MediumHead& H, Ray& r, Sampler& sampler, float hit_t, V3& path_weight,
                           float& bsdf_pdf)
{
    if (H.depth <= 0)
        return false;
    // compute_current_params (shading.h:473-519)
    MediumData cur = medium_vacuum();
    int novl       = 0;
    float total_cdf = 0.0f;
    for (int i = 0; i < H.depth; i++) {
        const MediumData pi = medium_entry(med, mbyte(H.mediums, i));
        if (i == 0)
            cur.priority = pi.priority;
        if (pi.priority != cur.priority)
            continue;
        cur.sigma_t = cur.sigma_t + pi.sigma_t;
        cur.sigma_s = cur.sigma_s + pi.sigma_s;
        const float avg = (pi.sigma_s.x + pi.sigma_s.y + pi.sigma_s.z) / 3.0f;
        total_cdf       = (novl > 0 ? total_cdf : 0.0f) + avg;
        novl++;
    }
    const bool normalise = novl > 1 && !medium_is_vacuum(cur) && total_cdf > 0.0f;
    cur.sigma_s = mkv(fminf(cur.sigma_s.x, cur.sigma_t.x), fminf(cur.sigma_s.y, cur.sigma_t.y),
                      fminf(cur.sigma_s.z, cur.sigma_t.z));
    if (medium_is_vacuum(cur))
        return false;
    float cw0 = path_weight.x * cur.sigma_s.x / cur.sigma_t.x;
    float cw1 = path_weight.y * cur.sigma_s.y / cur.sigma_t.y;
    float cw2 = path_weight.z * cur.sigma_s.z / cur.sigma_t.z;
    const float total = cw0 + cw1 + cw2;
    if (total <= 0.0f) {
        path_weight = path_weight * medium_transmittance(cur.sigma_t, hit_t);
        return false;
    }
    const float inv_total = 1.0f / total;
    cw0 *= inv_total; cw1 *= inv_total; cw2 *= inv_total;
    const V3 rnd = sampler.get();
    int channel;
    if (rnd.y < cw0)
        channel = 0;
    else if (rnd.y < cw0 + cw1)
        channel = 1;
    else
        channel = 2;
    const float sigma_t_channel = vcomp(cur.sigma_t, channel);
    const float t_volume        = -libm_logf(1.0f - rnd.x) / sigma_t_channel;
    const bool scatter = t_volume < hit_t;
    const float t      = scatter ? t_volume : hit_t;
    const V3 tr        = medium_transmittance(cur.sigma_t, t);
    const V3 density   = scatter ? (cur.sigma_t * tr) : tr;
    const float pdf    = density.x * cw0 + density.y * cw1 + density.z * cw2;
    if (pdf <= 0.0f)
        return false;
    if (!scatter) {
        path_weight = path_weight * vdiv(tr, pdf);
        return false;
    }
    path_weight = path_weight * vdiv(tr * cur.sigma_s, pdf);
    r.origin    = ray_point(r, t_volume);
    // which of the overlapping media scatters: walk the cdf again
    int medium_index = 0, k = 0;
    float cum        = 0.0f;
    for (int i = 0; i < H.depth; i++) {
        const MediumData pi = medium_entry(med, mbyte(H.mediums, i));
        if (pi.priority != cur.priority)
            continue;
        medium_index    = i;   // ends on the last overlapping entry when no cdf entry stops the walk
        const float avg = (pi.sigma_s.x + pi.sigma_s.y + pi.sigma_s.z) / 3.0f;
        cum             = (k > 0 ? cum : 0.0f) + avg;
        if (novl > 1 && k < novl - 1 && rnd.z < (normalise ? cum / total_cdf : cum))
            break;
        k++;
    }
    const V3 rp      = sampler.get();
    const BSample ps = medium_sample_phase(medium_entry(med, mbyte(H.mediums, medium_index)), -r.direction, rp.x, rp.y);
    if (ps.pdf > 0.0f) {
        path_weight = path_weight * ps.weight;
        r.direction = ps.wi;
        bsdf_pdf    = ps.pdf;
        return true;
    }
    return false;
}
// process_medium_closure (shading.cpp:1283-1447): the medium a surface encloses, gathered before
// the BSDF pass.  (refraction_ior is write-only in the reference: not kept.)
OSLD void process_medium_closure(const ClosurePool& pool, int closure, MediumData& md, V3 wo, bool backfacing,
                                 float path_roughness, const float* luts)
{
    int ptr_stack[OSLD_CLOSURE_STACK];
    V3 weight_stack[OSLD_CLOSURE_STACK];
    int sp    = 0;
    V3 weight = mkv(1.0f);
    while (closure) {
        int id = pool.id(closure);
        if (id == CL_MUL) {
            weight  = weight * pool.weight(closure);
            closure = __float_as_int(pool.w[closure + 4]);
        } else if (id == CL_ADD) {
            weight_stack[sp] = weight;
            ptr_stack[sp++]  = __float_as_int(pool.w[closure + 2]);
            closure          = __float_as_int(pool.w[closure + 1]);
        } else {
            const V3 cw     = weight * pool.weight(closure);
            const PoolPtr q = pool.w + (closure + 4);
            closure         = 0;
            if (id == MX_ANISOTROPIC_VDF_ID) {
                // params: albedo, extinction, anisotropy
                md.sigma_t  = cw * mkv(q[3], q[4], q[5]);
                md.sigma_s  = mkv(q[0], q[1], q[2]) * md.sigma_t;
                md.g        = q[6];
                md.priority = 0;
                md.sigma_s  = mkv(fminf(md.sigma_s.x, md.sigma_t.x), fminf(md.sigma_s.y, md.sigma_t.y),
                                  fminf(md.sigma_s.z, md.sigma_t.z));
            } else if (id == MX_MEDIUM_VDF_ID) {
                // params: albedo, transmission_depth, transmission_color, anisotropy, ior, priority
                const V3 albedo = mkv(q[0], q[1], q[2]), tc = mkv(q[4], q[5], q[6]);
                if (albedo.x == 0 && albedo.y == 0 && albedo.z == 0 && tc.x == 0 && tc.y == 0 && tc.z == 0) {
                    md.sigma_t = md.sigma_s = mkv(0.0f);
                } else {
                    const float epsilon = 1e-10f;
                    const V3 st = mkv(-fast_log(fmaxf(tc.x, epsilon)), -fast_log(fmaxf(tc.y, epsilon)),
                                      -fast_log(fmaxf(tc.z, epsilon)));
                    md.sigma_t  = st * vdiv(cw, q[3]);
                    md.sigma_s  = albedo * md.sigma_t;
                    md.sigma_s  = mkv(fminf(md.sigma_s.x, md.sigma_t.x), fminf(md.sigma_s.y, md.sigma_t.y),
                                      fminf(md.sigma_s.z, md.sigma_t.z));
                }
                md.g        = q[7];
                md.priority = __float_as_int(q[9]);
            }
#ifdef OSLD_GLOSSY_LOBES
            else if (id == MX_LAYER_ID) {
                const int top = __float_as_int(q[0]), base = __float_as_int(q[1]);
                V3 op     = evaluate_layer_opacity(pool, top, wo, backfacing, path_roughness, luts);
                op        = mkv(fminf(fmaxf(op.x, 0.f), 1.f), fminf(fmaxf(op.y, 0.f), 1.f), fminf(fmaxf(op.z, 0.f), 1.f));
                V3 base_w = weight * (mkv(1.0f) - op);
                closure   = top;
                ptr_stack[sp]      = base;
                weight_stack[sp++] = weight * base_w;   // (sic)
            }
#endif
        }
        if (closure == 0 && sp > 0) {
            closure = ptr_stack[--sp];
            weight  = weight_stack[sp];
        }
    }
}
#endif  // OSLD_HAS_MEDIA

// ---- path life cycle ---------------------------------------------------------------------------
// Sample `sid` of the round = sample plane sid / npix of work-set pixel sid % npix
// (antialias_pixel, simpleraytracer.cpp:1197-1213): camera sample -> initial path state.
OSLD void path_start(const RenderLaunch& L, int slot, int sid)
{
    const RenderScene& S = L.S;
    const int sb = sid / L.npix, pix = sid - sb * L.npix;
    const int xy = __ldg(L.pixmap + pix);
    const int x = xy & 0xffff, y = (int)((unsigned)xy >> 16);
    const int si = L.s0 + sb;
    Sampler sampler;
    sampler.init(x, y, si);
    V3 j = S.no_jitter ? mkv(0.5f, 0.5f, 0.0f) : sampler.get();
    j.x *= 2;
    j.x = j.x < 1 ? sqrtf(j.x) - 1 : 1 - sqrtf(2 - j.x);
    j.y *= 2;
    j.y = j.y < 1 ? sqrtf(j.y) - 1 : 1 - sqrtf(2 - j.y);
    Ray r = camera_ray(S, (float)x + 0.5f + j.x, (float)y + 0.5f + j.y);
    float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
    rec[0] = mkf4(r.origin, r.radius);
    rec[1] = mkf4(r.direction, r.spread);
    rec[2] = make_float4(1.0f, 1.0f, 1.0f, OSLD_INF);
    rec[3] = make_float4(0.0f, 0.0f, 0.0f, r.roughness);
    rec[5] = mki4(r.raytype, -1, 0, (int)sampler.seed);
    rec[6] = mki4((int)sampler.index, sid, 0, 0);
#ifdef OSLD_HAS_MEDIA
    L.medium[(size_t)slot * OSLD_MEDIUM_WORDS] = 0.0f;   // depth 0, empty pool
#endif
}
// a finished sample: its radiance goes to the sample's own slot (rt_resolve folds in order)
OSLD void path_finish(const RenderLaunch& L, int slot)
{
    const float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
    const float4 q3 = rec[3];
    const int sid   = __float_as_int(rec[6].y);
    float* r        = L.result + 3 * (size_t)sid;
    r[0] = q3.x; r[1] = q3.y; r[2] = q3.z;
}
// path regeneration: a freed slot takes the next sample of the round, if any is left
OSLD bool path_regen(const RenderLaunch& L, int slot, bool want)
{
    unsigned m = __ballot_sync(__activemask(), want);
    if (!want)
        return false;
    int lane = threadIdx.x & 31, leader = __ffs(m) - 1, base = 0;
    if (lane == leader)
        base = atomicAdd(L.counters + C_NEXT, __popc(m));
    base    = __shfl_sync(m, base, leader);
    int sid = base + __popc(m & ((1u << lane) - 1u));
    if (sid >= L.total)
        return false;
    path_start(L, slot, sid);
    return true;
}

// one thread per path slot of the initial fill
extern "C" __global__ void __launch_bounds__(256) rt_generate(const __grid_constant__ RenderLaunch L)
{
    const int n0 = L.nslots < L.total ? L.nslots : L.total;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n0; slot += gridDim.x * blockDim.x) {
        path_start(L, slot, slot);
        L.queue_in[slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        L.counters[C_LIVE]   = n0;
        L.counters[C_NEXT]   = n0;
        L.counters[C_OUT]    = 0;
        L.counters[C_SHADOW] = L.counters[C_SHADOW_OUT] = L.counters[C_FETCH] = 0;
    }
}

// ---- rt_trace: every ray of the step -----------------------------------------------------------
// Work items [0, live) are closest-hit queries of the live paths, [live, live + shadow) are the
// slots with pending NEE shadow rays (background ray, then light ray: two any-hit walks).
// Persistent warps: a lane that finishes its item takes the next one from the global cursor
// while the other lanes keep walking (rays of very different length share a warp; without
// this the warp idles on its longest ray: ncu showed 7 of 32 lanes active in round 1).
#ifndef OSLD_TRACE_CHUNK
#define OSLD_TRACE_CHUNK 64   // iterations (node visits or triangle tests) of one trav_step_warp call
#endif
#ifndef OSLD_TRACE_REFILL
#define OSLD_TRACE_REFILL 16  // idle lanes that trigger a refill from the queues (sweep: profiles/render_tune_r02.txt)
#endif
#define OSLD_TRACE_BLOCK 128
extern "C" __global__ void __launch_bounds__(OSLD_TRACE_BLOCK) rt_trace(const __grid_constant__ RenderLaunch L)
{
    extern __shared__ unsigned smem_[];
    __shared__ int sh_hist[64];
    const RenderScene& S = L.S;
    unsigned* const stk  = smem_ + threadIdx.x;
    const int stride     = blockDim.x;
    const int nlive = L.counters[C_LIVE], n = nlive + L.counters[C_SHADOW];
    if (threadIdx.x < 64)
        sh_hist[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    bool have = false, exhausted = false;
    int q = 0, slot = 0, phase = 0, flags = 0, vis = 0;
    Trav T;
    for (;;) {
        unsigned hm = __ballot_sync(0xffffffffu, have);
        if (!exhausted && __popc(hm) <= 32 - OSLD_TRACE_REFILL) {
            const unsigned need = ~hm;
            const int leader = __ffs(need) - 1;
            int base = 0;
            if (lane == leader)
                base = atomicAdd(L.counters + C_FETCH, __popc(need));
            base      = __shfl_sync(0xffffffffu, base, leader);
            exhausted = base + __popc(need) >= n;
            if (!have) {
                q = base + __popc(need & ((1u << lane) - 1u));
                if (q < n) {
                    have = true;
                    if (q < nlive) {
                        slot = L.queue_in[q];
                        const float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
                        const float4 q0 = rec[0], q1 = rec[1], q5 = rec[5];
                        phase = 0;
                        trav_init(S, T, stk, stride, xyz(q0), xyz(q1), OSLD_INF, (unsigned)__float_as_int(q5.y),
                                  ~0u, 0);
                    } else {
                        slot = L.queue_sh[q - nlive];
                        const float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
                        const float4* sh  = L.shrec + (size_t)slot * OSLD_SHADOW_QUADS;
                        const float4 q0 = rec[0], s0 = sh[0];
                        const unsigned hid = (unsigned)__float_as_int(rec[5].y);
                        flags = __float_as_int(s0.w);
                        vis   = 0;
                        if (flags & SH_BG) {
                            phase = 1;
                            trav_init(S, T, stk, stride, xyz(q0), xyz(s0), OSLD_INF, hid, ~0u, 1);
                        } else {
                            const float4 s2 = sh[2];
                            phase = 2;
                            trav_init(S, T, stk, stride, xyz(q0), xyz(s2), s2.w, hid,
                                      (unsigned)__float_as_int(sh[3].w), 1);
                        }
                    }
                }
            }
            hm = __ballot_sync(0xffffffffu, have);
        }
        if (hm == 0u)
            break;
        if (trav_step_warp(S, T, have, OSLD_TRACE_CHUNK, exhausted ? 0 : 32 - OSLD_TRACE_REFILL)) {
            if (phase == 0) {
                float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
                rec[4]      = make_float4(T.hit.t, T.hit.u, T.hit.v, __int_as_float((int)T.hit.id));
                if (L.sort_keys) {
                    const int key = (T.hit.t == OSLD_INF) ? 0 : __ldg(L.shader_key + __ldg(S.shaderids + T.hit.id));
                    L.sort_keys[q] = key;
                    atomicAdd(&sh_hist[key], 1);
                }
                have = false;
            } else {
                float4* sh = L.shrec + (size_t)slot * OSLD_SHADOW_QUADS;
                if (phase == 1) {
                    vis |= (T.hit.t == OSLD_INF) ? 1 : 0;
                    if (flags & SH_LIGHT) {
                        // second shadow ray of the slot: towards the sampled light point
                        const float4 s2 = sh[2];
                        const V3 org    = T.org;
                        const unsigned hid = T.skip1;
                        phase = 2;
                        trav_init(S, T, stk, stride, org, xyz(s2), s2.w, hid, (unsigned)__float_as_int(sh[3].w), 1);
                    } else
                        have = false;
                } else {
                    vis |= (T.hit.t == T.tmax0) ? 2 : 0;
                    have = false;
                }
                if (!have)
                    reinterpret_cast<int*>(sh + 1)[3] = vis;
            }
        }
    }
    if (L.sort_keys) {
        __syncthreads();
        if (threadIdx.x < 64 && sh_hist[threadIdx.x])
            atomicAdd(L.counters + C_HIST + threadIdx.x, sh_hist[threadIdx.x]);
    }
}

// Counting sort of the live queue by bucket (0 = miss, else the material's closure-signature
// rank); the histogram comes from rt_trace.  Stable order is not required: paths are independent.
extern "C" __global__ void __launch_bounds__(256) rt_sort_scatter(const __grid_constant__ RenderLaunch L)
{
    __shared__ int sh_prefix[64], sh_cnt[64], sh_base[64];
    const int n = L.counters[C_LIVE];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 64) {
        // exclusive prefix of the 64 bucket counts (two warps, shuffle scan)
        int c = L.counters[C_HIST + tid], v = c;
        for (int d = 1; d < 32; d <<= 1) {
            int o = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d)
                v += o;
        }
        sh_cnt[tid] = v;  // inclusive, per warp
        __syncwarp();
        sh_prefix[tid] = v - c;
    }
    __syncthreads();
    if (tid >= 32 && tid < 64)
        sh_prefix[tid] += sh_cnt[31];
    __syncthreads();
    for (int tile = blockIdx.x; tile * 256 < n; tile += gridDim.x) {
        if (tid < 64)
            sh_cnt[tid] = 0;
        __syncthreads();
        const int q   = tile * 256 + tid;
        const int key = q < n ? L.sort_keys[q] : 63;
        const int slot = q < n ? L.queue_in[q] : 0;
        const unsigned m = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(m) - 1;
        int r = 0;
        if (lane == leader)
            r = atomicAdd(&sh_cnt[key], __popc(m));
        r = __shfl_sync(0xffffffffu, r, leader) + __popc(m & ((1u << lane) - 1u));
        __syncthreads();
        if (tid < 63 && sh_cnt[tid])
            sh_base[tid] = atomicAdd(L.counters + C_CURSOR + tid, sh_cnt[tid]);
        __syncthreads();
        if (q < n)
            L.queue_out[sh_prefix[key] + sh_base[key] + r] = slot;
        __syncthreads();
    }
}

// ---- rt_shade ------------------------------------------------------------------------------------
// One bounce of subpixel_radiance (simpleraytracer.cpp:975-1189) between two closest-hit
// queries, minus the two NEE shadow-ray walks: their rays are written to the slot's shadow
// record and traced with everything else by the next rt_trace.
// Returns bit 0: the path continues (its record holds the next ray); bit 1: shadow rays pending.
OSLD int shade_path(const RenderLaunch& L, int slot, ClosurePool& pool)
{
    const RenderScene& S = L.S;
    bool alive           = false;
    int shflags          = 0;
    float4* rec = L.rec + (size_t)slot * OSLD_PATH_QUADS;
    float4* sh  = L.shrec + (size_t)slot * OSLD_SHADOW_QUADS;
    const float4 q0 = rec[0], q1 = rec[1], q2 = rec[2], q3 = rec[3], q4 = rec[4], q5 = rec[5], q6 = rec[6];
    Ray r;
    r.origin    = xyz(q0);
    r.direction = xyz(q1);
    r.radius    = q0.w;
    r.spread    = q1.w;
    r.roughness = q3.w;
    r.raytype   = __float_as_int(q5.x);
    const float ht = q4.x, hu = q4.y, hv = q4.z;
    const int hid  = __float_as_int(q4.w);
    const int b    = __float_as_int(q5.z);
    V3 path_weight   = xyz(q2);
    V3 path_radiance = xyz(q3);
    float bsdf_pdf   = q2.w;
    float out_rough  = q3.w;
    u32 seed_now     = (u32)__float_as_int(q5.w);
    V3 bg_dir        = mkv(0.0f);
    do {
        if (ht == OSLD_INF) {
            // miss: background (simpleraytracer.cpp:980-996)
#ifdef OSLD_HAS_BACKGROUND
            if (S.background_shader >= 0) {
                if (b > 0 && S.bg_values) {
                    float bg_pdf = 0;
                    V3 bg        = bg_eval(S, r.direction, bg_pdf);
                    path_radiance = path_radiance + path_weight * bg * power_heuristic<WEIGHT_WEIGHT>(bsdf_pdf, bg_pdf);
                } else {
                    path_radiance = path_radiance
                                    + path_weight * eval_background(S, r.direction, mkv(0.0f), mkv(0.0f), b, pool);
                }
            }
#endif
            break;
        }
#ifdef OSLD_HAS_MEDIA
        // the medium the ray crossed on its way to the hit (simpleraytracer.cpp:999-1002)
        float* const med    = L.medium + (size_t)slot * OSLD_MEDIUM_WORDS;
        const MediumHead mh = medium_head(med);
        Sampler sampler;
        sampler.seed  = seed_now;
        sampler.index = (u32)__float_as_int(q6.x);
        if (medium_integrate(med, mh, r, sampler, ht, path_weight, bsdf_pdf)) {
            // scattered inside: a bounce without a surface; prev_id stays
            rec[0] = mkf4(r.origin, r.radius);
            rec[1] = mkf4(r.direction, r.spread);
            rec[2] = mkf4(path_weight, bsdf_pdf);
            rec[5] = mki4(r.raytype, __float_as_int(q5.y), b + 1, (int)sampler.seed);
            alive  = b + 1 <= S.max_bounces;
            break;
        }
        seed_now = sampler.seed;
#endif
        SG sg;
        globals_from_hit(S, sg, r, ht, hid, hu, hv);
        if (S.show_globals) {
            V3 v = sg.Ng;
            if (S.show_globals == 2) v = sg.N;
            if (S.show_globals == 3) v = vnormalized(sg.dPdu);
            if (S.show_globals == 4) v = vnormalized(sg.dPdv);
            if (S.show_globals == 5) v = mkv(sg.u, sg.v, 0.0f);
            V3 c = v;
            if (S.show_globals != 5)
                c = c * 0.5f + mkv(0.5f);
            path_radiance = path_radiance + path_weight * c;
            break;
        }
        const float radius = r.radius + r.spread * ht;
        const int shaderID = __ldg(S.shaderids + hid);
        if (shaderID < 0)
            break;
        pool.reset();
        sg.pool = &pool;
        sg.Ci   = 0;
        osl_execute_shader(shaderID, sg);
        V3 Le = mkv(0.0f);
        CompositeBSDF bsdf;
        bsdf.num               = 0;
        const bool last_bounce = b == S.max_bounces;
#ifdef OSLD_HAS_MEDIA
        MediumData md = medium_vacuum();
        if (!last_bounce)
            process_medium_closure(pool, sg.Ci, md, -sg.I, sg.backfacing != 0, r.roughness, S.bsdl_luts);
        const bool false_isect = medium_false_intersection(
            md.priority, mh.depth > 0, mh.depth > 0 ? medium_entry_priority(med, mbyte(mh.mediums, 0)) : 0);
        process_closure(pool, sg.Ci, Le, bsdf, last_bounce, -sg.I, sg.backfacing != 0, r.roughness, S.bsdl_luts,
                        false_isect);
#else
        process_closure(pool, sg.Ci, Le, bsdf, last_bounce, -sg.I, sg.backfacing != 0, r.roughness, S.bsdl_luts);
#endif
        const int nlights = S.nlightprims;
        float k           = 1;
        if (__ldg(S.shader_is_light + shaderID) && nlights > 0) {
            const float light_pick_pdf = 1.0f / nlights;
            float light_pdf            = light_pick_pdf * scene_shapepdf(S, hid, r.origin, sg.P);
            k                          = power_heuristic<WEIGHT_EVAL>(bsdf_pdf, light_pdf);
        }
        path_radiance = path_radiance + path_weight * k * Le;
        if (last_bounce)
            break;
        const V3 wo = -sg.I;
        bsdf_prepare(bsdf, wo, path_weight, b >= S.rr_depth);
#ifndef OSLD_HAS_MEDIA
        Sampler sampler;
        sampler.seed  = seed_now;
        sampler.index = (u32)__float_as_int(q6.x);
#endif
        V3 s          = sampler.get();
        seed_now      = sampler.seed;
        const float xi = s.x, yi = s.y, zi = s.z;
#ifdef OSLD_HAS_BACKGROUND
        if (S.bg_values) {
            // one shadow ray towards an importance-sampled background direction
            float bg_pdf = 0;
            V3 bg        = bg_sample(S, xi, yi, bg_dir, bg_pdf);
            BSample bs   = bsdf_eval(bsdf, wo, bg_dir);
            V3 contrib   = path_weight * bs.weight * bg * power_heuristic<WEIGHT_WEIGHT>(bg_pdf, bs.pdf);
            if ((contrib.x + contrib.y + contrib.z) > 0) {
                sh[1] = mkf4(contrib, 0.0f);
                shflags |= SH_BG;
            }
        }
#endif
        if (nlights > 0) {
            const float light_pick_pdf = 1.0f / nlights;
            float xl = xi * nlights;
            int ls   = (int)floorf(xl);
            xl -= ls;
            unsigned lid = __ldg(S.lightprims + ls);
            if (lid != (unsigned)hid) {
                LightSample sample = scene_sample(S, (int)lid, sg.P, xl, yi);
                BSample bs         = bsdf_eval(bsdf, wo, sample.dir);
                V3 contrib = path_weight * bs.weight * power_heuristic<EVAL_WEIGHT>(light_pick_pdf * sample.pdf, bs.pdf);
                if ((contrib.x + contrib.y + contrib.z) > 0) {
                    sh[2]  = mkf4(sample.dir, sample.dist);
                    sh[3]  = mkf4(contrib, __int_as_float((int)lid));
                    rec[7] = make_float4(sample.u, sample.v, 0.0f, 0.0f);
                    shflags |= SH_LIGHT;
                }
            }
        }
        BSample p   = bsdf_sample(bsdf, wo, xi, yi, zi);
        path_weight = path_weight * p.weight;
        bsdf_pdf    = p.pdf;
        // the shadow rays start at P and skip this primitive, whether or not the path goes on
        rec[0] = mkf4(sg.P, radius);
        rec[5] = mki4(RAY_DIFFUSE, hid, b + 1, (int)seed_now);
#ifdef OSLD_HAS_MEDIA
        if (dot3(sg.Ng, p.wi) < 0) {   // the sampled direction crosses the surface (simpleraytracer.cpp:1174-1182)
            if (!sg.backfacing)
                medium_add(med, md);
            else
                medium_pop(med);
        }
#endif
        if (!(path_weight.x > 0) && !(path_weight.y > 0) && !(path_weight.z > 0))
            break;
        // continue the path
        rec[1]    = mkf4(p.wi, fmaxf(r.spread, p.roughness));
        rec[2]    = mkf4(path_weight, bsdf_pdf);
        out_rough = p.roughness;
        alive     = true;
    } while (false);
    rec[3] = mkf4(path_radiance, out_rough);
    if (shflags)
        sh[0] = mkf4(bg_dir, __int_as_float(shflags | (alive ? 0 : SH_DEAD)));
    return (alive ? 1 : 0) | (shflags ? 2 : 0);
}

// NEE of the previous bounce, after rt_trace has answered the visibility of its shadow rays
// (simpleraytracer.cpp:1083-1161): unoccluded background sample adds its contribution, an
// unoccluded light sample runs the light's shader for its emission.  Returns whether the
// path ended at that bounce.
OSLD bool light_path(const RenderLaunch& L, int slot, ClosurePool& pool)
{
    const RenderScene& S = L.S;
    float4* rec      = L.rec + (size_t)slot * OSLD_PATH_QUADS;
    const float4* sh = L.shrec + (size_t)slot * OSLD_SHADOW_QUADS;
    const float4 s0 = sh[0], s1 = sh[1], q3 = rec[3];
    const int flags = __float_as_int(s0.w), vis = __float_as_int(s1.w);
    V3 rad          = xyz(q3);
    if ((flags & SH_BG) && (vis & 1))
        rad = rad + xyz(s1);
    if ((flags & SH_LIGHT) && (vis & 2)) {
        const float4 s2 = sh[2], s3 = sh[3], q0 = rec[0], q7 = rec[7];
        const int lid = __float_as_int(s3.w);
        Ray shadow_ray;
        shadow_ray.origin    = xyz(q0);
        shadow_ray.direction = xyz(s2);
        shadow_ray.radius    = q0.w;
        shadow_ray.spread = shadow_ray.roughness = 0.0f;
        shadow_ray.raytype = RAY_SHADOW;
        SG lsg;
        globals_from_hit(S, lsg, shadow_ray, s2.w, lid, q7.x, q7.y);
        pool.reset();
        lsg.pool = &pool;
        lsg.Ci   = 0;
        osl_execute_shader(__ldg(S.shaderids + lid), lsg);
        V3 lLe = mkv(0.0f);
        CompositeBSDF dummy;
        dummy.num = 0;
        process_closure(pool, lsg.Ci, lLe, dummy, true);
        rad = rad + xyz(s3) * lLe;
    }
    rec[3] = mkf4(rad, q3.w);
    return (flags & SH_DEAD) != 0;
}

#define OSLD_SHADE_BLOCK 128
#ifndef OSLD_SHADE_MINBLOCKS
#define OSLD_SHADE_MINBLOCKS 8   // 64 registers: 32 resident warps per SM (4, 6 and 8 time the same)
#endif
extern "C" __global__ void __launch_bounds__(OSLD_SHADE_BLOCK, OSLD_SHADE_MINBLOCKS) rt_shade(const __grid_constant__ RenderLaunch L)
{
    extern __shared__ unsigned smem_[];
    OSLD_POOL_DECL(pool, smem_);
    const int n = L.counters[C_LIVE];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ((n + 31) & ~31); q += gridDim.x * blockDim.x) {
        int fl = 0, slot = 0;
        if (q < n) {
            slot = L.queue_in[q];
            fl   = shade_path(L, slot, pool);
        }
        bool alive = (fl & 1) != 0;
        const bool pending = (fl & 2) != 0, ended = q < n && fl == 0;
        if (ended)
            path_finish(L, slot);
        if (path_regen(L, slot, ended))
            alive = true;
        queue_push(L.queue_out, L.counters + C_OUT, slot, alive);
        queue_push(L.queue_sh, L.counters + C_SHADOW_OUT, slot, pending);
    }
}

extern "C" __global__ void __launch_bounds__(OSLD_SHADE_BLOCK) rt_light(const __grid_constant__ RenderLaunch L)
{
    extern __shared__ unsigned smem_[];
    OSLD_POOL_DECL(pool, smem_);
    const int n = L.counters[C_SHADOW];
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < ((n + 31) & ~31); q += gridDim.x * blockDim.x) {
        bool ended = false;
        int slot   = 0;
        if (q < n) {
            slot  = L.queue_sh[q];
            ended = light_path(L, slot, pool);
            if (ended)
                path_finish(L, slot);
        }
        const bool alive = path_regen(L, slot, ended);
        queue_push(L.queue_out, L.counters + C_OUT, slot, alive);
    }
}

// end of a step: the out queue becomes the in queue (the host rotates the pointers), the
// shadow entries shade appended become the next step's, and the host's copy of the loop
// state is refreshed (mapped pinned memory: the host polls it instead of synchronising)
extern "C" __global__ void __launch_bounds__(128) rt_swap(const __grid_constant__ RenderLaunch L)
{
    if (blockIdx.x != 0)
        return;
    L.counters[C_HIST + threadIdx.x] = 0;  // histogram + cursors
    if (threadIdx.x == 0) {
        const int live = L.counters[C_OUT], shadow = L.counters[C_SHADOW_OUT];
        L.counters[C_LIVE]       = live;
        L.counters[C_OUT]        = 0;
        L.counters[C_SHADOW]     = shadow;
        L.counters[C_SHADOW_OUT] = 0;
        L.counters[C_FETCH]      = 0;
        const int it             = L.counters[C_ITER] + 1;
        L.counters[C_ITER]       = it;
        if (L.host_state) {
            L.host_state[1] = live;
            L.host_state[2] = shadow;
            L.host_state[3] = L.counters[C_NEXT];
            __threadfence_system();
            L.host_state[0] = it;
        }
    }
}

// The end of a work set: once no sample is left to start and only a few paths are alive (glass
// interiors keep a handful going for 10^4..10^5 bounces) a step costs five launches for almost
// no work.  Each remaining path is then run to its end by one thread, one path per warp (paths
// of very different lengths sharing a warp would serialise each other's bounces).  Same
// per-path arithmetic as the staged kernels, hence the same pixels.  Expects the closest hit of
// every queued path in its record and no pending shadow entries (the host runs rt_trace +
// rt_light first).
extern "C" __global__ void __launch_bounds__(32) rt_tail(const __grid_constant__ RenderLaunch L)
{
    extern __shared__ unsigned smem_[];
    if (threadIdx.x != 0)
        return;
    const RenderScene& S = L.S;
    unsigned* const stk  = smem_;
    OSLD_POOL_DECL(pool, smem_ + OSLD_STK_WORDS * OSLD_BVH_STACK * 32);   // OSLD_POOL_STORE words per thread
    const int n = L.counters[C_LIVE];
    for (int q = blockIdx.x; q < n; q += gridDim.x) {
        const int slot = L.queue_in[q];
        float4* rec    = L.rec + (size_t)slot * OSLD_PATH_QUADS;
        float4* sh     = L.shrec + (size_t)slot * OSLD_SHADOW_QUADS;
        for (;;) {
            const int fl = shade_path(L, slot, pool);
            if (fl & 2) {
                const float4 q0 = rec[0], s0 = sh[0];
                const unsigned hid = (unsigned)__float_as_int(rec[5].y);
                const int flags    = __float_as_int(s0.w);
                int vis            = 0;
                if (flags & SH_BG)
                    vis |= scene_intersect(S, stk, 32, xyz(q0), xyz(s0), OSLD_INF, hid, ~0u, 1).t == OSLD_INF ? 1 : 0;
                if (flags & SH_LIGHT) {
                    const float4 s2 = sh[2];
                    vis |= scene_intersect(S, stk, 32, xyz(q0), xyz(s2), s2.w, hid, (unsigned)__float_as_int(sh[3].w), 1).t
                                   == s2.w ? 2 : 0;
                }
                reinterpret_cast<int*>(sh + 1)[3] = vis;
                light_path(L, slot, pool);
            }
            if (!(fl & 1)) {
                path_finish(L, slot);
                break;
            }
            const float4 q0 = rec[0], q1 = rec[1], q5 = rec[5];
            const Hit h = scene_intersect(S, stk, 32, xyz(q0), xyz(q1), OSLD_INF, (unsigned)__float_as_int(q5.y), ~0u);
            rec[4]      = make_float4(h.t, h.u, h.v, __int_as_float((int)h.id));
        }
    }
}

// fold the round's samples into the running image in the reference's order:
// result = lerp(result, r, 1/(si+1))  for si = s0 .. s0+nsamples-1  (simpleraytracer.cpp:1211-1213)
extern "C" __global__ void __launch_bounds__(256) rt_resolve(const __grid_constant__ RenderLaunch L)
{
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < L.npix; pix += gridDim.x * blockDim.x) {
        V3 result = L.s0 == 0 ? mkv(0.0f) : mkv(L.accum[3 * pix], L.accum[3 * pix + 1], L.accum[3 * pix + 2]);
        for (int sb = 0; sb < L.nsamples; ++sb) {
            const float* rp = L.result + 3 * ((size_t)sb * L.npix + pix);
            V3 r     = mkv(rp[0], rp[1], rp[2]);
            float t  = 1.0f / (float)(L.s0 + sb + 1);
            result   = result * (1.0f - t) + r * t;
        }
        L.accum[3 * pix]     = result.x;
        L.accum[3 * pix + 1] = result.y;
        L.accum[3 * pix + 2] = result.z;
    }
}
