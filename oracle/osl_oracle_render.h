// osl_oracle_render.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Scalar restatement of the reference path tracer `testrender`:
//   Ray / Camera / Scene queries      src/testrender/raytracer.h:40-344
//   BVH traversal + triangle test     src/testrender/bvh.cpp:221-356
//   Sampler (Owen-scrambled Sobol)    src/testrender/sampling.h:212-295
//   TangentFrame, Sampling, MIS       src/testrender/sampling.h:17-207
//   fresnel_*                         src/testrender/optics.h:13-111
//   Diffuse / Reflection / Refraction / Transparent lobes
//                                     src/testrender/shading.cpp:301-322, 1081-1150
//   CompositeBSDF                     src/testrender/shading.h:319-437
//   process_closure                   src/testrender/shading.cpp:1448-1706
//   globals_from_hit, subpixel_radiance, antialias_pixel
//                                     src/testrender/simpleraytracer.cpp:889-1216
// One pixel at a time, recursive-loop integrator, exactly the reference's
// control flow.  Scene arrays (tessellated meshes, BVH, light list) are
// prepared by the shared host harness and passed in.  Not restated yet:
// participating media (MediumStack), background importance sampling,
// displacement — none of which the cornell / bunny configs use.
#pragma once
#include "osl_oracle_closure.h"
#include "osl_oracle_runtime.h"

namespace oslo {

// ---- Imath-style helpers ------------------------------------------------------
inline float length2(const V3& v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline float vlength(const V3& v) { return imath_length(v); }
inline V3 normalized(V3 v)
{
    float l = imath_length(v);
    if (l != 0.0f) {
        v.x /= l;
        v.y /= l;
        v.z /= l;
    }
    return v;
}

struct RenderScene {
    int nverts, ntris, nnodes, nlightprims, nshaders, nmeshes;
    const float* verts;         // 3 per vertex
    const float* normals;       // 3 per normal
    const float* uvs;           // 2 per uv
    const int* triangles;       // 3 per triangle
    const int* n_triangles;     // 3 per triangle, -1 = none
    const int* uv_triangles;    // 3 per triangle, -1 = none
    const int* shaderids;       // per triangle
    const int* meshids;         // per triangle
    const float* mesh_surfacearea;
    const float* bvh_nodes;     // 8 words per node: bounds[6], child, nprims
    const unsigned* bvh_indices;
    const unsigned* lightprims;
    const int* shader_is_light;
    float eye[3], dir[3], up[3], fov;   // as given in the scene file
    float cx[3], cy[3], invw, invh;     // derived by camera_finalize()
    int xres, yres;
    int aa, max_bounces, rr_depth, no_jitter, show_globals;
    int background_shader, background_resolution;
};

inline V3 vert(const RenderScene& S, int i) { return V3(S.verts[3 * i], S.verts[3 * i + 1], S.verts[3 * i + 2]); }
inline V3 nrm(const RenderScene& S, int i) { return V3(S.normals[3 * i], S.normals[3 * i + 1], S.normals[3 * i + 2]); }

// ---- Ray / Camera -------------------------------------------------------------
enum RayType { RAY_CAMERA = 1, RAY_SHADOW = 2, RAY_REFLECTION = 4, RAY_REFRACTION = 8, RAY_DIFFUSE = 16 };

inline void ortho(const V3& n, V3& x, V3& y)
{
    x = normalized(std::fabs(n.x) > .01f ? V3(n.z, 0, -n.x) : V3(0, -n.z, n.y));
    y = cross(n, x);
}
struct Ray {
    V3 origin, direction;
    float radius, spread, roughness;
    int raytype;
    V3 point(float t) const { return origin + direction * t; }
    Dv dual_direction() const
    {
        Dv v;
        v.val = direction;
        ortho(direction, v.dx, v.dy);
        v.dx = v.dx * spread;
        v.dy = v.dy * spread;
        return v;
    }
    Dv point_dual(float t) const
    {
        const float r = radius + spread * t;
        Dv p;
        p.val = point(t);
        ortho(direction, p.dx, p.dy);
        p.dx = p.dx * r;
        p.dy = p.dy * r;
        return p;
    }
};
// Camera::lookat + resolution + finalize (raytracer.h:95-126)
inline void camera_finalize(RenderScene& S)
{
    V3 dir = normalized(V3(S.dir[0], S.dir[1], S.dir[2]));
    V3 up(S.up[0], S.up[1], S.up[2]);
    S.invw   = 1.0f / S.xres;
    S.invh   = 1.0f / S.yres;
    float k  = fast_tan(S.fov * float(M_PI / 360));
    V3 right = normalized(cross(dir, up));
    V3 cx    = right * (S.xres * k / S.yres);
    V3 cy    = normalized(cross(cx, dir)) * k;
    for (int i = 0; i < 3; ++i) {
        S.dir[i] = dir[i];
        S.cx[i]  = cx[i];
        S.cy[i]  = cy[i];
    }
}
inline Ray camera_ray(const RenderScene& S, float x, float y)
{
    V3 cx(S.cx[0], S.cx[1], S.cx[2]), cy(S.cy[0], S.cy[1], S.cy[2]), dir(S.dir[0], S.dir[1], S.dir[2]);
    const V3 v        = normalized(cx * (x * S.invw - 0.5f) + cy * (0.5f - y * S.invh) + dir);
    const float cos_a = dot(dir, v);
    const float spread = std::sqrt(S.invw * S.invh * vlength(cx) * vlength(cy) * cos_a) * cos_a;
    Ray r;
    r.origin    = V3(S.eye[0], S.eye[1], S.eye[2]);
    r.direction = v;
    r.radius    = 0;
    r.spread    = spread;
    r.roughness = 0.0f;
    r.raytype   = RAY_CAMERA;
    return r;
}

// ---- Sampler --------------------------------------------------------------------
struct Sampler {
    uint32_t seed, index;
    Sampler(int px, int py, int si) : seed(((px & 2047) << 22) | ((py & 2047) << 11)), index(reversebits(si)) {}
    static uint32_t hash(uint32_t s)
    {
        s ^= s >> 16; s *= 0x21f0aaadu; s ^= s >> 15; s *= 0xd35a2d97u; s ^= s >> 15;
        return s;
    }
    static uint32_t reversebits(uint32_t x)
    {
        x = (x << 16) | (x >> 16);
        x = ((x & 0x00ff00ff) << 8) | ((x & 0xff00ff00) >> 8);
        x = ((x & 0x0f0f0f0f) << 4) | ((x & 0xf0f0f0f0) >> 4);
        x = ((x & 0x33333333) << 2) | ((x & 0xcccccccc) >> 2);
        x = ((x & 0x55555555) << 1) | ((x & 0xaaaaaaaa) >> 1);
        return x;
    }
    static uint32_t owen_scramble(uint32_t p, uint32_t s)
    {
        p ^= p * 0x3d20adea; p += s; p *= (s >> 16) | 1; p ^= p * 0x05526c56; p ^= p * 0x53a22864;
        return reversebits(p);
    }
    V3 get()
    {
        static const uint32_t zmatrix[24]
            = { 0x000001u, 0x000003u, 0x000006u, 0x000009u, 0x000017u, 0x00003au, 0x000071u, 0x0000a3u,
                0x000116u, 0x000339u, 0x000677u, 0x0009aau, 0x001601u, 0x003903u, 0x007706u, 0x00aa09u,
                0x010117u, 0x03033au, 0x060671u, 0x0909a3u, 0x171616u, 0x3a3939u, 0x717777u, 0xa3aaaau };
        seed += 4;
        uint32_t si = owen_scramble(index, hash(seed - 4)) & 0xFFFFFF;
        uint32_t rx = si, ry = 0, rz = 0, ymatrix = 1;
        for (int c = 0; c < 24; c++) {
            uint32_t bit = (si >> c) & 1;
            ry ^= bit * ymatrix;
            rz ^= bit * zmatrix[c];
            ymatrix ^= ymatrix << 1;
        }
        return V3((owen_scramble(rx, hash(seed - 3)) >> 8) * 5.96046448e-8f,
                  (owen_scramble(ry, hash(seed - 2)) >> 8) * 5.96046448e-8f,
                  (owen_scramble(rz, hash(seed - 1)) >> 8) * 5.96046448e-8f);
    }
};

// ---- sampling helpers -------------------------------------------------------------
struct TangentFrame {
    V3 u, v, w;
    static TangentFrame from_normal(const V3& n)
    {
        const float sign = std::copysign(1.0f, n.z);
        const float a    = -1 / (sign + n.z);
        const float b    = n.x * n.y * a;
        TangentFrame f;
        f.u = V3(1 + sign * n.x * n.x * a, sign * b, -sign * n.x);
        f.v = V3(b, sign + n.y * n.y * a, -n.y);
        f.w = n;
        return f;
    }
    static TangentFrame from_normal_and_tangent(const V3& n, const V3& t)
    {
        V3 x        = t - n * dot(n, t);
        float xlen2 = dot(x, x);
        if (xlen2 > 0) {
            x = x * (1.0f / std::sqrt(xlen2));
            TangentFrame f;
            f.u = x;
            f.v = cross(n, x);
            f.w = n;
            return f;
        }
        return from_normal(n);
    }
    V3 get(float x, float y, float z) const { return x * u + y * v + z * w; }
    float getx(const V3& a) const { return dot(a, u); }
    float gety(const V3& a) const { return dot(a, v); }
    float getz(const V3& a) const { return dot(a, w); }
    V3 tolocal(const V3& a) const { return V3(dot(a, u), dot(a, v), dot(a, w)); }
    V3 toworld(const V3& a) const { return get(a.x, a.y, a.z); }
};
inline void to_unit_disk(float& x, float& y)
{
    const float PI_OVER_4 = float(M_PI_4), PI_OVER_2 = float(M_PI_2);
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
inline void sample_cosine_hemisphere(const V3& N, float rndx, float rndy, V3& out, float& pdf)
{
    to_unit_disk(rndx, rndy);
    float cos_theta = std::sqrt(std::max(1 - rndx * rndx - rndy * rndy, 0.0f));
    out             = TangentFrame::from_normal(N).get(rndx, rndy, cos_theta);
    pdf             = cos_theta * float(M_1_PI);
}
enum MISMode { WEIGHT_WEIGHT, WEIGHT_EVAL, EVAL_WEIGHT };
template<MISMode mode> inline float power_heuristic(float sampled_pdf, float other_pdf)
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
    const float MAX = std::numeric_limits<float>::max();
    switch (mode) {
    case WEIGHT_WEIGHT: return std::min(other_pdf, MAX) * mis;
    case WEIGHT_EVAL: return mis;
    case EVAL_WEIGHT: return mis * ((other_pdf > sampled_pdf) ? std::min(1 / r, MAX) : r);
    }
    return 0;
}
inline void update_eval(V3* w, float* pdf, V3 ow, float opdf, float b)
{
    if (b > std::numeric_limits<float>::min()) {
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

// ---- fresnel (optics.h) -----------------------------------------------------------
inline float fresnel_dielectric(float cosi, float eta)
{
    if (eta == 0)
        return 1;
    if (cosi < 0.0f)
        eta = 1.0f / eta;
    float c = std::fabs(cosi);
    float g = eta * eta - 1 + c * c;
    if (g > 0) {
        g       = std::sqrt(g);
        float A = (g - c) / (g + c);
        float B = (c * (g + c) - 1) / (c * (g - c) + 1);
        return 0.5f * A * A * (1 + B * B);
    }
    return 1.0f;
}
inline float fresnel_refraction(const V3& I, const V3& N, float eta, V3& T)
{
    float cosi = -dot(I, N);
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
        float dnp = std::sqrt(arg);
        float nK  = (neta * cosi) - dnp;
        T         = I * neta + Nn * nK;
        return 1 - fresnel_dielectric(cosi, eta);
    }
    T = V3(0.0f);
    return 0;
}

// ---- BSDF lobes -----------------------------------------------------------------
struct BSample {
    V3 wi, weight;
    float pdf = 0, roughness = 0;
    BSample() : wi(0.0f), weight(0.0f) {}
    BSample(V3 wi, V3 w, float pdf, float r) : wi(wi), weight(w), pdf(pdf), roughness(r) {}
};
enum LobeType { LOBE_DIFFUSE, LOBE_TRANSLUCENT, LOBE_REFLECTION, LOBE_REFRACTION, LOBE_TRANSPARENT,
                LOBE_PHONG, LOBE_WARD, LOBE_MICROFACET,
                LOBE_BSDL_OREN_NAYAR /* libbsdl mtx::OrenNayarDiffuseLobe through BSDL_WRAP */,
                LOBE_BSDL_BURLEY /* mtx::BurleyDiffuseLobe */,
                LOBE_BSDL_SHEEN /* mtx::SheenLobe, Conty-Kulla mode */,
                LOBE_MX_SPEC /* mtx::ConductorLobe / DielectricLobe / SchlickLobe */,
                LOBE_MX_TRANSLUCENT /* mtx::TranslucentLobe */,
                LOBE_SPI_THINLAYER /* spi::ThinLayerLobe (SpiThinLayer, shading.cpp:119-152) */ };
}  // namespace oslo
#include "osl_oracle_mxlobes.h"
namespace oslo {
struct Lobe;
// Phong / Ward / Microfacet live in osl_oracle_lobes.h
V3 ext_albedo(const Lobe& l, const V3& wo);
BSample ext_eval(const Lobe& l, const V3& wo, const V3& wi);
BSample ext_sample(const Lobe& l, const V3& wo, float rx, float ry, float rz);
struct Lobe {
    int type;
    V3 N;
    float eta;
    // Phong: exponent ; Ward: T, ax, ay ; Microfacet: U(=T), xalpha(=ax), yalpha(=ay), eta, refract, ggx
    V3 T;
    float ax = 0, ay = 0, exponent = 0;
    int refract = 0, ggx = 0;
    // libbsdl diffuse lobes: albedo, roughness (in ax), energy compensation flag
    V3 albedo = V3(1.0f);
    int energy_compensation = 0;
    // sheen: ax = sheen alpha, ay = regularized roughness, emiss = layering transmittance
    float emiss = 1.0f;
    bool backfacing = false;
    TangentFrame tf;
    MxSpec mx;   // conductor / dielectric / generalized schlick state (LOBE_MX_SPEC)
    ThinSpec thin;   // LOBE_SPI_THINLAYER
    V3 get_albedo(const V3& wo) const
    {
        switch (type) {
        case LOBE_REFLECTION: {
            float cosNO = dot(N, wo);
            if (cosNO > 0)
                return V3(fresnel_dielectric(cosNO, eta));
            return V3(1.0f);
        }
        case LOBE_REFRACTION: return V3(1 - fresnel_dielectric(dot(N, wo), eta));
        case LOBE_PHONG:
        case LOBE_WARD:
        case LOBE_MICROFACET: return ext_albedo(*this, wo);
        case LOBE_BSDL_OREN_NAYAR:
        case LOBE_BSDL_BURLEY: return albedo;  // BSDL_WRAP::get_albedo = albedo_impl().toRGB(0)
        case LOBE_BSDL_SHEEN: return albedo * (1 - emiss);
        case LOBE_MX_SPEC: return mx_albedo(mx);
        case LOBE_MX_TRANSLUCENT: return albedo;
        default: return V3(1.0f);
        }
    }
    BSample eval(const V3& wo, const V3& wi) const
    {
        if (type >= LOBE_PHONG)
            return ext_eval(*this, wo, wi);
        if (type == LOBE_DIFFUSE || type == LOBE_TRANSLUCENT) {
            const float pdf = std::max(dot(N, wi), 0.0f) * float(M_1_PI);
            return BSample(wi, V3(1.0f), pdf, 1.0f);
        }
        return BSample();
    }
    BSample sample(const V3& wo, float rx, float ry, float rz) const
    {
        switch (type) {
        case LOBE_DIFFUSE:
        case LOBE_TRANSLUCENT: {
            V3 out;
            float pdf;
            sample_cosine_hemisphere(N, rx, ry, out, pdf);
            return BSample(out, V3(1.0f), pdf, 1.0f);
        }
        case LOBE_REFLECTION: {
            float cosNO = dot(N, wo);
            if (cosNO > 0) {
                V3 wi = (2 * cosNO) * N - wo;
                return BSample(wi, V3(fresnel_dielectric(cosNO, eta)), std::numeric_limits<float>::infinity(), 0);
            }
            return BSample();
        }
        case LOBE_REFRACTION: {
            V3 wi;
            float Ft = fresnel_refraction(-wo, N, eta, wi);
            return BSample(wi, V3(Ft), std::numeric_limits<float>::infinity(), 0);
        }
        case LOBE_TRANSPARENT: return BSample(-wo, V3(1.0f), std::numeric_limits<float>::infinity(), 0);
        default: return ext_sample(*this, wo, rx, ry, rz);
        }
        return BSample();
    }
};
}  // namespace oslo
#include "osl_oracle_lobes.h"
namespace oslo {

struct CompositeBSDF {
    enum { MaxEntries = 8 };
    V3 weights[MaxEntries];
    float pdfs[MaxEntries];
    Lobe lobes[MaxEntries];
    int num = 0;
    bool add(const V3& w, const Lobe& l)
    {
        if (num >= MaxEntries)
            return false;
        weights[num] = w;
        lobes[num]   = l;
        ++num;
        return true;
    }
    void prepare(const V3& wo, const V3& path_weight, bool absorb)
    {
        float total = 0;
        for (int i = 0; i < num; i++) {
            pdfs[i] = dot(weights[i], path_weight * lobes[i].get_albedo(wo))
                      / (path_weight.x + path_weight.y + path_weight.z);
            total += pdfs[i];
        }
        if ((!absorb && total > 0) || total > 1)
            for (int i = 0; i < num; i++)
                pdfs[i] /= total;
    }
    BSample eval(const V3& wo, const V3& wi) const
    {
        BSample s;
        for (int i = 0; i < num; i++) {
            BSample b = lobes[i].eval(wo, wi);
            b.weight  = b.weight * weights[i];
            update_eval(&s.weight, &s.pdf, b.weight, b.pdf, pdfs[i]);
            s.roughness += b.roughness * pdfs[i];
        }
        return s;
    }
    BSample sample(const V3& wo, float rx, float ry, float rz) const
    {
        float accum = 0;
        for (int i = 0; i < num; i++) {
            if (rx < (pdfs[i] + accum)) {
                rx        = (rx - accum) / pdfs[i];
                rx        = std::min(rx, 0.99999994f);
                BSample s = lobes[i].sample(wo, rx, ry, rz);
                s.weight  = s.weight * (weights[i] * (1 / pdfs[i]));
                s.pdf *= pdfs[i];
                if (s.pdf == 0.0f)
                    return BSample();
                for (int j = 0; j < num; j++) {
                    if (i != j) {
                        BSample b = lobes[j].eval(wo, s.wi);
                        b.weight  = b.weight * weights[j];
                        update_eval(&s.weight, &s.pdf, b.weight, b.pdf, pdfs[j]);
                    }
                }
                return s;
            }
            accum += pdfs[i];
        }
        return BSample();
    }
};
// ---- participating media (shading.h:438-725, shading.cpp:1152-1196) ---------------------------
// OIIO::fast_sinpi / fast_cospi (fmath.h; not in the reference tree): parabola-based sine of pi*x
inline float fast_sinpi(float x)
{
    const float z = x - ((x + 25165824.0f) - 25165824.0f);   // strip the integral part: [-1, 1]
    const float y = z - z * std::fabs(z);
    const float Q = 3.10396624f;
    const float P = 3.584135056f;
    return y * (Q + P * std::fabs(y));
}
inline float fast_cospi(float x) { return fast_sinpi(x + 0.5f); }

struct MediumParams {
    V3 sigma_t = V3(0.0f);   // extinction coefficient; vacuum by default
    V3 sigma_s = V3(0.0f);   // scattering
    float medium_g       = 0.0f;
    float refraction_ior = 1.0f;
    int priority         = 0;
    bool is_vaccum() const { return sigma_t.x <= 0.0f && sigma_t.y <= 0.0f && sigma_t.z <= 0.0f; }
    bool is_special_priority() const { return priority == 0; }
    // HenyeyGreenstein = bsdl::spi::VolumeLobe{g, g, blend 0} in the frame Frame(-wo)
    // (shading.cpp:1152-1196, BSDL/SPI/bsdf_volume_impl.h)
    static float phase_func(float costheta, float g)
    {
        if (g == 0)
            return 0.25f * (1 / float(M_PI));
        const float num = 0.25f * (1 / float(M_PI)) * (1 - g * g);
        const float den = 1 + g * g + 2.0f * g * costheta;
        return num / std::sqrt(den * den * den);
    }
    BSample sample_phase_func(const V3& wo, float rx, float ry, float /*rz*/) const
    {
        if (is_vaccum())
            return BSample(V3(1.0f), V3(1.0f), 0.0f, 0.0f);
        const float g1 = mx_clamp(medium_g, -0.99f, 0.99f), g2 = g1, blend = 0.0f;
        // sample_phase
        float g, x;
        if (rx < blend) {
            g = g2;
            x = rx / blend;
        } else {
            g = g1;
            x = (rx - blend) / (1 - blend);
        }
        float cosTheta;
        if (std::fabs(g) < 1e-3f)
            cosTheta = 1 - 2 * x;
        else {
            float k  = (1 - g * g) / (1 - g + 2 * g * x);
            cosTheta = (1 + g * g - k * k) / (2 * g);
        }
        float sinTheta = std::sqrt(std::max(0.0f, 1.0f - cosTheta * cosTheta));
        float phi      = 2 * ry;
        V3 wi_l(sinTheta * fast_cospi(phi), sinTheta * fast_sinpi(phi), cosTheta);
        // eval_impl: pdf = lerp(blend, phase(g1), phase(g2)) of cos = clamp(-wi.z)
        float OdotI = mx_clamp(-wi_l.z, -1.0f, 1.0f);
        float pdf   = mx_lerp(blend, phase_func(OdotI, g1), phase_func(OdotI, g2));
        TangentFrame f = TangentFrame::from_normal(-wo);   // bsdl::Frame(Z) = the same orthonormal basis
        return BSample(f.get(wi_l.x, wi_l.y, wi_l.z), V3(1.0f), pdf, 1.0f);
    }
};

struct MediumStack {
    enum { MaxEntries = 8 };
    MediumParams pool[MaxEntries];
    int mediums[MaxEntries];       // pool indices, sorted by descending priority
    int entry_order[MaxEntries];   // pool indices in LIFO order
    float cdf[MaxEntries];
    int overlapping_medium_indices[MaxEntries];
    MediumParams current_params;
    int depth = 0, pool_size = 0, num_overlapping = 0;

    const MediumParams* get_current_params() const { return depth > 0 ? &pool[mediums[0]] : nullptr; }
    void compute_current_params()
    {
        MediumParams np;
        num_overlapping = 0;
        for (int i = 0; i < depth; i++) {
            const MediumParams& pi = pool[mediums[i]];
            if (i == 0)
                np.priority = pi.priority;
            if (pi.priority != np.priority)
                continue;
            overlapping_medium_indices[num_overlapping] = i;
            np.sigma_t = np.sigma_t + pi.sigma_t;
            np.sigma_s = np.sigma_s + pi.sigma_s;
            float avg  = (pi.sigma_s.x + pi.sigma_s.y + pi.sigma_s.z) / 3.0f;
            cdf[num_overlapping] = (num_overlapping > 0 ? cdf[num_overlapping - 1] : 0.0f) + avg;
            num_overlapping++;
        }
        if (num_overlapping > 1 && !np.is_vaccum()) {
            float total_cdf = cdf[num_overlapping - 1];
            if (total_cdf > 0.0f)
                for (int i = 0; i < num_overlapping; i++)
                    cdf[i] /= total_cdf;
        }
        np.sigma_s = V3(std::min(np.sigma_s.x, np.sigma_t.x), std::min(np.sigma_s.y, np.sigma_t.y),
                        std::min(np.sigma_s.z, np.sigma_t.z));
        current_params = np;
    }
    static V3 vdiv(const V3& a, float d) { return V3(a.x / d, a.y / d, a.z / d); }   // Imath Vec3 / T
    static V3 transmittance(const V3& sigma_t, float distance)
    {
        return V3(expf(-sigma_t.x * distance), expf(-sigma_t.y * distance), expf(-sigma_t.z * distance));
    }
    // returns true when the path scattered inside the medium (new origin / direction in r)
    template<class SamplerT>
    bool integrate(Ray& r, SamplerT& sampler, float hit_t, V3& path_weight, float& bsdf_pdf)
    {
        if (depth <= 0 || current_params.is_vaccum())
            return false;
        const V3 ws = V3(path_weight.x * current_params.sigma_s.x / current_params.sigma_t.x,
                         path_weight.y * current_params.sigma_s.y / current_params.sigma_t.y,
                         path_weight.z * current_params.sigma_s.z / current_params.sigma_t.z);
        float cw[3] = { ws.x, ws.y, ws.z };
        float total = cw[0] + cw[1] + cw[2];
        if (total <= 0.0f) {
            path_weight = path_weight * transmittance(current_params.sigma_t, hit_t);
            return false;
        }
        float inv_total = 1.0f / total;
        cw[0] *= inv_total; cw[1] *= inv_total; cw[2] *= inv_total;
        V3 rnd = sampler.get();
        int channel;
        if (rnd.y < cw[0])
            channel = 0;
        else if (rnd.y < cw[0] + cw[1])
            channel = 1;
        else
            channel = 2;
        float sigma_t_channel = current_params.sigma_t[channel];
        float t_volume        = -logf(1.0f - rnd.x) / sigma_t_channel;
        bool scatter = t_volume < hit_t;
        float t      = scatter ? t_volume : hit_t;
        V3 tr        = transmittance(current_params.sigma_t, t);
        V3 density   = scatter ? (current_params.sigma_t * tr) : tr;
        float pdf    = density.x * cw[0] + density.y * cw[1] + density.z * cw[2];
        if (pdf <= 0.0f)
            return false;
        if (scatter)
            path_weight = path_weight * vdiv(tr * current_params.sigma_s, pdf);
        else {
            path_weight = path_weight * vdiv(tr, pdf);
            return false;
        }
        r.origin  = r.origin + r.direction * t_volume;
        int index = 0;
        if (num_overlapping > 1)
            for (index = 0; index < num_overlapping - 1; ++index)
                if (rnd.z < cdf[index])
                    break;
        int medium_index = overlapping_medium_indices[index];
        V3 rp            = sampler.get();
        BSample ps       = pool[mediums[medium_index]].sample_phase_func(-r.direction, rp.x, rp.y, rp.z);
        if (ps.pdf > 0.0f) {
            path_weight = path_weight * ps.weight;
            r.direction = ps.wi;
            bsdf_pdf    = ps.pdf;
            return true;
        }
        return false;
    }
    bool add_medium(const MediumParams& np)
    {
        if (depth >= MaxEntries || pool_size >= MaxEntries)
            return false;
        const int p = pool_size++;
        pool[p]     = np;
        int insert_pos = depth;
        for (int i = 0; i < depth; ++i)
            if (np.priority > pool[mediums[i]].priority) {
                insert_pos = i;
                break;
            }
        for (int j = depth; j > insert_pos; --j)
            mediums[j] = mediums[j - 1];
        mediums[insert_pos] = p;
        entry_order[depth]  = p;
        depth++;
        compute_current_params();
        return true;
    }
    void pop_medium()
    {
        if (depth <= 0)
            return;
        depth--;
        const int p      = entry_order[depth];
        int sorted_index = -1;
        for (int i = 0; i <= depth; ++i)
            if (mediums[i] == p) {
                sorted_index = i;
                break;
            }
        if (sorted_index < 0)
            return;
        for (int j = sorted_index; j < depth; ++j)
            mediums[j] = mediums[j + 1];
        if (p == pool_size - 1)
            pool_size--;
        compute_current_params();
    }
    bool false_intersection_with(const MediumParams& entrant) const
    {
        const MediumParams* current = get_current_params();
        if (!current)
            return false;
        if (entrant.is_special_priority() && current->is_special_priority())
            return false;
        if (entrant.priority == current->priority)
            return true;
        return entrant.priority > current->priority;
    }
};

struct ShadingResult {
    V3 Le = V3(0.0f);
    CompositeBSDF bsdf;
    MediumParams medium_data;   // what the surface encloses (medium_vdf / anisotropic_vdf closures)
};

inline const Clos* clos_ptr_param(const ClosComp* c, int word)
{
    const Clos* p;
    std::memcpy(&p, c->params + word, sizeof p);
    return p;
}
inline V3 clamp01(const V3& c)
{
    auto f = [](float x) { return x < 0.f ? 0.f : (x > 1.f ? 1.f : x); };
    return V3(f(c.x), f(c.y), f(c.z));
}
// mtx::SheenLobe construction from an MX_SHEEN_ID component; false: unsupported mode
inline bool sheen_from_component(Lobe& l, const ClosComp* comp, const SG& sg, float path_roughness)
{
    l.type   = LOBE_BSDL_SHEEN;
    l.N      = V3(comp->params[0], comp->params[1], comp->params[2]);
    l.albedo = V3(comp->params[3], comp->params[4], comp->params[5]);
    // "mode" keyword: 1 = Zeltner-Burley LTC sheen, anything else = Conty-Kulla (SheenLobe::use_zeltner)
    lobes::sheen_setup(l, -sg.I.val, comp->params[6], sg.backfacing != 0, path_roughness, (int)f2u(comp->params[7]));
    return true;
}
// mtx::ConductorLobe / DielectricLobe / SchlickLobe from their closure components (parameter
// order = the Data structs' registration order; the distribution string takes one word)
inline void mx_from_component(Lobe& l, const ClosComp* comp, const SG& sg, float path_roughness)
{
    const float* p = comp->params;
    const V3 wo    = -sg.I.val;
    l.type         = LOBE_MX_SPEC;
    l.N            = V3(p[0], p[1], p[2]);
    const V3 Z     = lobes::bsdl_visible_normal(wo, l.N, l.N);
    l.tf           = lobes::bsdl_frame_zx(Z, V3(p[3], p[4], p[5]));
    const float cosNO     = dot(Z, wo);
    const bool backfacing = sg.backfacing != 0;
    if (comp->id == MX_CONDUCTOR_ID)
        l.mx = mx_conductor_setup(cosNO, p[6], p[7], V3(p[8], p[9], p[10]), V3(p[11], p[12], p[13]), path_roughness);
    else if (comp->id == MX_DIELECTRIC_ID)
        l.mx = mx_dielectric_setup(cosNO, V3(p[6], p[7], p[8]), V3(p[9], p[10], p[11]), p[12], p[13], p[14],
                                   V3(p[18], p[19], p[20]), backfacing, path_roughness);
    else
        l.mx = mx_schlick_setup(cosNO, V3(p[6], p[7], p[8]), V3(p[9], p[10], p[11]), p[12], p[13],
                                V3(p[14], p[15], p[16]), V3(p[17], p[18], p[19]), p[20], backfacing, path_roughness);
}
// evaluate_layer_opacity (shading.cpp:1198-1282): how much of the light the top stack of a
// layer() takes; returns the weight held when the walk ends, as the reference does
inline V3 evaluate_layer_opacity(const SG& sg, float path_roughness, const Clos* closure)
{
    if (closure == nullptr)
        return V3(0.0f);
    const int STACK_SIZE = 16;
    int stack_idx        = 0;
    const Clos* ptr_stack[STACK_SIZE];
    V3 weight_stack[STACK_SIZE];
    V3 weight(1.0f);
    while (closure) {
        switch (closure->id) {
        case CL_MUL:
            weight  = weight * ((const ClosMul*)closure)->weight;
            closure = ((const ClosMul*)closure)->closure;
            break;
        case CL_ADD:
            ptr_stack[stack_idx]      = ((const ClosAdd*)closure)->b;
            weight_stack[stack_idx++] = weight;
            closure                   = ((const ClosAdd*)closure)->a;
            break;
        default: {
            const ClosComp* comp = (const ClosComp*)closure;
            V3 w                 = comp->w;
            switch (comp->id) {
            case MX_LAYER_ID:
                closure                   = clos_ptr_param(comp, 0);
                ptr_stack[stack_idx]      = clos_ptr_param(comp, 2);
                weight_stack[stack_idx++] = weight * w;
                break;
            case REFLECTION_ID:
            case FRESNEL_REFLECTION_ID: {
                Lobe l;
                l.type = LOBE_REFLECTION;
                l.N    = V3(comp->params[0], comp->params[1], comp->params[2]);
                l.eta  = comp->id == FRESNEL_REFLECTION_ID ? comp->params[3] : 0.0f;
                weight  = weight * (w * l.get_albedo(-sg.I.val));
                closure = nullptr;
                break;
            }
            case MX_SHEEN_ID: {
                Lobe l;
                sheen_from_component(l, comp, sg, path_roughness);
                weight  = weight * (w * (V3(1.0f) - V3(l.emiss)));
                closure = nullptr;
                break;
            }
            case MX_DIELECTRIC_ID: {
                Lobe l;
                mx_from_component(l, comp, sg, path_roughness);
                weight  = weight * (w * (V3(1.0f) - mx_filter_o(l.mx, false)));
                closure = nullptr;
                break;
            }
            case MX_GENERALIZED_SCHLICK_ID: {
                closure = nullptr;
                // transmissive dielectrics are opaque to the layer below
                if (!(comp->params[9] == 0 && comp->params[10] == 0 && comp->params[11] == 0))
                    break;
                Lobe l;
                mx_from_component(l, comp, sg, path_roughness);
                weight = weight * (w * (V3(1.0f) - mx_filter_o(l.mx, true)));
                break;
            }
            default: closure = nullptr; break;   // unhandled BSDFs are opaque
            }
        }
        }
        if (closure == nullptr && stack_idx > 0) {
            closure = ptr_stack[--stack_idx];
            weight  = weight_stack[stack_idx];
        }
    }
    return weight;
}

// process_medium_closure (shading.cpp:1283-1447): what the surface encloses, found before the
// BSDF pass so that nested dielectrics can be resolved against the medium stack
inline bool is_black3(const V3& c) { return c.x == 0 && c.y == 0 && c.z == 0; }
inline void process_medium_closure(const SG& sg, float path_roughness, ShadingResult& result,
                                   const MediumStack& medium_stack, const Clos* closure)
{
    if (!closure)
        return;
    const int STACK_SIZE = 16;
    int stack_idx        = 0;
    const Clos* ptr_stack[STACK_SIZE];
    V3 weight_stack[STACK_SIZE];
    V3 weight(1.0f);
    auto clamp_s = [&]() {
        MediumParams& m = result.medium_data;
        m.sigma_s = V3(std::min(m.sigma_s.x, m.sigma_t.x), std::min(m.sigma_s.y, m.sigma_t.y),
                       std::min(m.sigma_s.z, m.sigma_t.z));
    };
    while (closure) {
        switch (closure->id) {
        case CL_MUL:
            weight  = weight * ((const ClosMul*)closure)->weight;
            closure = ((const ClosMul*)closure)->closure;
            break;
        case CL_ADD:
            weight_stack[stack_idx] = weight;
            ptr_stack[stack_idx++]  = ((const ClosAdd*)closure)->b;
            closure                 = ((const ClosAdd*)closure)->a;
            break;
        default: {
            const ClosComp* comp = (const ClosComp*)closure;
            const float* p       = comp->params;
            closure              = nullptr;
            switch (comp->id) {
            case MX_LAYER_ID: {
                const Clos* top  = clos_ptr_param(comp, 0);
                const Clos* base = clos_ptr_param(comp, 2);
                V3 base_w = weight * (V3(1.0f) - clamp01(evaluate_layer_opacity(sg, path_roughness, top)));
                closure                   = top;
                ptr_stack[stack_idx]      = base;
                weight_stack[stack_idx++] = weight * base_w;   // (sic) the weight enters twice
                break;
            }
            case MX_ANISOTROPIC_VDF_ID: {
                // params: albedo, extinction, anisotropy
                V3 cw                       = weight * comp->w;
                result.medium_data.sigma_t  = cw * V3(p[3], p[4], p[5]);
                result.medium_data.sigma_s  = V3(p[0], p[1], p[2]) * result.medium_data.sigma_t;
                result.medium_data.medium_g = p[6];
                result.medium_data.priority = 0;
                clamp_s();
                break;
            }
            case MX_MEDIUM_VDF_ID: {
                // params: albedo, transmission_depth, transmission_color, anisotropy, ior, priority
                V3 cw = weight * comp->w;
                const V3 albedo(p[0], p[1], p[2]), t_color(p[4], p[5], p[6]);
                if (is_black3(albedo) && is_black3(t_color)) {
                    result.medium_data.sigma_t = V3(0.0f);
                    result.medium_data.sigma_s = V3(0.0f);
                } else {
                    const float epsilon = 1e-10f;
                    V3 st(-fast_log(fmaxf(t_color.x, epsilon)), -fast_log(fmaxf(t_color.y, epsilon)),
                          -fast_log(fmaxf(t_color.z, epsilon)));
                    result.medium_data.sigma_t = st * MediumStack::vdiv(cw, p[3]);
                    result.medium_data.sigma_s = albedo * result.medium_data.sigma_t;
                    clamp_s();
                }
                result.medium_data.medium_g       = p[7];
                result.medium_data.refraction_ior = sg.backfacing ? 1.0f / p[8] : p[8];
                result.medium_data.priority       = (int)f2u(p[9]);
                break;
            }
            case MX_DIELECTRIC_ID:
            case MX_GENERALIZED_SCHLICK_ID: {
                // refr_tint: dielectric params[9..11], schlick params[9..11] as well (after N, U, refl_tint)
                if (!is_black3(weight * comp->w * V3(p[9], p[10], p[11]))) {
                    float ior;
                    if (comp->id == MX_DIELECTRIC_ID)
                        ior = p[14];
                    else {
                        // F0 = params[14..16]
                        float avg_F0  = mx_clamp((p[14] + p[15] + p[16]) / 3.0f, 0.0f, 0.99f);
                        float sqrt_F0 = sqrtf(avg_F0);
                        ior           = (1 + sqrt_F0) / (1 - sqrt_F0);
                    }
                    result.medium_data.refraction_ior = sg.backfacing ? 1.0f / ior : ior;
                    const MediumParams* current       = medium_stack.get_current_params();
                    if (current && result.medium_data.priority <= current->priority)
                        result.medium_data.refraction_ior = current->refraction_ior;
                }
                break;
            }
            default: break;
            }
            break;
        }
        }
        if (closure == nullptr && stack_idx > 0) {
            closure = ptr_stack[--stack_idx];
            weight  = weight_stack[stack_idx];
        }
    }
}

// process_bsdf_closure: explicit 16-deep stack, weights multiplied root->leaf
inline void process_closure(const SG& sg, ShadingResult& result, const Clos* closure, bool light_only,
                            float path_roughness = 0.0f, const MediumStack* medium_stack = nullptr)
{
    if (!closure)
        return;
    if (!light_only && medium_stack)
        process_medium_closure(sg, path_roughness, result, *medium_stack, closure);
    const int STACK_SIZE = 16;
    int stack_idx        = 0;
    const Clos* ptr_stack[STACK_SIZE];
    V3 weight_stack[STACK_SIZE];
    V3 weight(1.0f);
    while (closure) {
        switch (closure->id) {
        case CL_MUL:
            weight  = weight * ((const ClosMul*)closure)->weight;
            closure = ((const ClosMul*)closure)->closure;
            break;
        case CL_ADD:
            ptr_stack[stack_idx]      = ((const ClosAdd*)closure)->b;
            weight_stack[stack_idx++] = weight;
            closure                   = ((const ClosAdd*)closure)->a;
            break;
        default: {
            const ClosComp* comp = (const ClosComp*)closure;
            V3 cw                = weight * comp->w;
            closure              = nullptr;
            if (comp->id == EMISSION_ID)
                result.Le = result.Le + cw;
            else if (comp->id == MX_UNIFORM_EDF_ID)
                result.Le = result.Le + cw * V3(comp->params[0], comp->params[1], comp->params[2]);
            else if (!light_only) {
                Lobe l;
                l.N   = V3(comp->params[0], comp->params[1], comp->params[2]);
                l.eta = 0.0f;
                bool known = true;
                switch (comp->id) {
                case DIFFUSE_ID: l.type = LOBE_DIFFUSE; break;
                case TRANSLUCENT_ID: l.type = LOBE_TRANSLUCENT; l.N = -l.N; break;
                case REFLECTION_ID: l.type = LOBE_REFLECTION; break;
                case FRESNEL_REFLECTION_ID: l.type = LOBE_REFLECTION; l.eta = comp->params[3]; break;
                case REFRACTION_ID: l.type = LOBE_REFRACTION; l.eta = comp->params[3]; break;
                case TRANSPARENT_ID:
                case MX_TRANSPARENT_ID: l.type = LOBE_TRANSPARENT; break;
                case OREN_NAYAR_ID: {
                    // -> MxOrenNayarDiffuse{N, albedo 1, sigma, no energy compensation}
                    // (shading.cpp:1496-1503); frame on the visible normal, Nf = Ngf = N
                    l.type = LOBE_BSDL_OREN_NAYAR;
                    l.ax   = lobes::bsdl_clamp(comp->params[3], 0.0f, 1.0f);
                    l.tf   = TangentFrame::from_normal(lobes::bsdl_visible_normal(-sg.I.val, l.N, l.N));
                    break;
                }
                case MX_OREN_NAYAR_DIFFUSE_ID:
                case MX_BURLEY_DIFFUSE_ID: {
                    // params: N, albedo, roughness [, energy_compensation] (libbsdl Data structs)
                    l.type   = comp->id == MX_BURLEY_DIFFUSE_ID ? LOBE_BSDL_BURLEY : LOBE_BSDL_OREN_NAYAR;
                    l.albedo = V3(comp->params[3], comp->params[4], comp->params[5]);
                    l.ax     = lobes::bsdl_clamp(comp->params[6], 0.0f, 1.0f);
                    if (comp->id == MX_OREN_NAYAR_DIFFUSE_ID)
                        l.energy_compensation = f2u(comp->params[7]) != 0;
                    l.tf = TangentFrame::from_normal(lobes::bsdl_visible_normal(-sg.I.val, l.N, l.N));
                    break;
                }
                case MX_SHEEN_ID: known = sheen_from_component(l, comp, sg, path_roughness); break;
                case MX_CONDUCTOR_ID:
                case MX_DIELECTRIC_ID:
                case MX_GENERALIZED_SCHLICK_ID:
                    // a boundary the medium stack says is not there (nested dielectrics) is passed through
                    if (comp->id != MX_CONDUCTOR_ID && medium_stack
                        && medium_stack->false_intersection_with(result.medium_data))
                        l.type = LOBE_TRANSPARENT;
                    else
                        mx_from_component(l, comp, sg, path_roughness);
                    break;
                case MX_TRANSLUCENT_ID: {
                    // params: N, albedo (bsdf_translucent_impl.h): cosine lobe on the far side
                    l.type   = LOBE_MX_TRANSLUCENT;
                    l.albedo = V3(comp->params[3], comp->params[4], comp->params[5]);
                    l.tf     = TangentFrame::from_normal(lobes::bsdl_visible_normal(-sg.I.val, l.N, l.N));
                    break;
                }
                case SPI_THINLAYER: {
                    // params: N, T, IOR, roughness, anisotropy, thickness, refl_tint, refr_tint, sigma_t
                    // (ThinLayerLobe::Data registration order; shading.cpp:1668-1674)
                    const float* p = comp->params;
                    const V3 wo    = -sg.I.val;
                    l.type         = LOBE_SPI_THINLAYER;
                    const V3 Z     = lobes::bsdl_visible_normal(wo, l.N, l.N);
                    l.tf           = lobes::bsdl_frame_zx(Z, V3(p[3], p[4], p[5]));
                    l.thin = lobes::thin_setup(dot(wo, Z), p[6], p[7], p[8], p[9], V3(p[10], p[11], p[12]),
                                               V3(p[13], p[14], p[15]), V3(p[16], p[17], p[18]), path_roughness);
                    break;
                }
                case MX_SUBSURFACE_ID: {
                    // no BSSRDF in testrender: a diffuse lobe weighted by the albedo (shading.cpp:1626-1635)
                    l.type = LOBE_DIFFUSE;
                    cw     = cw * V3(comp->params[3], comp->params[4], comp->params[5]);
                    break;
                }
                case MX_LAYER_ID: {
                    // layer(top, base): the base is attenuated by what the top stack takes
                    // (shading.cpp:1645-1661)
                    const Clos* top  = clos_ptr_param(comp, 0);
                    const Clos* base = clos_ptr_param(comp, 2);
                    V3 base_w = weight * (V3(1.0f) - clamp01(evaluate_layer_opacity(sg, path_roughness, top)));
                    closure   = top;
                    weight    = cw;
                    if (!(base_w.x == 0 && base_w.y == 0 && base_w.z == 0)) {
                        ptr_stack[stack_idx]      = base;
                        weight_stack[stack_idx++] = base_w;
                    }
                    known = false;   // nothing to add for the layer node itself
                    break;
                }
                case PHONG_ID: l.type = LOBE_PHONG; l.exponent = comp->params[3]; break;
                case WARD_ID:
                    l.type = LOBE_WARD;
                    l.T    = V3(comp->params[3], comp->params[4], comp->params[5]);
                    l.ax   = comp->params[6];
                    l.ay   = comp->params[7];
                    l.tf   = TangentFrame::from_normal_and_tangent(l.N, l.T);
                    break;
                case MICROFACET_ID: {
                    // params: dist code, N, U, xalpha, yalpha, eta, refract (shading.h MicrofacetParams)
                    int dist = (int)f2u(comp->params[0]);
                    l.type   = LOBE_MICROFACET;
                    l.N      = V3(comp->params[1], comp->params[2], comp->params[3]);
                    l.T      = V3(comp->params[4], comp->params[5], comp->params[6]);
                    l.ax     = comp->params[7];
                    l.ay     = comp->params[8];
                    l.eta    = comp->params[9];
                    l.refract = (int)f2u(comp->params[10]);
                    l.ggx    = dist == 1;
                    l.tf     = TangentFrame::from_normal_and_tangent(l.N, l.T);
                    known    = (dist == 1 || dist == 2 || dist == 3) && l.refract >= 0 && l.refract <= 2;
                    break;
                }
                default: known = false; break;
                }
                if (known)
                    result.bsdf.add(cw, l);
            }
            break;
        }
        }
        if (closure == nullptr && stack_idx > 0) {
            closure = ptr_stack[--stack_idx];
            weight  = weight_stack[stack_idx];
        }
    }
}

// ---- scene queries ----------------------------------------------------------------
struct Intersection {
    float t, u, v;
    unsigned id;
};
inline float compf(const V3& v, int i) { return (&v.x)[i]; }
inline float minf_(float a, float b) { return b < a ? b : a; }
inline float maxf_(float a, float b) { return b > a ? b : a; }
inline bool box_intersect(const V3& org, const V3& rdir, float tmax, const float* bounds, float* dist)
{
    const float tx1 = (bounds[0] - org.x) * rdir.x, tx2 = (bounds[1] - org.x) * rdir.x;
    const float ty1 = (bounds[2] - org.y) * rdir.y, ty2 = (bounds[3] - org.y) * rdir.y;
    const float tz1 = (bounds[4] - org.z) * rdir.z, tz2 = (bounds[5] - org.z) * rdir.z;
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
inline float xorf(float a, unsigned b) { return u2f(f2u(a) ^ b); }

inline Intersection scene_intersect(const RenderScene& S, const Ray& ray, const float tmax, unsigned skipID1,
                                    unsigned skipID2 = ~0u)
{
    struct StackItem {
        int node;
        float dist;
    } stack[64];
    Intersection result;
    result.t  = tmax;
    result.u  = result.v = 0;
    result.id = 0;
    stack[0]  = { 0, result.t };
    const V3 org = ray.origin, dir = ray.direction;
    const V3 rdir(1 / dir.x, 1 / dir.y, 1 / dir.z);
    int kz = 0;
    if (std::fabs(dir.y) > std::fabs(compf(dir, kz)))
        kz = 1;
    if (std::fabs(dir.z) > std::fabs(compf(dir, kz)))
        kz = 2;
    int kx = kz == 2 ? 0 : kz + 1;
    int ky = kx == 2 ? 0 : kx + 1;
    const V3 shearDir(compf(dir, kx) / compf(dir, kz), compf(dir, ky) / compf(dir, kz), compf(rdir, kz));
    for (int stackPtr = 1; stackPtr != 0;) {
        if (result.t < stack[--stackPtr].dist)
            continue;
        const float* node = S.bvh_nodes + 8 * stack[stackPtr].node;
        unsigned child = f2u(node[6]), nprims = f2u(node[7]);
        if (nprims) {
            for (unsigned i = 0; i < nprims; i++) {
                unsigned id = S.bvh_indices[child + i];
                const V3 A = vert(S, S.triangles[3 * id]) - org, B = vert(S, S.triangles[3 * id + 1]) - org,
                         C = vert(S, S.triangles[3 * id + 2]) - org;
                const float Ax = compf(A, kx) - shearDir.x * compf(A, kz);
                const float Ay = compf(A, ky) - shearDir.y * compf(A, kz);
                const float Bx = compf(B, kx) - shearDir.x * compf(B, kz);
                const float By = compf(B, ky) - shearDir.y * compf(B, kz);
                const float Cx = compf(C, kx) - shearDir.x * compf(C, kz);
                const float Cy = compf(C, ky) - shearDir.y * compf(C, kz);
                const float U = Cx * By - Cy * Bx, V = Ax * Cy - Ay * Cx, W = Bx * Ay - By * Ax;
                if ((U < 0 || V < 0 || W < 0) && (U > 0 || V > 0 || W > 0))
                    continue;
                const float det = U + V + W;
                if (det == 0)
                    continue;
                const float Az = compf(A, kz), Bz = compf(B, kz), Cz = compf(C, kz);
                const float T       = shearDir.z * (U * Az + V * Bz + W * Cz);
                const unsigned mask = f2u(det) & 0x80000000u;
                if (xorf(T, mask) < 0)
                    continue;
                if (xorf(T, mask) > result.t * xorf(det, mask))
                    continue;
                if (id == skipID1 || id == skipID2)
                    continue;
                const float rcpDet = 1 / det;
                result.t  = T * rcpDet;
                result.u  = V * rcpDet;
                result.v  = W * rcpDet;
                result.id = id;
            }
        } else {
            int child1 = (int)child, child2 = child1 + 1;
            float dist1 = 0, dist2 = 0;
            bool hit1 = box_intersect(org, rdir, result.t, S.bvh_nodes + 8 * child1, &dist1);
            bool hit2 = box_intersect(org, rdir, result.t, S.bvh_nodes + 8 * child2, &dist2);
            if (dist1 > dist2) {
                std::swap(hit1, hit2);
                std::swap(dist1, dist2);
                std::swap(child1, child2);
            }
            stack[stackPtr] = { child2, dist2 };
            stackPtr += hit2;
            stack[stackPtr] = { child1, dist1 };
            stackPtr += hit1;
        }
    }
    return result;
}

struct LightSample {
    V3 dir;
    float dist, pdf, u, v;
};
inline LightSample scene_sample(const RenderScene& S, int primID, const V3& x, float xi, float yi)
{
    if (yi > xi) {
        xi *= 0.5f;
        yi -= xi;
    } else {
        yi *= 0.5f;
        xi -= yi;
    }
    const V3 va = vert(S, S.triangles[3 * primID]), vb = vert(S, S.triangles[3 * primID + 1]),
             vc = vert(S, S.triangles[3 * primID + 2]);
    const V3 n  = cross(va - vb, va - vc);
    V3 l        = ((1 - xi - yi) * va + xi * vb + yi * vc) - x;
    float d2    = length2(l);
    V3 dir      = normalized(l);
    float pdf   = d2 / (0.5f * std::fabs(dot(dir, n)));
    return { dir, std::sqrt(d2), pdf, xi, yi };
}
inline float scene_shapepdf(const RenderScene& S, int primID, const V3& x, const V3& p)
{
    const V3 va = vert(S, S.triangles[3 * primID]), vb = vert(S, S.triangles[3 * primID + 1]),
             vc = vert(S, S.triangles[3 * primID + 2]);
    const V3 n  = cross(va - vb, va - vc);
    V3 l        = p - x;
    float d2    = length2(l);
    V3 dir      = normalized(l);
    return d2 / (0.5f * std::fabs(dot(dir, n)));
}
inline V3 scene_normal(const RenderScene& S, V3& Ng, int primID, float u, float v)
{
    const V3 va = vert(S, S.triangles[3 * primID]), vb = vert(S, S.triangles[3 * primID + 1]),
             vc = vert(S, S.triangles[3 * primID + 2]);
    Ng = normalized(cross(va - vb, va - vc));
    if (S.n_triangles[3 * primID] < 0)
        return Ng;
    const V3 na = nrm(S, S.n_triangles[3 * primID]), nb = nrm(S, S.n_triangles[3 * primID + 1]),
             nc = nrm(S, S.n_triangles[3 * primID + 2]);
    return normalized((1 - u - v) * na + u * nb + v * nc);
}
inline void scene_project(Dv& p, const V3& N, const V3& I)
{
    V3 nI      = normalized(I);
    float cosI = dot(-nI, N);
    if (std::fabs(cosI) > 1e-3f) {
        float deltaX = dot(p.dx, N) / cosI;
        float deltaY = dot(p.dy, N) / cosI;
        p.dx = p.dx + nI * deltaX;
        p.dy = p.dy + nI * deltaY;
    }
}
struct V2 {
    float x, y;
};
// Scene::uv (raytracer.h:263-300): returns uv value + screen derivatives
inline void scene_uv(const RenderScene& S, const Dv& p, const V3& n, V3& dPdu, V3& dPdv, int primID, float u,
                     float v, V2& uvv, V2& uvdx, V2& uvdy)
{
    uvv = uvdx = uvdy = { 0, 0 };
    if (S.uv_triangles[3 * primID] < 0)
        return;
    auto UV = [&](int i) { return V2 { S.uvs[2 * i], S.uvs[2 * i + 1] }; };
    const V2 ta = UV(S.uv_triangles[3 * primID]), tb = UV(S.uv_triangles[3 * primID + 1]),
             tc = UV(S.uv_triangles[3 * primID + 2]);
    const V3 va = vert(S, S.triangles[3 * primID]), vb = vert(S, S.triangles[3 * primID + 1]),
             vc = vert(S, S.triangles[3 * primID + 2]);
    const V2 dt02 = { ta.x - tc.x, ta.y - tc.y }, dt12 = { tb.x - tc.x, tb.y - tc.y };
    const V3 dp02 = va - vc, dp12 = vb - vc;
    const float det = dt02.x * dt12.y - dt02.y * dt12.x;
    if (det != 0) {
        float invdet = 1 / det;
        dPdu         = (dt12.y * dp02 - dt02.y * dp12) * invdet;
        dPdv         = (-dt12.x * dp02 + dt02.x * dp12) * invdet;
    }
    V3 La = cross(n, vc - vb);
    La    = La / dot(va - vb, La);
    V3 Lb = cross(n, va - vc);
    Lb    = Lb / dot(vb - vc, Lb);
    V3 Lc = cross(n, vb - va);
    Lc    = Lc / dot(vc - va, Lc);
    auto comb = [&](float a, float b, float c) {
        return V2 { a * ta.x + b * tb.x + c * tc.x, a * ta.y + b * tb.y + c * tc.y };
    };
    uvdx = comb(dot(La, p.dx), dot(Lb, p.dx), dot(Lc, p.dx));
    uvdy = comb(dot(La, p.dy), dot(Lb, p.dy), dot(Lc, p.dy));
    uvv  = comb(1 - u - v, u, v);
}

inline void globals_from_hit(const RenderScene& S, SG& sg, const Ray& r, float t, int id, float u, float v)
{
    Dv direction = r.dual_direction();
    sg.I         = direction;
    Dv P         = r.point_dual(t);
    V3 Ng;
    sg.N = scene_normal(S, Ng, id, u, v);
    scene_project(P, sg.N, sg.I.val);
    sg.P = P;
    V3 dPdu(0.0f), dPdv(0.0f);
    V2 uv, uvdx, uvdy;
    scene_uv(S, P, sg.N, dPdu, dPdv, id, u, v, uv, uvdx, uvdy);
    sg.dPdu = dPdu;
    sg.dPdv = dPdv;
    sg.u    = Df(uv.x, uvdx.x, uvdy.x);
    sg.v    = Df(uv.y, uvdx.y, uvdy.y);
    sg.surfacearea = S.mesh_surfacearea[S.meshids[id]];
    sg.backfacing  = dot(Ng, sg.I.val) > 0;
    if (sg.backfacing) {
        sg.N = -sg.N;
        Ng   = -Ng;
    }
    sg.Ng             = Ng;
    sg.raytype        = r.raytype;
    sg.flipHandedness = dot(cross(sg.P.dx, sg.P.dy), sg.N) < 0;
    sg.dPdz = V3(0.0f);
    sg.time = sg.dtime = 0.0f;
    sg.dPdtime = V3(0.0f);
    sg.Ps      = Dv(V3(0.0f));
}

typedef void (*ShaderFn)(SG& sg);

// process_background_closure (shading.cpp:1709-1746): walks the tree and returns the
// weight it holds when the walk ends (the last visited branch), as the reference does.
inline V3 process_background_closure(const Clos* closure)
{
    if (!closure)
        return V3(0.0f);
    const int STACK_SIZE = 16;
    int stack_idx        = 0;
    const Clos* ptr_stack[STACK_SIZE];
    V3 weight_stack[STACK_SIZE];
    V3 weight(1.0f);
    while (closure) {
        switch (closure->id) {
        case CL_MUL:
            weight  = weight * ((const ClosMul*)closure)->weight;
            closure = ((const ClosMul*)closure)->closure;
            break;
        case CL_ADD:
            ptr_stack[stack_idx]      = ((const ClosAdd*)closure)->b;
            weight_stack[stack_idx++] = weight;
            closure                   = ((const ClosAdd*)closure)->a;
            break;
        case BACKGROUND_ID:
            weight  = weight * ((const ClosComp*)closure)->w;
            closure = nullptr;
            break;
        default:
            // the reference's switch has no default: an unexpected component would spin
            // forever there; treat it as the end of this branch
            closure = nullptr;
            break;
        }
        if (closure == nullptr && stack_idx > 0) {
            closure = ptr_stack[--stack_idx];
            weight  = weight_stack[stack_idx];
        }
    }
    return weight;
}

// Importance table over the sphere of directions (background.h:38-276)
struct Background {
    std::vector<V3> values;
    std::vector<float> rows, cols;
    int res = -1;
    float invres = 0.0f, invjacobian = 0.0f;

    Dv map(float x, float y) const
    {
        Df u     = Df(x, 1, 0) * invres;
        Df v     = Df(y, 0, 1) * invres;
        Df theta = u * float(2 * M_PI);
        float s, c;
        fast_sincos(theta.val, &s, &c);
        Df st = dualfunc(theta, s, c), ct = dualfunc(theta, c, -s);
        Df cos_phi = 1.0f - 2.0f * v;
        Df sin_phi = d_sqrt(1.0f - cos_phi * cos_phi);
        return make_dv(sin_phi * ct, sin_phi * st, cos_phi);
    }
    template<class F> void prepare(int resolution, F cb)
    {
        res = resolution;
        if (res < 32)
            res = 32;
        invres      = 1.0f / res;
        invjacobian = res * res / float(4 * M_PI);
        values.assign((size_t)res * res, V3(0.0f));
        rows.assign(res, 0.0f);
        cols.assign((size_t)res * res, 0.0f);
        for (int y = 0, i = 0; y < res; y++) {
            for (int x = 0; x < res; x++, i++) {
                values[i] = cb(map(x + 0.5f, y + 0.5f));
                cols[i]   = std::max(std::max(values[i].x, values[i].y), values[i].z) + ((x > 0) ? cols[i - 1] : 0.0f);
            }
            rows[y] = cols[i - 1] + ((y > 0) ? rows[y - 1] : 0.0f);
            if (cols[i - 1] > 0)
                for (int x = 0; x < res; x++)
                    cols[i - res + x] /= cols[i - 1];
        }
        for (int y = 0; y < res; y++)
            rows[y] /= rows[res - 1];
        for (int y = 0, i = 0; y < res; y++) {
            float row_pdf = rows[y] - (y > 0 ? rows[y - 1] : 0.0f);
            for (int x = 0; x < res; x++, i++) {
                float col_pdf = cols[i] - (x > 0 ? cols[i - 1] : 0.0f);
                values[i]     = values[i] / (row_pdf * col_pdf * invjacobian);
            }
        }
    }
    V3 eval(const V3& dir, float& pdf) const
    {
        float u = fast_atan2(dir.y, dir.x) * float(M_1_PI * 0.5f);
        if (u < 0)
            u++;
        float v = (1 - dir.z) * 0.5f;
        int x   = (int)(u * res);
        if (x < 0) x = 0;
        else if (x >= res) x = res - 1;
        int y = (int)(v * res);
        if (y < 0) y = 0;
        else if (y >= res) y = res - 1;
        int i         = y * res + x;
        float row_pdf = rows[y] - (y > 0 ? rows[y - 1] : 0.0f);
        float col_pdf = cols[i] - (x > 0 ? cols[i - 1] : 0.0f);
        pdf           = std::max(0.0f, row_pdf * col_pdf * invjacobian);
        return values[i];
    }
    static float sample_cdf(const float* data, unsigned n, float x, unsigned* idx, float* pdf)
    {
        // upper_bound (background.h:14-34)
        const float* first = data;
        int len            = (int)n;
        while (len != 0) {
            int l2         = len / 2;
            const float* m = first + l2;
            if (x < *m)
                len = l2;
            else {
                first = m + 1;
                len -= l2 + 1;
            }
        }
        *idx = (unsigned)(first - data);
        float scaled_sample;
        if (*idx == 0) {
            *pdf          = data[0];
            scaled_sample = x / data[0];
        } else {
            *pdf          = data[*idx] - data[*idx - 1];
            scaled_sample = (x - data[*idx - 1]) / (data[*idx] - data[*idx - 1]);
        }
        return std::min(scaled_sample, 0.99999994f);
    }
    V3 sample(float rx, float ry, Dv& dir, float& pdf) const
    {
        float row_pdf, col_pdf;
        unsigned x, y;
        ry  = sample_cdf(rows.data(), res, ry, &y, &row_pdf);
        rx  = sample_cdf(cols.data() + (size_t)y * res, res, rx, &x, &col_pdf);
        dir = map(x + rx, y + ry);
        pdf = std::max(0.0f, row_pdf * col_pdf * invjacobian);
        return values[(size_t)y * res + x];
    }
};

struct Renderer {
    const RenderScene& S;
    const ShaderFn* shaders;
    Ctx* ctx;
    const Background* background = nullptr;   // importance table, when S.background_resolution > 0

    // SimpleRaytracer::eval_background (simpleraytracer.cpp:937-954)
    V3 eval_background(const Dv& dir, int bounce, ClosurePool& pool) const
    {
        SG sg;
        std::memset((void*)&sg, 0, sizeof(SG));
        sg.I = dir;
        if (bounce >= 0)
            sg.raytype = bounce > 0 ? RAY_DIFFUSE : RAY_CAMERA;
        execute(S.background_shader, sg, pool);
        return process_background_closure(sg.Ci);
    }

    void execute(int shaderID, SG& sg, ClosurePool& pool) const
    {
        pool.reset();
        sg.pool = &pool;
        sg.Ci   = nullptr;
        sg.ctx  = ctx;
        sg.shadeindex = 0;
        shaders[shaderID](sg);
    }

    V3 subpixel_radiance(float x, float y, Sampler& sampler) const
    {
        const float inf = std::numeric_limits<float>::infinity();
        Ray r           = camera_ray(S, x, y);
        V3 path_weight(1.0f), path_radiance(0.0f);
        int prev_id    = -1;
        float bsdf_pdf = inf;
        ClosurePool pool, light_pool;
        MediumStack medium_stack;
        for (int b = 0; b <= S.max_bounces; b++) {
            SG sg;
            Intersection hit = scene_intersect(S, r, inf, (unsigned)prev_id);
            if (hit.t == inf) {
                if (S.background_shader >= 0) {
                    if (b > 0 && background) {
                        float bg_pdf = 0;
                        V3 bg        = background->eval(r.direction, bg_pdf);
                        path_radiance = path_radiance
                                        + path_weight * bg * power_heuristic<WEIGHT_WEIGHT>(bsdf_pdf, bg_pdf);
                    } else {
                        path_radiance = path_radiance + path_weight * eval_background(Dv(r.direction), b, pool);
                    }
                }
                break;
            }
            if (medium_stack.integrate(r, sampler, hit.t, path_weight, bsdf_pdf))
                continue;   // scattered inside the medium: a bounce without a surface
            globals_from_hit(S, sg, r, hit.t, hit.id, hit.u, hit.v);
            if (S.show_globals) {
                V3 v = sg.Ng;
                if (S.show_globals == 2) v = sg.N;
                if (S.show_globals == 3) v = normalized(sg.dPdu);
                if (S.show_globals == 4) v = normalized(sg.dPdv);
                if (S.show_globals == 5) v = V3(sg.u.val, sg.v.val, 0);
                V3 c = v;
                if (S.show_globals != 5)
                    c = c * 0.5f + V3(0.5f);
                path_radiance = path_radiance + path_weight * c;
                break;
            }
            const float radius = r.radius + r.spread * hit.t;
            int shaderID       = S.shaderids[hit.id];
            if (shaderID < 0)
                break;
            execute(shaderID, sg, pool);
            ShadingResult result;
            bool last_bounce = b == S.max_bounces;
            process_closure(sg, result, sg.Ci, last_bounce, r.roughness, &medium_stack);
            const int nlights = S.nlightprims;
            float k           = 1;
            if (S.shader_is_light[shaderID] && nlights > 0) {
                const float light_pick_pdf = 1.0f / nlights;
                float light_pdf = light_pick_pdf * scene_shapepdf(S, hit.id, r.origin, sg.P.val);
                k               = power_heuristic<WEIGHT_EVAL>(bsdf_pdf, light_pdf);
            }
            path_radiance = path_radiance + path_weight * k * result.Le;
            if (last_bounce)
                break;
            result.bsdf.prepare(-sg.I.val, path_weight, b >= S.rr_depth);
            V3 s     = sampler.get();
            float xi = s.x, yi = s.y, zi = s.z;
            if (background) {
                // one shadow ray towards an importance-sampled background direction
                Dv bg_dir;
                float bg_pdf = 0;
                V3 bg        = background->sample(xi, yi, bg_dir, bg_pdf);
                BSample bs   = result.bsdf.eval(-sg.I.val, bg_dir.val);
                V3 contrib   = path_weight * bs.weight * bg * power_heuristic<WEIGHT_WEIGHT>(bg_pdf, bs.pdf);
                if ((contrib.x + contrib.y + contrib.z) > 0) {
                    Ray shadow_ray { sg.P.val, bg_dir.val, radius, 0, 0, RAY_SHADOW };
                    Intersection sh = scene_intersect(S, shadow_ray, inf, hit.id);
                    if (sh.t == inf)
                        path_radiance = path_radiance + contrib;
                }
            }
            if (nlights > 0) {
                const float light_pick_pdf = 1.0f / nlights;
                float xl = xi * nlights;
                int ls   = (int)std::floor(xl);
                xl -= ls;
                unsigned lid = S.lightprims[ls];
                if (lid != hit.id) {
                    int lshader       = S.shaderids[lid];
                    LightSample sample = scene_sample(S, lid, sg.P.val, xl, yi);
                    BSample bs         = result.bsdf.eval(-sg.I.val, sample.dir);
                    V3 contrib = path_weight * bs.weight
                                 * power_heuristic<EVAL_WEIGHT>(light_pick_pdf * sample.pdf, bs.pdf);
                    if ((contrib.x + contrib.y + contrib.z) > 0) {
                        Ray shadow_ray { sg.P.val, sample.dir, radius, 0, 0, RAY_SHADOW };
                        Intersection sh = scene_intersect(S, shadow_ray, sample.dist, hit.id, lid);
                        if (sh.t == sample.dist) {
                            SG lsg;
                            globals_from_hit(S, lsg, shadow_ray, sample.dist, lid, sample.u, sample.v);
                            execute(lshader, lsg, light_pool);
                            ShadingResult lres;
                            process_closure(lsg, lres, lsg.Ci, true, r.roughness);
                            path_radiance = path_radiance + contrib * lres.Le;
                        }
                    }
                }
            }
            BSample p   = result.bsdf.sample(-sg.I.val, xi, yi, zi);
            path_weight = path_weight * p.weight;
            bsdf_pdf    = p.pdf;
            r.raytype   = RAY_DIFFUSE;
            r.direction = p.wi;
            r.radius    = radius;
            r.spread    = std::max(r.spread, p.roughness);
            r.roughness = p.roughness;
            if (dot(sg.Ng, p.wi) < 0) {   // the sampled direction crosses the surface
                if (!sg.backfacing)
                    medium_stack.add_medium(result.medium_data);
                else
                    medium_stack.pop_medium();
            }
            if (!(path_weight.x > 0) && !(path_weight.y > 0) && !(path_weight.z > 0))
                break;
            prev_id  = hit.id;
            r.origin = sg.P.val;
        }
        return path_radiance;
    }

    V3 antialias_pixel(int x, int y) const
    {
        V3 result(0.0f);
        for (int si = 0, n = S.aa * S.aa; si < n; si++) {
            Sampler sampler(x, y, si);
            V3 j = S.no_jitter ? V3(0.5f, 0.5f, 0) : sampler.get();
            j.x *= 2;
            j.x = j.x < 1 ? std::sqrt(j.x) - 1 : 1 - std::sqrt(2 - j.x);
            j.y *= 2;
            j.y = j.y < 1 ? std::sqrt(j.y) - 1 : 1 - std::sqrt(2 - j.y);
            V3 r     = subpixel_radiance(x + 0.5f + j.x, y + 0.5f + j.y, sampler);
            float t  = 1.0f / (si + 1);
            result   = result * (1.0f - t) + r * t;
        }
        return result;
    }
};

}  // namespace oslo
