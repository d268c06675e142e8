// osl_b200_mxlobes.cuh — the MaterialX microfacet closures of the wavefront integrator (product code).
//
// conductor_bsdf, dielectric_bsdf, generalized_schlick_bsdf and translucent_bsdf are libbsdl lobes
// in the reference, wrapped by testrender's BSDL_WRAP (src/testrender/shading.cpp:73-115) with
// three RGB channels and OIIO fast_* math (src/testrender/bsdl_config.h):
//   GGXDist, TabulatedEnergyCurve, eval/sample_turquin_microms_reflection
//                                    src/libbsdl/include/BSDL/microfacet_tools_{decl,impl}.h
//   mtx::ConductorFresnel, ConductorLobe        BSDL/MTX/bsdf_conductor_impl.h
//   mtx::DielectricFresnel, DielectricBSDF, DielectricLobe   BSDL/MTX/bsdf_dielectric_impl.h
//   mtx::SchlickFresnel, SchlickLobe            BSDL/MTX/bsdf_schlick_impl.h
//   reflect, refract                            BSDL/tools.h
// A lobe is set up once per shading point (the reference constructs the lobe object inside
// CompositeBSDF's pool): MxSpec below is that state.  The energy-compensation tables are DATA
// (openshadinglanguage_b200/data/bsdl_luts.bin, baked by tools/bake_bsdl_luts.cpp and equal to the
// reference's genluts output): 99 KB in HBM behind RenderScene::bsdl_luts, read through the
// read-only cache - every thread of a material-sorted warp touches the same few rows.
// Included from osl_b200_lobes.cuh when the scene's materials create these closures
// (OSLD_MX_LOBES from the host code generator).
#pragma once

namespace osld {


// tools.h: CLAMP = MIN(MAX(x, a), b), LERP clamps its parameter, SQR, MAX_ABS_XYZ
OSLD float mx_clamp(float x, float a, float b)
{
    float m = x > a ? x : a;
    return m < b ? m : b;
}
OSLD float mx_lerp(float f, float a, float b)
{
    f = mx_clamp(f, 0.0f, 1.0f);
    return (1 - f) * a + f * b;
}
OSLD float mx_sqr(float x) { return x * x; }
OSLD float mx_max_abs(V3 v) { return fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z))); }

// ---- energy tables ---------------------------------------------------------------------------
// layout of the table block: [MiniMicrofacetGGX 1x16x16][ReflFront 32x16x16][BothFront][BothBack]
enum { LUT_GGX = 0, LUT_REFL_FRONT = 256, LUT_BOTH_FRONT = 256 + 8192, LUT_BOTH_BACK = 256 + 2 * 8192,
       LUT_WORDS = 256 + 3 * 8192,
       // then the Zeltner-Burley LTC coefficients (32x32x3) and spi::Thinlayer 32x16x16 (thinlayer closure)
       LUT_THINLAYER = LUT_WORDS + 32 * 32 * 3 };
struct EnergyCurve {   // TabulatedEnergyCurve<BSDF> (microfacet_tools_impl.h:110-207)
    const float* storedE;
    int Nf;            // Nr = Nc = 16 for every table used here
    bool sqr_cosines;  // MiniMicrofacet spaces its cosines quadratically, DielectricBSDF linearly
    float roughness, fresnel_index;
    float get_cosine(int i) const
    {
        const float x = float(i) * (1.0f / 15);
        return fmaxf(sqr_cosines ? x * x : x, 1e-6f);
    }
    float interpolate_emiss(int i) const
    {
        const int Nr = 16, Nc = 16;
        float rf = roughness * (Nr - 1);
        int ra   = (int)(rf);
        int rb   = ra + 1 < Nr - 1 ? ra + 1 : Nr - 1;
        rf -= ra;
        if (Nf == 1)
            return mx_lerp(rf, storedE[ra * Nc + i], storedE[rb * Nc + i]);
        float ff = fresnel_index * (Nf - 1);
        int fa   = (int)(ff);
        int fb   = fa + 1 < Nf - 1 ? fa + 1 : Nf - 1;
        ff -= fa;
        return mx_lerp(ff,
                         mx_lerp(rf, storedE[(fa * Nr + ra) * Nc + i], storedE[(fa * Nr + rb) * Nc + i]),
                         mx_lerp(rf, storedE[(fb * Nr + ra) * Nc + i], storedE[(fb * Nr + rb) * Nc + i]));
    }
    float Emiss_eval(float c) const
    {
        float cos0 = get_cosine(0);
        if (c <= cos0)
            return interpolate_emiss(0);
        for (int i = 1; i < 16; i++) {
            const float cos1 = get_cosine(i);
            if (c < cos1) {
                float q = (c - cos0) / (cos1 - cos0);
                return mx_lerp(q, interpolate_emiss(i - 1), interpolate_emiss(i));
            }
            cos0 = cos1;
        }
        return interpolate_emiss(15);
    }
};
OSLD float ggx_energy(const float* luts, float roughness, float cosNO)
{
    EnergyCurve c;
    c.storedE = luts + LUT_GGX; c.Nf = 1; c.sqr_cosines = true; c.roughness = roughness; c.fresnel_index = 0.0f;
    return c.Emiss_eval(cosNO);
}
OSLD float dielectric_energy(const float* luts, int table, float roughness, float fresnel_index, float cosNO)
{
    EnergyCurve c;
    c.storedE = luts + table; c.Nf = 32; c.sqr_cosines = false; c.roughness = roughness; c.fresnel_index = fresnel_index;
    return c.Emiss_eval(cosNO);
}

// ---- GGXDist (microfacet_tools_decl.h:13-46, _impl.h:19-107) ---------------------------------
struct GGXD {
    float ax = 0, ay = 0;
    GGXD() {}
    GGXD(float rough, float aniso, bool flip_aniso = false) : ax(mx_sqr(rough)), ay(mx_sqr(rough))
    {
        if (flip_aniso)
            aniso = -aniso;
        const float ALPHA_MIN = 1e-5f;
        ax = fmaxf(ax * (1 + aniso), ALPHA_MIN);
        ay = fmaxf(ay * (1 - aniso), ALPHA_MIN);
    }
    float roughness() const { return fmaxf(ax, ay); }
    float D(const V3& Hr) const
    {
        const float cosPhi2st2 = mx_sqr(Hr.x / ax);
        const float sinPhi2st2 = mx_sqr(Hr.y / ay);
        const float cosThetaM2 = mx_sqr(Hr.z);
        const float sinThetaM2 = cosPhi2st2 + sinPhi2st2;
        return 1.0f / ((float)OSLD_PI * ax * ay * mx_sqr(cosThetaM2 + sinThetaM2));
    }
    float G1(V3 w) const
    {
        w = mkv(w.x * ax, w.y * ay, w.z);
        return 2.0f * w.z / (w.z + imath_length(w));
    }
    float G2_G1(V3 wi, V3 wo) const
    {
        wi             = mkv(wi.x * ax, wi.y * ay, wi.z);
        wo             = mkv(wo.x * ax, wo.y * ay, wo.z);
        const float nl = imath_length(wi);
        const float nv = imath_length(wo);
        return wi.z * (wo.z + nv) / (wo.z * nl + wi.z * nv);
    }
    V3 sample_for_refl(const V3& wo, float randu, float randv) const
    {
        V3 i_std        = vnormalized(mkv(wo.x * ax, wo.y * ay, wo.z));
        const float phi = 2.0f * (float)OSLD_PI * randu;
        const float a   = mx_clamp(fminf(ax, ay), 0.0f, 1.0f);
        const float s   = 1 + sqrtf(mx_sqr(wo.x) + mx_sqr(wo.y));
        const float a2  = mx_sqr(a);
        const float s2  = mx_sqr(s);
        const float k   = (1 - a2) * s2 / (s2 + a2 * mx_sqr(wo.z));
        const float b   = k * i_std.z;
        const float z   = (1 - randv) * (1 + b) - b;
        const float sinTheta = sqrtf(mx_clamp(1 - mx_sqr(z), 0.0f, 1.0f));
        V3 o_std = mkv(sinTheta * fast_cos(phi), sinTheta * fast_sin(phi), z);
        V3 m_std = i_std + o_std;
        return vnormalized(mkv(m_std.x * ax, m_std.y * ay, m_std.z));
    }
    float D_refl_D(const V3& wo, const V3& m) const
    {
        (void)m;
        const float len2 = mx_sqr(wo.x * ax) + mx_sqr(wo.y * ay);
        const float t    = sqrtf(len2 + mx_sqr(wo.z));
        const float a    = mx_clamp(fminf(ax, ay), 0.0f, 1.0f);
        const float s    = 1 + sqrtf(mx_sqr(wo.x) + mx_sqr(wo.y));
        const float a2   = mx_sqr(a);
        const float s2   = mx_sqr(s);
        const float k    = (1 - a2) * s2 / (s2 + a2 * mx_sqr(wo.z));
        return 2 * wo.z / (k * wo.z + t);
    }
};

OSLD V3 bsdl_reflect(const V3& E, const V3& N) { return N * (2 * dot3(N, E)) - E; }
OSLD V3 bsdl_refract(const V3& E, const V3& N, float eta)
{
    V3 R = mkv(0.0f);
    if (eta == 0)
        return R;
    V3 Nn;
    float cosi = dot3(E, N), neta;
    if (cosi > 0) {
        neta = 1 / eta;
        Nn   = N;
    } else {
        cosi = -cosi;
        neta = eta;
        Nn   = -N;
    }
    float arg = 1 - (neta * neta * (1 - (cosi * cosi)));
    if (arg >= 0) {
        float dnp = sqrtf(arg);
        float nK  = (neta * cosi) - dnp;
        R         = vnormalized(E * (-neta) + Nn * nK);
    }
    return R;
}
OSLD float v3max(const V3& v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }
OSLD V3 v3clamped(const V3& v, float a, float b) { return mkv(mx_clamp(v.x, a, b), mx_clamp(v.y, a, b), mx_clamp(v.z, a, b)); }
OSLD V3 v3sqrt(const V3& v) { return mkv(sqrtf(v.x), sqrtf(v.y), sqrtf(v.z)); }
OSLD V3 v3div(const V3& a, const V3& b) { return mkv(a.x / b.x, a.y / b.y, a.z / b.z); }

#ifdef OSLD_THINLAYER
// what spi::ThinLayerLobe boils down to (functions in osl_b200_thinlayer.cuh)
struct ThinSpec {
    GGXD d;
    V3 sigma_t;
    float eta, thickness, roughness, prob_clamp;
    V3 refl_tint, refr_tint;
    float Eo;   // energy the microfacet lobes lose, handed to the diffuse / translucent pair
};
#endif

// ---- Fresnel terms ----------------------------------------------------------------------------
enum { MXF_CONDUCTOR, MXF_DIELECTRIC, MXF_SCHLICK };
struct MxFresnel {
    int kind = MXF_DIELECTRIC;
    V3 ior = mkv(0.0f), extinction = mkv(0.0f);   // conductor
    float eta = 1.5f;                           // dielectric / schlick: relative IOR as seen from wo's side
    V3 F0 = mkv(0.0f), F90 = mkv(1.0f);           // schlick
    float exponent = 5.0f, tir_cos = 0.0f;
    static float clamp_eta(float e, bool backside)   // DielectricFresnel::DielectricFresnel
    {
        const float IOR_MIN = 1.001f, IOR_MAX = 5.0f;
        if (backside)
            e = 1 / e;
        return e >= 1 ? mx_clamp(e, IOR_MIN, IOR_MAX) : mx_clamp(e, 1 / IOR_MAX, 1 / IOR_MIN);
    }
    float table_index() const
    {
        const float IOR_MIN = 1.001f, IOR_MAX = 5.0f;
        const float seta = mx_clamp(eta < 1 ? 1 / eta : eta, IOR_MIN, IOR_MAX);
        const float x    = (seta - IOR_MIN) * (1 / (IOR_MAX - IOR_MIN));
        return sqrtf(x);
    }
    V3 eval(float c) const
    {
        if (kind == MXF_CONDUCTOR) {   // ConductorFresnel::eval (bsdf_conductor_impl.h:22-45)
            const float FLOAT_MIN = 1.17549435e-38f, BIG = 1e12f;
            const float cos_theta = mx_clamp(c, 0.0f, 1.0f);
            const V3 one = mkv(1.0f);
            const V3 cosTheta2 = mkv(cos_theta * cos_theta);
            const V3 sinTheta2 = one - cosTheta2;
            const V3 n2        = ior * ior;
            const V3 k2        = extinction * extinction;
            const V3 t0        = n2 - k2 - sinTheta2;
            const V3 a2plusb2  = v3sqrt(t0 * t0 + 4 * n2 * k2);
            const V3 t1        = a2plusb2 + cosTheta2;
            const V3 a         = v3sqrt(0.5f * (a2plusb2 + t0));
            const V3 t2        = (2.0f * cos_theta) * a;
            const V3 rs        = v3div(t1 - t2, v3clamped(t1 + t2, FLOAT_MIN, BIG));
            const V3 t3        = cosTheta2 * a2plusb2 + sinTheta2 * sinTheta2;
            const V3 t4        = t2 * sinTheta2;
            const V3 rp        = v3div(rs * (t3 - t4), v3clamped(t3 + t4, FLOAT_MIN, BIG));
            return 0.5f * v3clamped(rp + rs, 0, 2);
        }
        if (kind == MXF_SCHLICK) {     // SchlickFresnel::eval (bsdf_schlick_impl.h:29-37)
            c = mx_clamp(c, 0.0f, 1.0f);
            if (c < tir_cos)
                return mkv(1.0f);
            const float f = mx_clamp(fast_safe_pow(1 - c, exponent), 0.0f, 1.0f);
            return (1 - f) * F0 + f * F90;
        }
        // DielectricFresnel::eval (bsdf_dielectric_impl.h:29-41)
        float g = (eta - 1.0f) * (eta + 1.0f) + c * c;
        if (g > 0) {
            g       = sqrtf(g);
            float A = (g - c) / (g + c);
            float B = (c * (g + c) - 1) / (c * (g - c) + 1);
            return mkv(0.5f * A * A * (1 + B * B));
        }
        return mkv(1.0f);
    }
    V3 conductor_avg() const   // ConductorFresnel::avg
    {
        const float a = -0.32775145f, b = 0.18346033f, c = 0.61146583f, d = -0.07785134f;
        float r[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float x = vcomp(ior, i), y = vcomp(extinction, i);
            const float p = a + b * x + c * y + d * x * y;
            r[i]          = mx_clamp(p / (1 + p), 0.0f, 1.0f);
        }
        return mkv(r[0], r[1], r[2]);
    }
};

// ---- the lobe state every MX microfacet closure boils down to ---------------------------------
struct MxSpec {
    GGXD d;
    MxFresnel f;
    float E_ms_spec = 0;   // DielectricBSDF::E_ms / ConductorLobe::E_ms: GGX single-scatter loss
    float E_ms      = 0;   // the lobe's own table look-up (DielectricLobe / SchlickLobe)
    bool dorefl = true, dorefr = false;
    bool conductor = false;
    V3 refl_tint = mkv(1.0f), refr_tint = mkv(0.0f);
    V3 wo_absorption = mkv(1.0f);
    float roughness  = 0;
};

// eval_turquin_microms_reflection (microfacet_tools_impl.h:333-362)
OSLD BSample mx_eval_turquin(const GGXD& dist, const MxFresnel& fresnel, float E_ms, const V3& wo, const V3& wi)
{
    const float cosNO = wo.z;
    const float cosNI = wi.z;
    if (cosNI <= 0 || cosNO <= 0)
        return bs_null();
    V3 m            = vnormalized(wo + wi);
    float cosMO     = dot3(m, wo);
    const float D   = dist.D(m);
    float D_refl_D  = dist.D_refl_D(wo, m);
    float D_refl    = D_refl_D * D;
    const float G1  = dist.G1(wo);
    float pdf       = D_refl / (4 * cosNO);
    const float out = dist.G2_G1(wi, wo) * G1 / D_refl_D;
    const V3 F      = fresnel.eval(cosMO);
    const V3 F_ms   = F;
    const float msf = E_ms / fmaxf(0.01f, 1 - E_ms);
    const V3 one = mkv(1.0f);
    const V3 O = out * F * (one + F_ms * msf);
    return bs_make(wi, O, pdf, -1);
}
OSLD BSample mx_sample_turquin(const GGXD& dist, const MxFresnel& fresnel, float E_ms, const V3& wo, const V3& rnd)
{
    if (wo.z <= 0)
        return bs_null();
    V3 m  = dist.sample_for_refl(wo, rnd.x, rnd.y);
    V3 wi = bsdl_reflect(wo, m);
    if (wi.z <= 0)
        return bs_null();
    return mx_eval_turquin(dist, fresnel, E_ms, wo, wi);
}
// DielectricBSDF<Fresnel>::eval / ::sample with use_bvn_refraction (bsdf_dielectric_impl.h:114-217)
OSLD BSample mx_spec_eval(const MxSpec& s, const V3& wo, const V3& wi)
{
    if (!s.dorefr)
        return mx_eval_turquin(s.d, s.f, s.E_ms_spec, wo, wi);
    const float cosNO = wo.z;
    const float cosNI = wi.z;
    if (cosNI > 0) {
        const V3 m        = vnormalized(wo + wi);
        const float cosMO = dot3(m, wo);
        if (cosMO <= 0)
            return bs_null();
        const float D  = s.d.D(m);
        const float G1 = s.d.G1(wo);
        const V3 F     = s.f.eval(cosMO);
        if (v3max(F) <= 0)
            return bs_null();
        const float D_refl_D = s.d.D_refl_D(wo, m);
        const float D_refl   = D_refl_D * D;
        const V3 out    = F * (s.d.G2_G1(wi, wo) * G1 / (D_refl_D * v3max(F)));
        const float pdf = D_refl / (4.0f * cosNO) * v3max(F);
        return bs_make(wi, out, pdf, 0);
    } else if (cosNI < 0) {
        const float eta = s.f.eta;
        const V3 Ht     = vnormalized(eta * wi + wo) * ((eta > 1) ? -1.0f : 1.0f);
        const float cosHO = dot3(Ht, wo);
        const float cosHI = dot3(Ht, wi);
        if (cosHO <= 0 || cosHI >= 0)
            return bs_null();
        const V3 Ft = mkv(1.0f) - s.f.eval(cosHO);
        if (Ht.z <= 0 || v3max(Ft) <= 0)
            return bs_null();
        const float D  = s.d.D(Ht);
        const float G1 = s.d.G1(wo);
        float J        = (-cosHI * cosHO * mx_sqr(eta)) / (wo.z * mx_sqr(cosHI * eta + cosHO));
        const float D_refl_D = s.d.D_refl_D(wo, Ht);
        const float D_refl   = D_refl_D * D;
        float pdf            = D_refl * J * v3max(Ft);
        const V3 out         = Ft * (s.d.G2_G1(mkv(wi.x, wi.y, -wi.z), wo) * G1 / (D_refl_D * v3max(Ft)));
        return bs_make(wi, out, pdf, 0);
    }
    return bs_null();
}
OSLD BSample mx_spec_sample(const MxSpec& s, const V3& wo, float randu, float randv, float randw)
{
    if (!s.dorefr)
        return mx_sample_turquin(s.d, s.f, s.E_ms_spec, wo, mkv(randu, randv, randw));
    V3 m              = s.d.sample_for_refl(wo, randu, randv);
    const float cosMO = dot3(wo, m);
    if (cosMO <= 0)
        return bs_null();
    const float F       = v3max(s.f.eval(cosMO));
    bool choose_reflect = randw < F;
    const V3 wi         = choose_reflect ? bsdl_reflect(wo, m) : bsdl_refract(wo, m, s.f.eta);
    if ((choose_reflect && wi.z <= 0) || (!choose_reflect && wi.z >= 0))
        return bs_null();
    return mx_spec_eval(s, wo, wi);
}

// the rx / ry -> (roughness, anisotropy) reparametrisation shared by the three constructors
OSLD void mx_roughness(float roughness_x, float roughness_y, float path_roughness, float& roughness, float& aniso,
                         bool& flip)
{
    const float EPSILON = 1e-4f;
    const float rx = mx_clamp(roughness_x, EPSILON, 2.0f);
    const float ry = mx_clamp(roughness_y, EPSILON, 2.0f);
    const float ax = fmaxf(rx, ry);
    const float ay = fminf(rx, ry);
    const float b  = ay / ax;
    aniso          = (1 - b) / (1 + b);
    roughness      = 1.0f - (1.0f - sqrtf(ax / (1 + aniso))) * (1.0f - path_roughness);   // regularize_roughness
    flip           = rx < ry;
}
// ConductorLobe::ConductorLobe (bsdf_conductor_impl.h:82-110)
OSLD MxSpec mx_conductor_setup(const float* luts, float cosNO, float rx, float ry, const V3& ior, const V3& extinction, float path_roughness)
{
    MxSpec s;
    float roughness, aniso;
    bool flip;
    mx_roughness(rx, ry, path_roughness, roughness, aniso, flip);
    s.d            = GGXD(roughness, aniso, flip);
    s.f.kind       = MXF_CONDUCTOR;
    s.f.ior        = ior;
    s.f.extinction = extinction;
    s.E_ms_spec    = ggx_energy(luts, roughness, cosNO);
    s.conductor    = true;
    s.dorefl       = true;
    s.dorefr       = false;
    s.roughness    = roughness;
    return s;
}
// DielectricLobe / SchlickLobe constructors (bsdf_dielectric_impl.h:219-293, bsdf_schlick_impl.h:39-102)
OSLD void mx_dielectric_finish(const float* luts, MxSpec& s, float cosNO, float roughness, float aniso, bool flip, bool backfacing)
{
    s.d         = GGXD(roughness, aniso, flip);
    s.E_ms      = 0;
    s.E_ms_spec = 0;
    if (!s.dorefr)
        s.E_ms_spec = ggx_energy(luts, roughness, cosNO);   // DielectricBSDF's constructor
    if (s.dorefl && !s.dorefr)
        s.E_ms = dielectric_energy(luts, LUT_REFL_FRONT, roughness, s.f.table_index(), cosNO);
    else if (s.dorefr)
        s.E_ms = dielectric_energy(luts, backfacing ? LUT_BOTH_BACK : LUT_BOTH_FRONT, roughness, s.f.table_index(), cosNO);
    s.roughness = roughness;
}
OSLD MxSpec mx_dielectric_setup(const float* luts, float cosNO, const V3& refl_tint, const V3& refr_tint, float rx, float ry, float ior,
                                  const V3& absorption, bool backfacing, float path_roughness)
{
    MxSpec s;
    s.refl_tint = refl_tint;
    s.refr_tint = refr_tint;
    s.dorefl    = v3max(refl_tint) > 0;
    s.dorefr    = v3max(refr_tint) > 0;
    float roughness, aniso;
    bool flip;
    mx_roughness(rx, ry, path_roughness, roughness, aniso, flip);
    const float IOR = mx_clamp(ior, 1.001f, 5.0f);
    s.f.kind        = MXF_DIELECTRIC;
    s.f.eta         = MxFresnel::clamp_eta(IOR / 1.0f, backfacing);   // relative_eta: outer_ior = 1
    mx_dielectric_finish(luts, s, cosNO, roughness, aniso, flip, backfacing);
    if (v3max(absorption) > 0 && s.dorefl && !s.dorefr) {
        const float FLOAT_MIN = 1.17549435e-38f;
        const float sinNO2  = 1 - mx_sqr(cosNO);
        const float inveta2 = mx_sqr(1 / s.f.eta);
        const float cos_p   = sqrtf(1 - fminf(1.0f, inveta2 * sinNO2));
        const float dist    = 1 / fmaxf(cos_p, FLOAT_MIN);
        s.wo_absorption     = mkv(fast_exp(-absorption.x * dist), fast_exp(-absorption.y * dist), fast_exp(-absorption.z * dist));
    }
    return s;
}
OSLD MxSpec mx_schlick_setup(const float* luts, float cosNO, const V3& refl_tint, const V3& refr_tint, float rx, float ry, const V3& F0in,
                               const V3& F90in, float exponent, bool backfacing, float path_roughness)
{
    MxSpec s;
    s.refl_tint = refl_tint;
    s.refr_tint = refr_tint;
    s.dorefl    = v3max(refl_tint) > 0;
    s.dorefr    = v3max(refr_tint) > 0;
    float roughness, aniso;
    bool flip;
    mx_roughness(rx, ry, path_roughness, roughness, aniso, flip);
    const float avg_F0         = mx_clamp((F0in.x + F0in.y + F0in.z) * (1.0f / 3), 0.0f, 0.99f);
    const float sqrt_F0        = sqrtf(avg_F0);
    const float refraction_ior = (1 + sqrt_F0) / (1 - sqrt_F0);
    s.f.kind     = MXF_SCHLICK;
    s.f.eta      = MxFresnel::clamp_eta(refraction_ior, backfacing);
    s.f.F0       = v3clamped(F0in, 0, 1);
    s.f.F90      = v3clamped(F90in, 0, 1);
    s.f.exponent = exponent;
    s.f.tir_cos  = s.f.eta >= 1 ? 0 : sqrtf(1 - mx_sqr(s.f.eta));
    mx_dielectric_finish(luts, s, cosNO, roughness, aniso, flip, backfacing);
    return s;
}
// ConductorLobe / DielectricLobe / SchlickLobe ::eval_impl, ::sample_impl (local frame)
OSLD BSample mx_eval_local(const MxSpec& s, const V3& wo, const V3& wi)
{
    if (s.conductor) {
        BSample r   = mx_eval_turquin(s.d, s.f, s.E_ms_spec, wo, wi);
        r.roughness = s.roughness;
        return r;
    }
    if (!s.dorefl && !s.dorefr)
        return bs_null();
    BSample r = mx_spec_eval(s, wo, wi);
    if (s.dorefr)
        r.weight = r.weight * (1 / fmaxf(0.01f, 1 - s.E_ms));
    r.weight    = r.weight * (r.wi.z > 0 ? s.refl_tint : s.refr_tint);
    r.roughness = s.roughness;
    return r;
}
OSLD BSample mx_sample_local(const MxSpec& s, const V3& wo, float rx, float ry, float rz)
{
    if (s.conductor) {
        BSample r   = mx_sample_turquin(s.d, s.f, s.E_ms_spec, wo, mkv(rx, ry, rz));
        r.roughness = s.roughness;
        return r;
    }
    if (!s.dorefl && !s.dorefr)
        return bs_null();
    BSample r = mx_spec_sample(s, wo, rx, ry, rz);
    if (s.dorefr)
        r.weight = r.weight * (1 / fmaxf(0.01f, 1 - s.E_ms));
    if (mx_max_abs(r.wi) < 1e-4f)
        return bs_null();
    r.weight    = r.weight * (r.wi.z > 0 ? s.refl_tint : s.refr_tint);
    r.roughness = s.roughness;
    return r;
}
OSLD V3 mx_albedo(const MxSpec& s)
{
    if (s.conductor)
        return s.f.conductor_avg();
    return !s.dorefr ? s.refl_tint * (1 - s.E_ms) : mkv(1.0f);
}
// filter_o: what a layer() lets through to its base (DielectricLobe / SchlickLobe)
OSLD V3 mx_filter_o(const MxSpec& s, bool schlick)
{
    if (s.dorefr)
        return mkv(0.0f);
    return schlick ? mkv(s.E_ms) : s.E_ms * s.wo_absorption;
}


}  // namespace osld
