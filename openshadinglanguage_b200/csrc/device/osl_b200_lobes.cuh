// osl_b200_lobes.cuh — testrender's glossy lobes for the wavefront integrator.
//
//   Phong                         src/testrender/shading.cpp:323-364
//   Ward                          src/testrender/shading.cpp:366-448
//   GGXDist / BeckmannDist        src/testrender/shading.cpp:472-560
//   Microfacet<Dist, 0|1|2>       src/testrender/shading.cpp:563-799
//
// Included from osl_b200_render.cuh after `struct Lobe` when the scene's
// materials use one of these closures (the host code generator defines
// OSLD_GLOSSY_LOBES); scenes without them keep the small 5-word lobe record.
// The reference picks the distribution and the reflect/refract/both variant
// through template parameters; here both are fields of the lobe record and the
// branches are warp-uniform for a material-sorted wavefront.
#pragma once

namespace osld {

OSLD float sqr_(float x) { return x * x; }
OSLD V3 frame_tolocal(const Lobe& l, V3 a) { return mkv(dot3(a, l.fu), dot3(a, l.fv), dot3(a, l.N)); }
OSLD V3 frame_toworld(const Lobe& l, V3 a) { return a.x * l.fu + a.y * l.fv + a.z * l.N; }
OSLD void lobe_set_frame(Lobe& l, V3 t)
{
    // TangentFrame::from_normal_and_tangent (sampling.h:30-42)
    V3 x        = t - l.N * dot3(l.N, t);
    float xlen2 = dot3(x, x);
    if (xlen2 > 0) {
        x    = x * (1.0f / sqrtf(xlen2));
        l.fu = x;
        l.fv = cross3(l.N, x);
    } else {
        TangentFrame f = frame_from_normal(l.N);
        l.fu           = f.u;
        l.fv           = f.v;
    }
}

// ---- distributions ---------------------------------------------------------------------
OSLD float dist_F(bool ggx, float tan_m2)
{
    if (ggx)
        return 1 / ((float)OSLD_PI * (1 + tan_m2) * (1 + tan_m2));
    return (float)(1 / OSLD_PI) * fast_exp(-tan_m2);
}
OSLD float dist_Lambda(bool ggx, float a2)
{
    if (ggx)
        return 0.5f * (-1.0f + sqrtf(1.0f + 1.0f / a2));
    const float a = sqrtf(a2);
    return a < 1.6f ? (1.0f - 1.259f * a + 0.396f * a2) / (3.535f * a + 2.181f * a2) : 0.0f;
}
OSLD void ggx_sample_slope(float cos_theta, float randu, float randv, float& sx, float& sy)
{
    float c   = cos_theta < 1e-6f ? 1e-6f : cos_theta;
    float Q   = (1 + c) * randu - c;
    float num = c * sqrtf((1 - c) * (1 + c)) - Q * sqrtf((1 - Q) * (1 + Q));
    float den = (Q - c) * (Q + c);
    float eps = 1.0f / 4294967296.0f;
    den       = fabsf(den) < eps ? copysignf(eps, den) : den;
    sx        = num / den;
    float Ru  = 1 - 2 * randv;
    float u2  = fabsf(Ru);
    float z   = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f))
              / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    sy = copysignf(1.0f, Ru) * z * sqrtf(1.0f + sx * sx);
}
OSLD void beckmann_sample_slope(float cos_theta, float randu, float randv, float& sx, float& sy)
{
    const float SQRT_PI_INV = 1 / sqrtf((float)OSLD_PI);
    float ct                = cos_theta < 1e-6f ? 1e-6f : cos_theta;
    float tanThetaI         = sqrtf(1 - ct * ct) / ct;
    float cotThetaI         = 1 / tanThetaI;
    float c       = fast_erf(cotThetaI);
    float K       = tanThetaI * SQRT_PI_INV;
    float yApprox = randu * (1.0f + c + K * (1 - c * c));
    float yExact  = randu * (1.0f + c + K * fast_exp(-cotThetaI * cotThetaI));
    float b = K > 0 ? (0.5f - sqrtf(K * (K - yApprox + 1.0f) + 0.25f)) / K : yApprox - 1.0f;
    float invErf = fast_ierf(b);
    float value  = 1.0f + b + K * fast_exp(-invErf * invErf) - yExact;
    if (fabsf(value) > 1e-6f) {
        b -= value / (1 - invErf * tanThetaI);
        invErf = fast_ierf(b);
        value  = 1.0f + b + K * fast_exp(-invErf * invErf) - yExact;
        b -= value / (1 - invErf * tanThetaI);
        sx = fast_ierf(b);
    } else {
        sx = invErf;
    }
    sy = fast_ierf(2.0f * randv - 1.0f);
}

// ---- Microfacet ------------------------------------------------------------------------
OSLD float mf_lambda(const Lobe& l, V3 w)
{
    float cosTheta2  = sqr_(w.z);
    float cosPhi2st2 = sqr_(w.x * l.ax);
    float sinPhi2st2 = sqr_(w.y * l.ay);
    return dist_Lambda(l.ggx, cosTheta2 / (cosPhi2st2 + sinPhi2st2));
}
OSLD float mf_D(const Lobe& l, V3 Hr)
{
    float cosThetaM = Hr.z;
    if (cosThetaM > 0) {
        float cosPhi2st2 = sqr_(Hr.x / l.ax);
        float sinPhi2st2 = sqr_(Hr.y / l.ay);
        float cosThetaM2 = sqr_(cosThetaM);
        float cosThetaM4 = sqr_(cosThetaM2);
        float tanThetaM2 = (cosPhi2st2 + sinPhi2st2) / cosThetaM2;
        return dist_F(l.ggx, tanThetaM2) / (l.ax * l.ay * cosThetaM4);
    }
    return 0;
}
OSLD V3 mf_sample_micronormal(const Lobe& l, V3 wo, float randu, float randv)
{
    V3 swo = wo;
    swo.x *= l.ax;
    swo.y *= l.ay;
    swo             = vnormalized(swo);
    float cos_theta = fmaxf(swo.z, 0.0f);
    float cos_phi = 1, sin_phi = 0;
    if (cos_theta < 0.99999f) {
        float invnorm = 1 / sqrtf(sqr_(swo.x) + sqr_(swo.y));
        cos_phi       = swo.x * invnorm;
        sin_phi       = swo.y * invnorm;
    }
    float slx, sly;
    if (l.ggx)
        ggx_sample_slope(cos_theta, randu, randv, slx, sly);
    else
        beckmann_sample_slope(cos_theta, randu, randv, slx, sly);
    float sx = cos_phi * slx - sin_phi * sly, sy = sin_phi * slx + cos_phi * sly;
    sx *= l.ax;
    sy *= l.ay;
    float mlen = sqrtf(sx * sx + sy * sy + 1);
    return mkv(fabsf(sx) < mlen ? -sx / mlen : 1.0f, fabsf(sy) < mlen ? -sy / mlen : 1.0f, 1.0f / mlen);
}
OSLD V3 mf_albedo(const Lobe& l, V3 wo)
{
    if (l.refract == 2)
        return mkv(1.0f);
    float fr = fresnel_dielectric(dot3(l.N, wo), l.eta);
    return mkv(l.refract ? 1 - fr : fr);
}
OSLD BSample mf_eval(const Lobe& l, V3 wo, V3 wi)
{
    const int Refract = l.refract;
    const float rough = fmaxf(l.ax, l.ay);
    const V3 wo_l = frame_tolocal(l, wo), wi_l = frame_tolocal(l, wi);
    if (Refract == 0 || Refract == 2) {
        if (wo_l.z > 0 && wi_l.z > 0) {
            const V3 m           = vnormalized(wi_l + wo_l);
            const float D        = mf_D(l, m);
            const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
            const float G2 = 1 / (Lambda_o + Lambda_i + 1), G1 = 1 / (Lambda_o + 1);
            const float Fr = fresnel_dielectric(dot3(m, wo_l), l.eta);
            float pdf      = (G1 * D * 0.25f) / wo_l.z;
            float out      = G2 / G1;
            if (Refract == 2) {
                pdf *= Fr;
                return bs_make(wi, mkv(out), pdf, rough);
            }
            return bs_make(wi, mkv(out * Fr), pdf, rough);
        }
    }
    if (Refract == 1 || Refract == 2) {
        if (wi_l.z < 0 && wo_l.z > 0.0f) {
            V3 ht = -(l.eta * wi_l + wo_l);
            if (l.eta < 1.0f)
                ht = -ht;
            V3 Ht             = vnormalized(ht);
            const float cosHO = dot3(Ht, wo_l);
            const float Ft    = 1.0f - fresnel_dielectric(cosHO, l.eta);
            if (Ft > 0) {
                const float cosHI = dot3(Ht, wi_l);
                if (Ht.z <= 0.0f)
                    return bs_null();
                const float Dt       = mf_D(l, Ht);
                const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
                const float G2 = 1 / (Lambda_o + Lambda_i + 1), G1 = 1 / (Lambda_o + 1);
                float invHt2 = 1 / dot3(ht, ht);
                float pdf    = (fabsf(cosHI * cosHO) * (l.eta * l.eta) * (G1 * Dt) * invHt2) / wo_l.z;
                float out    = G2 / G1;
                if (Refract == 2) {
                    pdf *= Ft;
                    return bs_make(wi, mkv(out), pdf, rough);
                }
                return bs_make(wi, mkv(out * Ft), pdf, rough);
            }
        }
    }
    return bs_null();
}
OSLD BSample mf_sample(const Lobe& l, V3 wo, float rx, float ry, float rz)
{
    const int Refract = l.refract;
    const float rough = fmaxf(l.ax, l.ay);
    const V3 wo_l     = frame_tolocal(l, wo);
    const float cosNO = wo_l.z;
    if (!(cosNO > 0))
        return bs_null();
    const V3 m        = mf_sample_micronormal(l, wo_l, rx, ry);
    const float cosMO = dot3(m, wo_l);
    const float F     = fresnel_dielectric(cosMO, l.eta);
    if (Refract == 0 || (Refract == 2 && rz < F)) {
        const V3 wi_l        = (2.0f * cosMO) * m - wo_l;
        const float D        = mf_D(l, m);
        const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
        const float G2 = 1 / (Lambda_o + Lambda_i + 1), G1 = 1 / (Lambda_o + 1);
        V3 wi     = frame_toworld(l, wi_l);
        float pdf = (G1 * D * 0.25f) / cosNO;
        float out = G2 / G1;
        if (Refract == 2) {
            pdf *= F;
            return bs_make(wi, mkv(out), pdf, rough);
        }
        return bs_make(wi, mkv(F * out), pdf, rough);
    }
    const V3 M = frame_toworld(l, m);
    V3 wi;
    float Ft             = fresnel_refraction(-wo, M, l.eta, wi);
    const V3 wi_l        = frame_tolocal(l, wi);
    const float cosHO    = dot3(m, wo_l), cosHI = dot3(m, wi_l);
    const float D        = mf_D(l, m);
    const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
    const float G2 = 1 / (Lambda_o + Lambda_i + 1), G1 = 1 / (Lambda_o + 1);
    const V3 ht        = -(l.eta * wi_l + wo_l);
    const float invHt2 = 1.0f / dot3(ht, ht);
    float pdf = (fabsf(cosHI * cosHO) * (l.eta * l.eta) * (G1 * D) * invHt2) / fabsf(wo_l.z);
    float out = G2 / G1;
    if (Refract == 2) {
        pdf *= Ft;
        return bs_make(wi, mkv(out), pdf, rough);
    }
    return bs_make(wi, mkv(Ft * out), pdf, rough);
}

// ---- libbsdl lobes behind testrender's BSDL_WRAP (shading.cpp:73-115) ----------------------
// BSDL_WRAP builds BsdfGlobals(wo, N, N, backfacing, path_roughness, 1, 0); the lobe frame is
// Frame(visible_normal(N)) (libbsdl/include/BSDL/bsdf_impl.h:48-76, tools.h:459-510) and
// eval / sample run in that local frame.  The frame's Z is kept in l.N.
OSLD float bsdl_max_abs_xyz(V3 v) { return fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z))); }
OSLD float bsdl_clamp(float x, float a, float b)
{
    float m = fmaxf(x, a);
    return m < b ? m : b;
}
OSLD V3 bsdl_visible_normal(V3 wo, V3 Ngf, V3 N)
{
    if (dot3(wo, N) > 0.0f)
        return N;
    V3 V = cross3(wo, N);
    if (bsdl_max_abs_xyz(V) < 1e-4f) {
        V = cross3(wo, Ngf);
        if (bsdl_max_abs_xyz(V) < 1e-4f) {
            const float s = copysignf(1.0f, wo.z);
            const float a = -1.0f / (s + wo.z);
            V             = mkv(wo.x * wo.y * a, s + wo.y * wo.y * a, -wo.y);
        } else
            V = vnormalized(V);
    } else
        V = vnormalized(V);
    return vnormalized(cross3(V, wo) + 1e-4f * wo);
}
OSLD void lobe_set_bsdl_frame(Lobe& l, V3 wo)
{
    TangentFrame f = frame_from_normal(bsdl_visible_normal(wo, l.N, l.N));
    l.fu           = f.u;
    l.fv           = f.v;
    l.N            = f.w;
}
#ifdef OSLD_MX_LOBES
// mtx::ConductorLobe / DielectricLobe / SchlickLobe from their closure components (parameter
// order = the libbsdl Data structs' registration order; the distribution string takes one word).
// Frame(Z = visible normal, X = U) (tools.h:483-495).
OSLD void mx_set_frame_zx(Lobe& l, V3 Z, V3 X)
{
    if (bsdl_max_abs_xyz(X) < 1e-4f || fabsf(dot3(Z, vnormalized(X))) > 0.999f) {
        TangentFrame f = frame_from_normal(Z);
        l.fu           = f.u;
        l.fv           = f.v;
    } else {
        l.fv = vnormalized(cross3(Z, X));
        l.fu = cross3(l.fv, Z);
    }
    l.N = Z;
}
OSLD void mx_from_component(const float* luts, Lobe& l, int id, PoolPtr p, V3 wo, bool backfacing, float path_roughness)
{
    l.type = LOBE_MX_SPEC;
    l.N    = mkv(p[0], p[1], p[2]);
    const V3 Z = bsdl_visible_normal(wo, l.N, l.N);
    mx_set_frame_zx(l, Z, mkv(p[3], p[4], p[5]));
    const float cosNO = dot3(Z, wo);
    if (id == MX_CONDUCTOR_ID)
        l.mx = mx_conductor_setup(luts, cosNO, p[6], p[7], mkv(p[8], p[9], p[10]), mkv(p[11], p[12], p[13]),
                                  path_roughness);
    else if (id == MX_DIELECTRIC_ID)
        l.mx = mx_dielectric_setup(luts, cosNO, mkv(p[6], p[7], p[8]), mkv(p[9], p[10], p[11]), p[12], p[13], p[14],
                                   mkv(p[18], p[19], p[20]), backfacing, path_roughness);
    else
        l.mx = mx_schlick_setup(luts, cosNO, mkv(p[6], p[7], p[8]), mkv(p[9], p[10], p[11]), p[12], p[13],
                                mkv(p[14], p[15], p[16]), mkv(p[17], p[18], p[19]), p[20], backfacing, path_roughness);
}
#endif
// tools.h:200-278
OSLD V3 bsdl_sample_cos_hemisphere(float randu, float randv)
{
    const float a = 2 * randu - 1, qa = fabsf(a);
    const float b = 2 * randv - 1, qb = fabsf(b);
    const float rad = qa > qb ? qa : qb;
    const float phi = qa > qb ? qb / qa : ((qa == qb) ? 1.0f : 2 - qa / qb);
    const float x2  = phi * phi;
    float cp = 0.01578646f + -0.00029826362f * x2;
    cp       = -0.30837047f + cp * x2;
    cp       = 0.99998736f + cp * x2;
    float sp = 0.0024843954015523195266723632812500f + -0.0000341485538228880614042282104492f * x2;
    sp       = -0.0807407423853874206542968750000000f + sp * x2;
    sp       = 0.7853975892066955566406250000000000f + sp * x2;
    sp       = sp * phi;
    return mkv(copysignf(rad * cp, a), copysignf(rad * sp, b), sqrtf(1 - rad * rad));
}
// mtx::OrenNayarDiffuseLobe::eval_impl without energy compensation (OREN_NAYAR_ID maps to
// {N, albedo 1, sigma, energy_compensation false}, shading.cpp:1496-1503;
// MTX/bsdf_oren_nayar_diffuse_impl.h:40-62).  sigma (clamped) is kept in l.ax.
OSLD BSample oren_nayar_eval_local(const Lobe& l, V3 wo, V3 wi)
{
    const float ONEOVERPI = 1 / (float)OSLD_PI;
    const float cosNI = bsdl_clamp(wi.z, 0.0f, 1.0f);
    const float cosNO = bsdl_clamp(wo.z, 0.0f, 1.0f);
    if (cosNI <= 0.0f || cosNO <= 0.0f)
        return bs_null();
    const float cosIO = bsdl_clamp(dot3(wo, wi), -1.0f, 1.0f);
    const float s     = cosIO - cosNI * cosNO;
    const float pdf   = cosNI * ONEOVERPI;
    if (!l.refract) {
        const float s2    = sqr_(l.ax);
        const float A     = 1.0f - 0.50f * s2 / (s2 + 0.33f);
        const float B     = 0.45f * s2 / (s2 + 0.09f);
        const float stinv = s > 0.0f ? s / fmaxf(cosNI, cosNO) : 0.0f;
        const float f_ss  = A + B * stinv;
        return bs_make(wi, l.albedo * f_ss, pdf, 1.0f);
    }
    // energy-preserving Oren-Nayar (MTX/bsdf_oren_nayar_diffuse_impl.h:24-37, 63-93)
    const float PI_F          = (float)OSLD_PI;
    const float constant1_FON = 0.5f - 2.0f / (3.0f * PI_F);
    const float constant2_FON = 2.0f / 3.0f - 28.0f / (15.0f * PI_F);
    const float sigma         = l.ax;
    const float AF            = 1.0f / (1.0f + constant1_FON * sigma);
    const float BF            = sigma * AF;
    float EF[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float mu = k == 0 ? cosNO : cosNI;
        const float Si = sqrtf(fmaxf(0.0f, 1.0f - mu * mu));
        const float G  = Si * (fast_acos(mu) - Si * mu)
                        + 2.0f * ((Si / fmaxf(mu, 1e-7f)) * (1.0f - Si * Si * Si) - Si) * (1.0f / 3.0f);
        EF[k] = AF + (BF * ONEOVERPI) * G;
    }
    const float stinv = s > 0.0f ? s / fmaxf(cosNI, cosNO) : s;
    const float f_ss  = AF * (1.0f + sigma * stinv);
    const float avgEF = AF * (1.0f + constant2_FON * sigma);
    const float om    = fmaxf(0.0f, 1.0f - avgEF);
    const V3 a        = l.albedo;
    const V3 rho_ms   = mkv(sqr_(a.x) * avgEF / (1 - a.x * om), sqr_(a.y) * avgEF / (1 - a.y * om),
                            sqr_(a.z) * avgEF / (1 - a.z * om));
    const float f_ms = fmaxf(1e-7f, 1.0f - EF[0]) * fmaxf(1e-7f, 1.0f - EF[1]) / fmaxf(1e-7f, 1.0f - avgEF);
    return bs_make(wi, l.albedo * f_ss + rho_ms * f_ms, pdf, 1.0f);
}
// mtx::BurleyDiffuseLobe (MTX/bsdf_burley_diffuse_impl.h:24-58)
OSLD float burley_fresnel(float cos_theta, float F90)
{
    const float x  = bsdl_clamp(1.0f - cos_theta, 0.0f, 1.0f);
    const float x2 = x * x;
    float f        = bsdl_clamp(x2 * x2 * x, 0.0f, 1.0f);
    return (1 - f) * 1.0f + f * F90;
}
OSLD BSample burley_eval_local(const Lobe& l, V3 wo, V3 wi)
{
    const float ONEOVERPI = 1 / (float)OSLD_PI;
    if (wo.z <= 0.0f || wi.z <= 0.0f)
        return bs_null();
    const V3 H = wi + wo;
    if (bsdl_max_abs_xyz(H) < 1e-4f)
        return bs_null();
    const V3 Hn       = vnormalized(H);
    const float cosHI = bsdl_clamp(dot3(wi, Hn), 0.0f, 1.0f);
    const float cosNO = bsdl_clamp(wo.z, 0.0f, 1.0f);
    const float cosNI = bsdl_clamp(wi.z, 0.0f, 1.0f);
    const float F90   = 0.5f + 2.0f * l.ax * sqr_(cosHI);
    const float refL  = burley_fresnel(cosNI, F90);
    const float refV  = burley_fresnel(cosNO, F90);
    return bs_make(wi, l.albedo * (refL * refV), cosNI * ONEOVERPI, 1.0f);
}
OSLD BSample bsdl_diffuse_eval_local(const Lobe& l, V3 wo, V3 wi)
{
    return l.type == LOBE_BSDL_BURLEY ? burley_eval_local(l, wo, wi) : oren_nayar_eval_local(l, wo, wi);
}
// BSDL_WRAP::eval / ::sample (shading.cpp:88-103)
OSLD BSample bsdl_diffuse_eval(const Lobe& l, V3 wo, V3 wi)
{
    BSample s = bsdl_diffuse_eval_local(l, frame_tolocal(l, wo), frame_tolocal(l, wi));
    s.wi      = wi;
    return s;
}
OSLD BSample bsdl_diffuse_sample(const Lobe& l, V3 wo, float rx, float ry)
{
    const V3 wo_l = frame_tolocal(l, wo);
    BSample s     = bs_null();
    if (!(wo_l.z <= 0.0f))
        s = bsdl_diffuse_eval_local(l, wo_l, bsdl_sample_cos_hemisphere(rx, ry));
    s.wi = frame_toworld(l, s.wi);
    return s;
}

// ---- mtx::SheenLobe, Conty-Kulla mode (MTX/bsdf_sheen_impl.h:17-175) ------------------------
// Lobe fields: frame (fu, fv, N) = Frame(visible normal, wo); ax = sheen alpha, ay = regularized
// roughness, eta = Emiss (what the lobe lets through to a layer below), refract = backfacing,
// albedo = tint.
OSLD V3 bsdl_sample_uniform_hemisphere(float randu, float randv)
{
    const float a = 2 * randu - 1, qa = fabsf(a);
    const float b = 2 * randv - 1, qb = fabsf(b);
    const float rad = qa > qb ? qa : qb;
    const float phi = qa > qb ? qb / qa : ((qa == qb) ? 1.0f : 2 - qa / qb);
    const float x2  = phi * phi;
    float cp = 0.01578646f + -0.00029826362f * x2;
    cp       = -0.30837047f + cp * x2;
    cp       = 0.99998736f + cp * x2;
    float sp = 0.0024843954015523195266723632812500f + -0.0000341485538228880614042282104492f * x2;
    sp       = -0.0807407423853874206542968750000000f + sp * x2;
    sp       = 0.7853975892066955566406250000000000f + sp * x2;
    sp       = sp * phi;
    const float x = copysignf(rad * cp, a), y = copysignf(rad * sp, b);
    const float cos_theta = 1 - rad * rad;
    const float sin_theta = sqrtf(2 - rad * rad);
    return mkv(sin_theta * x, sin_theta * y, cos_theta);
}
OSLD float sheen_conty_albedo(float cosNO, float rough)
{
    float rx = 13.67300f, ry = 1.0f;
    rx = rx + -68.78018f * cosNO;               ry = ry + 61.57746f * cosNO;
    rx = rx + 799.08825f * rough;               ry = ry + 442.78211f * rough;
    rx = rx + -905.00061f * cosNO * rough;      ry = ry + 2597.49308f * cosNO * rough;
    rx = rx + 60.28956f * cosNO * cosNO;        ry = ry + 121.81241f * cosNO * cosNO;
    rx = rx + 1086.96473f * rough * rough;      ry = ry + 3045.55075f * rough * rough;
    return bsdl_clamp(rx / ry, 0.0f, 1.0f);
}
// ---- mtx::ZeltnerBurleySheen (MTX/bsdf_sheen_impl.h:205-355): sheen as a linearly transformed
// cosine.  The (A, B, R) coefficients are bilinear look-ups in a 32 x 32 table over (roughness,
// cos theta_o) that follows the energy tables in the LUT block (data/zeltner_ltc.bin); they depend
// on the view direction only, so the lobe fetches them once at set-up (l.ltc).  Compiled into
// modules whose materials pass a "mode" keyword to sheen_bsdf (OSLD_SHEEN_LTC).
#ifdef OSLD_SHEEN_LTC
#define OSLD_LUT_ZELTNER (256 + 3 * 8192)
OSLD V3 zeltner_fetch_coeffs(const float* luts, float roughness, float cosNO)
{
    const float ALMOSTONE = 0.999999940395355224609375f;
    const float row = bsdl_clamp(roughness, 0.0f, ALMOSTONE) * 31;
    const float col = bsdl_clamp(cosNO, 0.0f, ALMOSTONE) * 31;
    const float r = floorf(row), c = floorf(col);
    const float rf = row - r, cf = col - c;
    const float* T = luts + OSLD_LUT_ZELTNER + 3 * ((int)r * 32 + (int)c);
    const V3 v1 = mkv(__ldg(T), __ldg(T + 1), __ldg(T + 2)), v2 = mkv(__ldg(T + 3), __ldg(T + 4), __ldg(T + 5));
    const V3 v3 = mkv(__ldg(T + 96), __ldg(T + 97), __ldg(T + 98)), v4 = mkv(__ldg(T + 99), __ldg(T + 100), __ldg(T + 101));
    const V3 a = (1 - cf) * v1 + cf * v2, b = (1 - cf) * v3 + cf * v4;
    return (1 - rf) * a + rf * b;
}
OSLD BSample zeltner_eval_ltc(V3 wi, V3 ltc)
{
    const float a_inv = ltc.x, b_inv = ltc.y, r_coeff = ltc.z;
    const V3 wi_orig  = mkv(a_inv * wi.x + b_inv * wi.z, a_inv * wi.y, wi.z);
    const float q        = a_inv / dot3(wi_orig, wi_orig);
    const float jacobian = q * q;
    const float pdf      = jacobian * fmaxf(wi_orig.z, 0.0f) * (1 / (float)OSLD_PI);
    if (pdf > 1.17549435e-38f)
        return bs_make(wi, mkv(r_coeff), pdf, 0.0f);
    return bs_null();
}
#endif
// everything SheenLobe's constructor derives; l.N (shading normal) and l.albedo set by the caller.
// l.refract: bit 0 = backfacing, bit 1 = Zeltner mode ("mode" keyword == 1)
OSLD void sheen_setup(Lobe& l, V3 wo, float roughness_param, bool backfacing, float path_roughness, int mode = 0,
                      const float* luts = nullptr)
{
    const V3 Z = bsdl_visible_normal(wo, l.N, l.N);
    // Frame(Z, X = wo) (tools.h:483-495)
    if (bsdl_max_abs_xyz(wo) < 1e-4f || fabsf(dot3(Z, vnormalized(wo))) > 0.999f) {
        TangentFrame f = frame_from_normal(Z);
        l.fu           = f.u;
        l.fv           = f.v;
    } else {
        l.fv = vnormalized(cross3(Z, wo));
        l.fu = cross3(l.fv, Z);
    }
    l.N           = Z;
    const float r = bsdl_clamp(roughness_param, 0.0f, 1.0f);
    l.ay          = 1.0f - (1.0f - r) * (1.0f - path_roughness);
    l.ax          = fmaxf(0.06f, l.ay);
    l.refract     = backfacing ? 1 : 0;
    const float cosNO = bsdl_clamp(dot3(Z, wo), 0.0f, 1.0f);
    const float tmax  = fmaxf(l.albedo.x, fmaxf(l.albedo.y, l.albedo.z));
#ifdef OSLD_SHEEN_LTC
    if (mode == 1) {
        l.refract |= 2;
        l.ax = bsdl_clamp(fmaxf(0.02f, sqrtf(l.ay)), 0.02f, 1.0f);   // sheen_alpha, then ZeltnerBurleySheen's clamp
        // eval / sample look the coefficients up at wo.z in the lobe's frame
        l.ltc = zeltner_fetch_coeffs(luts, l.ax, dot3(Z, wo));
        l.eta = backfacing ? 1.0f : 1 - fminf(zeltner_fetch_coeffs(luts, l.ax, cosNO).z * tmax, 1.0f);
        return;
    }
#endif
    l.eta = backfacing ? 1.0f : 1 - fminf(sheen_conty_albedo(cosNO, bsdl_clamp(l.ax, 0.06f, 1.0f)) * tmax, 1.0f);
}
// SheenMicrofacet<ContyKullaDist<false>>::eval
OSLD BSample sheen_conty_eval(const Lobe& l, V3 wo, V3 wi)
{
    const float PI_F = (float)OSLD_PI, ONEOVERPI = 1 / (float)OSLD_PI;
    const float cosNO = wo.z, cosNI = wi.z;
    BSample s = bs_null();
    if (!(cosNI <= 1e-5f || cosNO <= 1e-5f)) {
        const float a   = bsdl_clamp(l.ax, 0.06f, 1.0f);
        const V3 Hr     = vnormalized(wo + wi);
        float cos_theta = bsdl_clamp(Hr.z, 0.0f, 1.0f);
        float sin_theta = sqrtf(1.0f - sqr_(cos_theta));
        const float D   = fast_safe_pow(sin_theta, 1 / a) * (2 + 1 / a) * 0.5f * ONEOVERPI;
        if (!((double)D < 1e-6)) {
            float cI = fminf(1.0f, wi.z), cO = fminf(1.0f, wo.z);
            const float G2 = (cI * cO) / (cI + cO - cI * cO);
            s = bs_make(wi, mkv(D * G2 * 0.5f * PI_F / cosNO), 0.5f * ONEOVERPI, 0.0f);
        }
    }
    return s;
}
// SheenLobe::eval_impl / sample_impl: the mode's lobe, then the tint and the roughness tag
OSLD BSample sheen_eval(const Lobe& l, V3 wo, V3 wi)
{
    const V3 wo_l = frame_tolocal(l, wo), wi_l = frame_tolocal(l, wi);
    BSample s     = bs_null();
    if (wi_l.z > 0 && wo_l.z >= 0 && !(l.refract & 1)) {
#ifdef OSLD_SHEEN_LTC
        if (l.refract & 2)
            s = zeltner_eval_ltc(wi_l, l.ltc);
        else
#endif
            s = sheen_conty_eval(l, wo_l, wi_l);
        s.weight    = s.weight * l.albedo;
        s.roughness = l.ay;
    }
    s.wi = wi;
    return s;
}
OSLD BSample sheen_sample(const Lobe& l, V3 wo, float rx, float ry)
{
    BSample s = bs_null();
    if (!(l.refract & 1)) {
        const V3 wo_l = frame_tolocal(l, wo);
#ifdef OSLD_SHEEN_LTC
        if (l.refract & 2) {
            if (!(wo_l.z < 0)) {   // cosine base distribution transformed by M
                const V3 o = bsdl_sample_cos_hemisphere(rx, ry);
                s = zeltner_eval_ltc(vnormalized(mkv(o.x - o.z * l.ltc.y, o.y, l.ltc.x * o.z)), l.ltc);
            }
        } else
#endif
            s = sheen_conty_eval(l, wo_l, bsdl_sample_uniform_hemisphere(rx, ry));
        s.weight    = s.weight * l.albedo;
        s.roughness = l.ay;
    }
    s.wi = frame_toworld(l, s.wi);
    return s;
}

// ---- Phong (exponent kept in l.ax) -------------------------------------------------------
OSLD BSample phong_eval(const Lobe& l, V3 wo, V3 wi)
{
    const float exponent = l.ax;
    float cosNI = dot3(l.N, wi), cosNO = dot3(l.N, wo);
    if (cosNI > 0 && cosNO > 0) {
        V3 R        = (2 * cosNO) * l.N - wo;
        float cosRI = dot3(R, wi);
        if (cosRI > 0) {
            const float pdf = (exponent + 1) * (float)(0.31830988618379067154 / 2) * fast_safe_pow(cosRI, exponent);
            return bs_make(wi, mkv(cosNI * (exponent + 2) / (exponent + 1)), pdf, 1 / (1 + exponent));
        }
    }
    return bs_null();
}
OSLD BSample phong_sample(const Lobe& l, V3 wo, float rx, float ry)
{
    const float exponent = l.ax;
    float cosNO          = dot3(l.N, wo);
    if (cosNO > 0) {
        V3 R      = (2 * cosNO) * l.N - wo;
        float phi = 2 * (float)OSLD_PI * rx;
        float sp, cp;
        fast_sincos(phi, &sp, &cp);
        float cosTheta  = fast_safe_pow(ry, 1 / (exponent + 1));
        float sinTheta2 = 1 - cosTheta * cosTheta;
        float sinTheta  = sinTheta2 > 0 ? sqrtf(sinTheta2) : 0;
        V3 wi           = frame_get(frame_from_normal(R), cp * sinTheta, sp * sinTheta, cosTheta);
        return phong_eval(l, wo, wi);
    }
    return bs_null();
}

// ---- Ward ------------------------------------------------------------------------------
OSLD BSample ward_eval(const Lobe& l, V3 wo, V3 wi)
{
    float cosNO = dot3(l.N, wo), cosNI = dot3(l.N, wi);
    if (cosNI > 0 && cosNO > 0) {
        V3 H       = vnormalized(wi + wo);
        float dotx = dot3(H, l.fu) / l.ax, doty = dot3(H, l.fv) / l.ay, dotn = dot3(H, l.N);
        float oh   = dot3(H, wi);
        float e    = fast_exp(-(dotx * dotx + doty * doty) / (dotn * dotn));
        float c    = (float)(4 * OSLD_PI) * l.ax * l.ay;
        float k    = oh * dotn * dotn * dotn;
        float pdf  = e / (c * k);
        return bs_make(wi, mkv(k * sqrtf(cosNI / cosNO)), pdf, fmaxf(l.ax, l.ay));
    }
    return bs_null();
}
OSLD BSample ward_sample(const Lobe& l, V3 wo, float rx, float ry)
{
    float cosNO = dot3(l.N, wo);
    if (cosNO > 0) {
        float phi = 2 * (float)OSLD_PI * rx;
        float sp, cp;
        fast_sincos(phi, &sp, &cp);
        float cosPhi = l.ax * cp, sinPhi = l.ay * sp;
        float k      = 1 / sqrtf(cosPhi * cosPhi + sinPhi * sinPhi);
        cosPhi *= k;
        sinPhi *= k;
        float thetaDenom = (cosPhi * cosPhi) / (l.ax * l.ax) + (sinPhi * sinPhi) / (l.ay * l.ay);
        float tanTheta2  = -fast_log(1 - ry) / thetaDenom;
        float cosTheta   = 1 / sqrtf(1 + tanTheta2);
        float sinTheta   = cosTheta * sqrtf(tanTheta2);
        V3 h             = mkv(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        float dotx = h.x / l.ax, doty = h.y / l.ay, dotn = h.z;
        h           = frame_toworld(l, h);
        float oh    = dot3(h, wo);
        V3 wi       = 2 * oh * h - wo;
        float cosNI = dot3(l.N, wi);
        if (cosNI > 0) {
            float e   = fast_exp(-(dotx * dotx + doty * doty) / (dotn * dotn));
            float c   = (float)(4 * OSLD_PI) * l.ax * l.ay;
            float kk  = oh * dotn * dotn * dotn;
            float pdf = e / (c * kk);
            return bs_make(wi, mkv(kk * sqrtf(cosNI / cosNO)), pdf, fmaxf(l.ax, l.ay));
        }
    }
    return bs_null();
}

}  // namespace osld
