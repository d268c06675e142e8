// osl_oracle_lobes.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of testrender's native glossy lobes:
//   Phong                         src/testrender/shading.cpp:323-364
//   Ward                          src/testrender/shading.cpp:366-448
//   GGXDist / BeckmannDist        src/testrender/shading.cpp:472-560
//   Microfacet<Dist, 0|1|2>       src/testrender/shading.cpp:563-799
// Included from osl_oracle_render.h right after `struct Lobe`.
#pragma once

namespace oslo {

namespace lobes {

inline float SQR(float x) { return x * x; }

struct V2f {
    float x, y;
};

// ---- distributions ---------------------------------------------------------------
inline float dist_F(bool ggx, float tan_m2)
{
    if (ggx)
        return 1 / (float(M_PI) * (1 + tan_m2) * (1 + tan_m2));
    return float(1 / M_PI) * fast_exp(-tan_m2);
}
inline float dist_Lambda(bool ggx, float a2)
{
    if (ggx)
        return 0.5f * (-1.0f + std::sqrt(1.0f + 1.0f / a2));
    const float a = std::sqrt(a2);
    return a < 1.6f ? (1.0f - 1.259f * a + 0.396f * a2) / (3.535f * a + 2.181f * a2) : 0.0f;
}
inline V2f ggx_sample_slope(float cos_theta, float randu, float randv)
{
    V2f slope;
    float c   = cos_theta < 1e-6f ? 1e-6f : cos_theta;
    float Q   = (1 + c) * randu - c;
    float num = c * std::sqrt((1 - c) * (1 + c)) - Q * std::sqrt((1 - Q) * (1 + Q));
    float den = (Q - c) * (Q + c);
    float eps = 1.0f / 4294967296.0f;
    den       = std::fabs(den) < eps ? std::copysign(eps, den) : den;
    slope.x   = num / den;
    float Ru  = 1 - 2 * randv;
    float u2  = std::fabs(Ru);
    float z   = (u2 * (u2 * (u2 * 0.27385f - 0.73369f) + 0.46341f))
              / (u2 * (u2 * (u2 * 0.093073f + 0.309420f) - 1.0f) + 0.597999f);
    slope.y = std::copysign(1.0f, Ru) * z * std::sqrt(1.0f + slope.x * slope.x);
    return slope;
}
inline V2f beckmann_sample_slope(float cos_theta, float randu, float randv)
{
    const float SQRT_PI_INV = 1 / std::sqrt(float(M_PI));
    float ct                = cos_theta < 1e-6f ? 1e-6f : cos_theta;
    float tanThetaI         = std::sqrt(1 - ct * ct) / ct;
    float cotThetaI         = 1 / tanThetaI;
    float c       = fast_erf(cotThetaI);
    float K       = tanThetaI * SQRT_PI_INV;
    float yApprox = randu * (1.0f + c + K * (1 - c * c));
    float yExact  = randu * (1.0f + c + K * fast_exp(-cotThetaI * cotThetaI));
    float b = K > 0 ? (0.5f - std::sqrt(K * (K - yApprox + 1.0f) + 0.25f)) / K : yApprox - 1.0f;
    float invErf = fast_ierf(b);
    float value  = 1.0f + b + K * fast_exp(-invErf * invErf) - yExact;
    V2f slope;
    if (std::fabs(value) > 1e-6f) {
        b -= value / (1 - invErf * tanThetaI);
        invErf = fast_ierf(b);
        value  = 1.0f + b + K * fast_exp(-invErf * invErf) - yExact;
        b -= value / (1 - invErf * tanThetaI);
        slope.x = fast_ierf(b);
    } else {
        slope.x = invErf;
    }
    slope.y = fast_ierf(2.0f * randv - 1.0f);
    return slope;
}

// ---- Microfacet -------------------------------------------------------------------
inline float mf_lambda(const Lobe& l, const V3& w)
{
    float cosTheta2  = SQR(w.z);
    float cosPhi2st2 = SQR(w.x * l.ax);
    float sinPhi2st2 = SQR(w.y * l.ay);
    return dist_Lambda(l.ggx, cosTheta2 / (cosPhi2st2 + sinPhi2st2));
}
inline float mf_G2(float Li, float Lo) { return 1 / (Li + Lo + 1); }
inline float mf_G1(float Lv) { return 1 / (Lv + 1); }
inline float mf_D(const Lobe& l, const V3& Hr)
{
    float cosThetaM = Hr.z;
    if (cosThetaM > 0) {
        float cosPhi2st2 = SQR(Hr.x / l.ax);
        float sinPhi2st2 = SQR(Hr.y / l.ay);
        float cosThetaM2 = SQR(cosThetaM);
        float cosThetaM4 = SQR(cosThetaM2);
        float tanThetaM2 = (cosPhi2st2 + sinPhi2st2) / cosThetaM2;
        return dist_F(l.ggx, tanThetaM2) / (l.ax * l.ay * cosThetaM4);
    }
    return 0;
}
inline V3 mf_sample_micronormal(const Lobe& l, const V3& wo, float randu, float randv)
{
    V3 swo = wo;
    swo.x *= l.ax;
    swo.y *= l.ay;
    swo             = normalized(swo);
    float cos_theta = std::max(swo.z, 0.0f);
    float cos_phi = 1, sin_phi = 0;
    if (cos_theta < 0.99999f) {
        float invnorm = 1 / std::sqrt(SQR(swo.x) + SQR(swo.y));
        cos_phi       = swo.x * invnorm;
        sin_phi       = swo.y * invnorm;
    }
    V2f slope = l.ggx ? ggx_sample_slope(cos_theta, randu, randv) : beckmann_sample_slope(cos_theta, randu, randv);
    V2f s { cos_phi * slope.x - sin_phi * slope.y, sin_phi * slope.x + cos_phi * slope.y };
    s.x *= l.ax;
    s.y *= l.ay;
    float mlen = std::sqrt(s.x * s.x + s.y * s.y + 1);
    return V3(std::fabs(s.x) < mlen ? -s.x / mlen : 1.0f, std::fabs(s.y) < mlen ? -s.y / mlen : 1.0f, 1.0f / mlen);
}
inline V3 mf_albedo(const Lobe& l, const V3& wo)
{
    if (l.refract == 2)
        return V3(1.0f);
    float fr = fresnel_dielectric(dot(l.N, wo), l.eta);
    return V3(l.refract ? 1 - fr : fr);
}
inline BSample mf_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    const int Refract = l.refract;
    const float rough = std::max(l.ax, l.ay);
    const V3 wo_l = l.tf.tolocal(wo), wi_l = l.tf.tolocal(wi);
    if (Refract == 0 || Refract == 2) {
        if (wo_l.z > 0 && wi_l.z > 0) {
            const V3 m           = normalized(wi_l + wo_l);
            const float D        = mf_D(l, m);
            const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
            const float G2 = mf_G2(Lambda_o, Lambda_i), G1 = mf_G1(Lambda_o);
            const float Fr = fresnel_dielectric(dot(m, wo_l), l.eta);
            float pdf      = (G1 * D * 0.25f) / wo_l.z;
            float out      = G2 / G1;
            if (Refract == 2) {
                pdf *= Fr;
                return BSample(wi, V3(out), pdf, rough);
            }
            return BSample(wi, V3(out * Fr), pdf, rough);
        }
    }
    if (Refract == 1 || Refract == 2) {
        if (wi_l.z < 0 && wo_l.z > 0.0f) {
            V3 ht = -(l.eta * wi_l + wo_l);
            if (l.eta < 1.0f)
                ht = -ht;
            V3 Ht             = normalized(ht);
            const float cosHO = dot(Ht, wo_l);
            const float Ft    = 1.0f - fresnel_dielectric(cosHO, l.eta);
            if (Ft > 0) {
                const float cosHI = dot(Ht, wi_l);
                if (Ht.z <= 0.0f)
                    return BSample();
                const float Dt       = mf_D(l, Ht);
                const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
                const float G2 = mf_G2(Lambda_o, Lambda_i), G1 = mf_G1(Lambda_o);
                float invHt2 = 1 / dot(ht, ht);
                float pdf    = (std::fabs(cosHI * cosHO) * (l.eta * l.eta) * (G1 * Dt) * invHt2) / wo_l.z;
                float out    = G2 / G1;
                if (Refract == 2) {
                    pdf *= Ft;
                    return BSample(wi, V3(out), pdf, rough);
                }
                return BSample(wi, V3(out * Ft), pdf, rough);
            }
        }
    }
    return BSample();
}
inline BSample mf_sample(const Lobe& l, const V3& wo, float rx, float ry, float rz)
{
    const int Refract = l.refract;
    const float rough = std::max(l.ax, l.ay);
    const V3 wo_l     = l.tf.tolocal(wo);
    const float cosNO = wo_l.z;
    if (!(cosNO > 0))
        return BSample();
    const V3 m        = mf_sample_micronormal(l, wo_l, rx, ry);
    const float cosMO = dot(m, wo_l);
    const float F     = fresnel_dielectric(cosMO, l.eta);
    if (Refract == 0 || (Refract == 2 && rz < F)) {
        const V3 wi_l        = (2.0f * cosMO) * m - wo_l;
        const float D        = mf_D(l, m);
        const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
        const float G2 = mf_G2(Lambda_o, Lambda_i), G1 = mf_G1(Lambda_o);
        V3 wi     = l.tf.toworld(wi_l);
        float pdf = (G1 * D * 0.25f) / cosNO;
        float out = G2 / G1;
        if (Refract == 2) {
            pdf *= F;
            return BSample(wi, V3(out), pdf, rough);
        }
        return BSample(wi, V3(F * out), pdf, rough);
    }
    const V3 M = l.tf.toworld(m);
    V3 wi;
    float Ft             = fresnel_refraction(-wo, M, l.eta, wi);
    const V3 wi_l        = l.tf.tolocal(wi);
    const float cosHO    = dot(m, wo_l), cosHI = dot(m, wi_l);
    const float D        = mf_D(l, m);
    const float Lambda_o = mf_lambda(l, wo_l), Lambda_i = mf_lambda(l, wi_l);
    const float G2 = mf_G2(Lambda_o, Lambda_i), G1 = mf_G1(Lambda_o);
    const V3 ht        = -(l.eta * wi_l + wo_l);
    const float invHt2 = 1.0f / dot(ht, ht);
    float pdf = (std::fabs(cosHI * cosHO) * (l.eta * l.eta) * (G1 * D) * invHt2) / std::fabs(wo_l.z);
    float out = G2 / G1;
    if (Refract == 2) {
        pdf *= Ft;
        return BSample(wi, V3(out), pdf, rough);
    }
    return BSample(wi, V3(Ft * out), pdf, rough);
}

// ---- Phong ------------------------------------------------------------------------
inline BSample phong_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    float cosNI = dot(l.N, wi), cosNO = dot(l.N, wo);
    if (cosNI > 0 && cosNO > 0) {
        V3 R        = (2 * cosNO) * l.N - wo;
        float cosRI = dot(R, wi);
        if (cosRI > 0) {
            const float pdf = (l.exponent + 1) * float(M_1_PI / 2) * fast_safe_pow(cosRI, l.exponent);
            return BSample(wi, V3(cosNI * (l.exponent + 2) / (l.exponent + 1)), pdf, 1 / (1 + l.exponent));
        }
    }
    return BSample();
}
inline BSample phong_sample(const Lobe& l, const V3& wo, float rx, float ry)
{
    float cosNO = dot(l.N, wo);
    if (cosNO > 0) {
        V3 R      = (2 * cosNO) * l.N - wo;
        float phi = 2 * float(M_PI) * rx;
        float sp, cp;
        fast_sincos(phi, &sp, &cp);
        float cosTheta  = fast_safe_pow(ry, 1 / (l.exponent + 1));
        float sinTheta2 = 1 - cosTheta * cosTheta;
        float sinTheta  = sinTheta2 > 0 ? std::sqrt(sinTheta2) : 0;
        V3 wi = TangentFrame::from_normal(R).get(cp * sinTheta, sp * sinTheta, cosTheta);
        return phong_eval(l, wo, wi);
    }
    return BSample();
}

// ---- Ward -------------------------------------------------------------------------
inline BSample ward_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    float cosNO = dot(l.N, wo), cosNI = dot(l.N, wi);
    if (cosNI > 0 && cosNO > 0) {
        V3 H       = normalized(wi + wo);
        float dotx = l.tf.getx(H) / l.ax, doty = l.tf.gety(H) / l.ay, dotn = l.tf.getz(H);
        float oh   = dot(H, wi);
        float e    = fast_exp(-(dotx * dotx + doty * doty) / (dotn * dotn));
        float c    = float(4 * M_PI) * l.ax * l.ay;
        float k    = oh * dotn * dotn * dotn;
        float pdf  = e / (c * k);
        return BSample(wi, V3(k * std::sqrt(cosNI / cosNO)), pdf, std::max(l.ax, l.ay));
    }
    return BSample();
}
inline BSample ward_sample(const Lobe& l, const V3& wo, float rx, float ry)
{
    float cosNO = dot(l.N, wo);
    if (cosNO > 0) {
        float phi = 2 * float(M_PI) * rx;
        float sp, cp;
        fast_sincos(phi, &sp, &cp);
        float cosPhi = l.ax * cp, sinPhi = l.ay * sp;
        float k      = 1 / std::sqrt(cosPhi * cosPhi + sinPhi * sinPhi);
        cosPhi *= k;
        sinPhi *= k;
        float thetaDenom = (cosPhi * cosPhi) / (l.ax * l.ax) + (sinPhi * sinPhi) / (l.ay * l.ay);
        float tanTheta2  = -fast_log(1 - ry) / thetaDenom;
        float cosTheta   = 1 / std::sqrt(1 + tanTheta2);
        float sinTheta   = cosTheta * std::sqrt(tanTheta2);
        V3 h(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
        float dotx = h.x / l.ax, doty = h.y / l.ay, dotn = h.z;
        h           = l.tf.get(h.x, h.y, h.z);
        float oh    = dot(h, wo);
        V3 wi       = 2 * oh * h - wo;
        float cosNI = dot(l.N, wi);
        if (cosNI > 0) {
            float e   = fast_exp(-(dotx * dotx + doty * doty) / (dotn * dotn));
            float c   = float(4 * M_PI) * l.ax * l.ay;
            float kk  = oh * dotn * dotn * dotn;
            float pdf = e / (c * kk);
            return BSample(wi, V3(kk * std::sqrt(cosNI / cosNO)), pdf, std::max(l.ax, l.ay));
        }
    }
    return BSample();
}

// ---- libbsdl lobes behind testrender's BSDL_WRAP (shading.cpp:73-115) ----------------
// BSDL_WRAP builds bsdl::BsdfGlobals(wo, N, N, backfacing, path_roughness, 1, 0), the
// lobe frame is Frame(visible_normal(N)) and eval/sample run in that local frame
// (libbsdl/include/BSDL/bsdf_impl.h:48-76, tools.h:459-510).
inline float bsdl_max_abs_xyz(const V3& v) { return std::max(std::fabs(v.x), std::max(std::fabs(v.y), std::fabs(v.z))); }
inline float bsdl_clamp(float x, float a, float b)
{
    float m = std::max(x, a);
    return m < b ? m : b;
}
inline V3 bsdl_visible_normal(const V3& wo, const V3& Ngf, const V3& N)
{
    if (dot(wo, N) > 0.0f)
        return N;
    V3 V = cross(wo, N);
    if (bsdl_max_abs_xyz(V) < 1e-4f) {
        V = cross(wo, Ngf);
        if (bsdl_max_abs_xyz(V) < 1e-4f) {
            const float s = std::copysign(1.0f, wo.z);
            const float a = -1.0f / (s + wo.z);
            V             = V3(wo.x * wo.y * a, s + wo.y * wo.y * a, -wo.y);
        } else
            V = normalized(V);
    } else
        V = normalized(V);
    return normalized(cross(V, wo) + 1e-4f * wo);
}
// tools.h:200-278
inline float bsdl_fast_cos_quadrant(float x)
{
    const float c1 = 0.99998736f, c2 = -0.30837047f, c3 = 0.01578646f, c4 = -0.00029826362f;
    float x2 = x * x;
    float cp = c3 + c4 * x2;
    cp       = c2 + cp * x2;
    cp       = c1 + cp * x2;
    return cp;
}
inline float bsdl_fast_sin_quadrant(float x)
{
    const float s1 = 0.7853975892066955566406250000000000f, s2 = -0.0807407423853874206542968750000000f,
                s3 = 0.0024843954015523195266723632812500f, s4 = -0.0000341485538228880614042282104492f;
    float x2 = x * x;
    float sp = s3 + s4 * x2;
    sp       = s2 + sp * x2;
    sp       = s1 + sp * x2;
    return sp * x;
}
inline V3 bsdl_sample_cos_hemisphere(float randu, float randv)
{
    const float a = 2 * randu - 1, qa = std::fabs(a);
    const float b = 2 * randv - 1, qb = std::fabs(b);
    const float rad = qa > qb ? qa : qb;
    const float phi = qa > qb ? qb / qa : ((qa == qb) ? 1.0f : 2 - qa / qb);
    const float x   = std::copysign(rad * bsdl_fast_cos_quadrant(phi), a);
    const float y   = std::copysign(rad * bsdl_fast_sin_quadrant(phi), b);
    return V3(x, y, std::sqrt(1 - rad * rad));
}
// mtx::OrenNayarDiffuseLobe::eval_impl without energy compensation (the OREN_NAYAR_ID
// closure maps to {N, albedo 1, sigma, energy_compensation false}, shading.cpp:1496-1503);
// MTX/bsdf_oren_nayar_diffuse_impl.h:40-62.  wo, wi in the lobe frame.
inline BSample oren_nayar_eval_local(const Lobe& l, const V3& wo, const V3& wi)
{
    const float ONEOVERPI = 1 / float(M_PI);
    const float cosNI = bsdl_clamp(wi.z, 0.0f, 1.0f);
    const float cosNO = bsdl_clamp(wo.z, 0.0f, 1.0f);
    if (cosNI <= 0.0f || cosNO <= 0.0f)
        return BSample();
    const float cosIO = bsdl_clamp(dot(wo, wi), -1.0f, 1.0f);
    const float s     = cosIO - cosNI * cosNO;
    const float pdf   = cosNI * ONEOVERPI;
    if (!l.energy_compensation) {
        const float s2    = SQR(l.ax);
        const float A     = 1.0f - 0.50f * s2 / (s2 + 0.33f);
        const float B     = 0.45f * s2 / (s2 + 0.09f);
        const float stinv = s > 0.0f ? s / std::max(cosNI, cosNO) : 0.0f;
        const float f_ss  = A + B * stinv;
        return BSample(wi, l.albedo * f_ss, pdf, 1.0f);
    }
    // energy-preserving Oren-Nayar (Portsmouth), :63-93
    const float PI_F          = float(M_PI);
    const float constant1_FON = 0.5f - 2.0f / (3.0f * PI_F);
    const float constant2_FON = 2.0f / 3.0f - 28.0f / (15.0f * PI_F);
    const float sigma         = l.ax;
    auto E_FON_analytic = [&](float mu) {
        const float AF = 1.0f / (1.0f + constant1_FON * sigma);
        const float BF = sigma * AF;
        const float Si = std::sqrt(std::max(0.0f, 1.0f - mu * mu));
        const float G  = Si * (fast_acos(mu) - Si * mu)
                        + 2.0f * ((Si / std::max(mu, 1e-7f)) * (1.0f - Si * Si * Si) - Si) * (1.0f / 3.0f);
        return AF + (BF * ONEOVERPI) * G;
    };
    const float AF    = 1.0f / (1.0f + constant1_FON * sigma);
    const float stinv = s > 0.0f ? s / std::max(cosNI, cosNO) : s;
    const float f_ss  = AF * (1.0f + sigma * stinv);
    const float EFo   = E_FON_analytic(cosNO);
    const float EFi   = E_FON_analytic(cosNI);
    const float avgEF = AF * (1.0f + constant2_FON * sigma);
    auto rho = [&](float a) { return SQR(a) * avgEF / (1 - a * std::max(0.0f, 1.0f - avgEF)); };
    const V3 rho_ms(rho(l.albedo.x), rho(l.albedo.y), rho(l.albedo.z));
    const float f_ms = std::max(1e-7f, 1.0f - EFo) * std::max(1e-7f, 1.0f - EFi) / std::max(1e-7f, 1.0f - avgEF);
    return BSample(wi, l.albedo * f_ss + rho_ms * f_ms, pdf, 1.0f);
}
// mtx::BurleyDiffuseLobe (MTX/bsdf_burley_diffuse_impl.h:24-58)
inline float burley_fresnel(float cos_theta, float F90)
{
    const float x  = bsdl_clamp(1.0f - cos_theta, 0.0f, 1.0f);
    const float x2 = x * x;            // pown<5>: rec = pown<2>(x); rec * rec * x
    float f        = bsdl_clamp(x2 * x2 * x, 0.0f, 1.0f);   // LERP clamps its parameter
    return (1 - f) * 1.0f + f * F90;
}
inline BSample burley_eval_local(const Lobe& l, const V3& wo, const V3& wi)
{
    const float ONEOVERPI = 1 / float(M_PI);
    if (wo.z <= 0.0f || wi.z <= 0.0f)
        return BSample();
    const V3 H = wi + wo;
    if (bsdl_max_abs_xyz(H) < 1e-4f)
        return BSample();
    const V3 Hn       = normalized(H);
    const float cosHI = bsdl_clamp(dot(wi, Hn), 0.0f, 1.0f);
    const float cosNO = bsdl_clamp(wo.z, 0.0f, 1.0f);
    const float cosNI = bsdl_clamp(wi.z, 0.0f, 1.0f);
    const float F90   = 0.5f + 2.0f * l.ax * SQR(cosHI);
    const float refL  = burley_fresnel(cosNI, F90);
    const float refV  = burley_fresnel(cosNO, F90);
    const float pdf   = cosNI * ONEOVERPI;
    return BSample(wi, l.albedo * (refL * refV), pdf, 1.0f);
}
inline BSample bsdl_diffuse_eval_local(const Lobe& l, const V3& wo, const V3& wi)
{
    return l.type == LOBE_BSDL_BURLEY ? burley_eval_local(l, wo, wi) : oren_nayar_eval_local(l, wo, wi);
}
// BSDL_WRAP::eval / ::sample (shading.cpp:88-103): both diffuse lobes sample the cosine lobe
inline BSample bsdl_diffuse_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    BSample s = bsdl_diffuse_eval_local(l, l.tf.tolocal(wo), l.tf.tolocal(wi));
    return BSample(wi, s.weight, s.pdf, s.roughness);
}
inline BSample bsdl_diffuse_sample(const Lobe& l, const V3& wo, float rx, float ry)
{
    const V3 wo_l = l.tf.tolocal(wo);
    BSample s;
    if (!(wo_l.z <= 0.0f))
        s = bsdl_diffuse_eval_local(l, wo_l, bsdl_sample_cos_hemisphere(rx, ry));
    return BSample(l.tf.toworld(s.wi), s.weight, s.pdf, s.roughness);
}

// ---- mtx::SheenLobe, Conty-Kulla mode (MTX/bsdf_sheen_impl.h:17-175) ----------------------
// Frame(Z = visible normal, X = wo) (tools.h:483-495)
inline TangentFrame bsdl_frame_zx(const V3& Z, const V3& Xin)
{
    if (bsdl_max_abs_xyz(Xin) < 1e-4f || std::fabs(dot(Z, normalized(Xin))) > 0.999f)
        return TangentFrame::from_normal(Z);
    TangentFrame f;
    f.w = Z;
    f.v = normalized(cross(Z, Xin));
    f.u = cross(f.v, Z);
    return f;
}
inline V3 bsdl_sample_uniform_hemisphere(float randu, float randv)
{
    const float a = 2 * randu - 1, qa = std::fabs(a);
    const float b = 2 * randv - 1, qb = std::fabs(b);
    const float rad = qa > qb ? qa : qb;
    const float phi = qa > qb ? qb / qa : ((qa == qb) ? 1.0f : 2 - qa / qb);
    const float x   = std::copysign(rad * bsdl_fast_cos_quadrant(phi), a);
    const float y   = std::copysign(rad * bsdl_fast_sin_quadrant(phi), b);
    const float cos_theta = 1 - rad * rad;
    const float sin_theta = std::sqrt(2 - rad * rad);
    return V3(sin_theta * x, sin_theta * y, cos_theta);
}
// ContyKullaSheenMTX::albedo: rational fit in (cosNO, roughness)
inline float sheen_conty_albedo(float cosNO, float rough)
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
// cosine; the (A, B, R) coefficients come from a 32 x 32 table over (roughness, cos theta_o)
// that follows the energy tables in the LUT block (data/zeltner_ltc.bin)
enum { LUT_ZELTNER = LUT_WORDS, LUT_WORDS_ALL = LUT_THINLAYER + 32 * 16 * 16 };
inline V3 zeltner_fetch_coeffs(float roughness, float cosNO)
{
    const float ALMOSTONE = 0.999999940395355224609375f;
    const int ltc_res     = 32;
    float row = bsdl_clamp(roughness, 0.0f, ALMOSTONE) * (ltc_res - 1);
    float col = bsdl_clamp(cosNO, 0.0f, ALMOSTONE) * (ltc_res - 1);
    float r   = std::floor(row);
    float c   = std::floor(col);
    float rf  = row - r;
    float cf  = col - c;
    int ri    = (int)r;
    int ci    = (int)c;
    const float* T = bsdl_luts() + LUT_ZELTNER;
    auto at = [&](int rr, int cc) { const float* q = T + 3 * (rr * 32 + cc); return V3(q[0], q[1], q[2]); };
    auto lerp3 = [](float f, const V3& a, const V3& b) { return (1 - f) * a + f * b; };
    return lerp3(rf, lerp3(cf, at(ri, ci), at(ri, ci + 1)), lerp3(cf, at(ri + 1, ci), at(ri + 1, ci + 1)));
}
inline BSample zeltner_eval_ltc(float roughness, const V3& wi, const V3& ltc)
{
    const float a_inv = ltc.x, b_inv = ltc.y, r_coeff = ltc.z;
    const V3 wi_orig(a_inv * wi.x + b_inv * wi.z, a_inv * wi.y, wi.z);
    const float q        = a_inv / dot(wi_orig, wi_orig);
    const float jacobian = q * q;
    const float pdf      = jacobian * std::max(wi_orig.z, 0.0f) * (1 / float(M_PI));
    if (pdf > std::numeric_limits<float>::min())
        return BSample(wi, V3(r_coeff), pdf, roughness);
    return BSample();
}
// set up everything SheenLobe's constructor derives (sheen_alpha in ax, regularized
// roughness in ay, Emiss in emiss, mode in refract)
inline void sheen_setup(Lobe& l, const V3& wo, float roughness_param, bool backfacing, float path_roughness, int mode = 0)
{
    const V3 Z   = bsdl_visible_normal(wo, l.N, l.N);
    l.tf         = bsdl_frame_zx(Z, wo);
    const float r = bsdl_clamp(roughness_param, 0.0f, 1.0f);
    l.ay         = 1.0f - (1.0f - r) * (1.0f - path_roughness);   // regularize_roughness
    l.refract    = mode == 1;                                      // SheenLobe::ZELTNER
    // MIN_ROUGHNESS of ZeltnerBurleySheen / ContyKullaDist
    l.ax         = l.refract ? std::max(0.02f, std::sqrt(l.ay)) : std::max(0.06f, l.ay);
    l.backfacing = backfacing;
    const float cosNO = bsdl_clamp(dot(Z, wo), 0.0f, 1.0f);
    const float tmax  = std::max(l.albedo.x, std::max(l.albedo.y, l.albedo.z));
    if (backfacing)
        l.emiss = 1.0f;
    else if (l.refract)
        l.emiss = 1 - std::min(zeltner_fetch_coeffs(bsdl_clamp(l.ax, 0.02f, 1.0f), cosNO).z * tmax, 1.0f);
    else
        l.emiss = 1 - std::min(sheen_conty_albedo(cosNO, bsdl_clamp(l.ax, 0.06f, 1.0f)) * tmax, 1.0f);
}
// SheenMicrofacet<ContyKullaDist<false>>::eval
inline BSample sheen_conty_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    const float PI_F = float(M_PI), ONEOVERPI = 1 / float(M_PI);
    const float cosNO = wo.z, cosNI = wi.z;
    if (cosNI <= 1e-5f || cosNO <= 1e-5f)
        return BSample();
    const float a   = bsdl_clamp(l.ax, 0.06f, 1.0f);
    const V3 Hr     = normalized(wo + wi);
    float cos_theta = bsdl_clamp(Hr.z, 0.0f, 1.0f);
    float sin_theta = std::sqrt(1.0f - SQR(cos_theta));
    const float D   = fast_safe_pow(sin_theta, 1 / a) * (2 + 1 / a) * 0.5f * ONEOVERPI;
    if (D < 1e-6)
        return BSample();
    float cI = std::min(1.0f, wi.z), cO = std::min(1.0f, wo.z);
    const float G2 = (cI * cO) / (cI + cO - cI * cO);
    return BSample(wi, V3(D * G2 * 0.5f * PI_F / cosNO), 0.5f * ONEOVERPI, 0);
}
inline BSample sheen_eval_local(const Lobe& l, const V3& wo, const V3& wi)
{
    const float cosNO = wo.z, cosNI = wi.z;
    const bool is_reflection = cosNI > 0 && cosNO >= 0;
    BSample s;
    if (is_reflection && !l.backfacing) {
        if (l.refract) {
            const float rough = bsdl_clamp(l.ax, 0.02f, 1.0f);
            if (!(wo.z < 0 || wi.z <= 0))
                s = zeltner_eval_ltc(rough, wi, zeltner_fetch_coeffs(rough, wo.z));
        } else
            s = sheen_conty_eval(l, wo, wi);
        s.weight    = s.weight * l.albedo;
        s.roughness = l.ay;
    }
    return s;
}
inline BSample sheen_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    BSample s = sheen_eval_local(l, l.tf.tolocal(wo), l.tf.tolocal(wi));
    return BSample(wi, s.weight, s.pdf, s.roughness);
}
inline BSample sheen_sample(const Lobe& l, const V3& wo, float rx, float ry)
{
    BSample s;
    if (!l.backfacing) {
        const V3 wo_l = l.tf.tolocal(wo);
        if (l.refract) {
            // cosine base distribution transformed by M
            const float rough = bsdl_clamp(l.ax, 0.02f, 1.0f);
            if (!(wo_l.z < 0)) {
                const V3 ltc = zeltner_fetch_coeffs(rough, wo_l.z);
                const V3 o   = bsdl_sample_cos_hemisphere(rx, ry);
                const V3 wi(o.x - o.z * ltc.y, o.y, ltc.x * o.z);
                s = zeltner_eval_ltc(rough, normalized(wi), ltc);
            }
        } else
            // SheenMicrofacet::sample -> eval
            s = sheen_conty_eval(l, wo_l, bsdl_sample_uniform_hemisphere(rx, ry));
        // SheenLobe::sample_impl scales and tags it
        s.weight    = s.weight * l.albedo;
        s.roughness = l.ay;
    }
    return BSample(l.tf.toworld(s.wi), s.weight, s.pdf, s.roughness);
}

#include "osl_oracle_thinlayer.h"

}  // namespace lobes

inline V3 ext_albedo(const Lobe& l, const V3& wo)
{
    return l.type == LOBE_MICROFACET ? lobes::mf_albedo(l, wo) : V3(1.0f);
}
inline BSample ext_eval(const Lobe& l, const V3& wo, const V3& wi)
{
    switch (l.type) {
    case LOBE_PHONG: return lobes::phong_eval(l, wo, wi);
    case LOBE_WARD: return lobes::ward_eval(l, wo, wi);
    case LOBE_MICROFACET: return lobes::mf_eval(l, wo, wi);
    case LOBE_BSDL_OREN_NAYAR:
    case LOBE_BSDL_BURLEY: return lobes::bsdl_diffuse_eval(l, wo, wi);
    case LOBE_BSDL_SHEEN: return lobes::sheen_eval(l, wo, wi);
    case LOBE_MX_SPEC: {   // BSDL_WRAP::eval (shading.cpp:88-93)
        BSample s = mx_eval_local(l.mx, l.tf.tolocal(wo), l.tf.tolocal(wi));
        return BSample(wi, s.weight, s.pdf, s.roughness);
    }
    case LOBE_SPI_THINLAYER: {   // SpiThinLayer::eval (shading.cpp:138-143)
        BSample s = lobes::thin_eval_local(l.thin, l.tf.tolocal(wo), l.tf.tolocal(wi));
        return BSample(wi, s.weight, s.pdf, s.roughness);
    }
    case LOBE_MX_TRANSLUCENT: {
        const V3 wi_l = l.tf.tolocal(wi);
        if (wi_l.z >= 0.0f)
            return BSample(wi, V3(0.0f), 0, 0);
        return BSample(wi, l.albedo, std::fabs(wi_l.z) * (1 / float(M_PI)), 1.0f);
    }
    }
    return BSample();
}
inline BSample ext_sample(const Lobe& l, const V3& wo, float rx, float ry, float rz)
{
    switch (l.type) {
    case LOBE_PHONG: return lobes::phong_sample(l, wo, rx, ry);
    case LOBE_WARD: return lobes::ward_sample(l, wo, rx, ry);
    case LOBE_MICROFACET: return lobes::mf_sample(l, wo, rx, ry, rz);
    case LOBE_BSDL_OREN_NAYAR:
    case LOBE_BSDL_BURLEY: return lobes::bsdl_diffuse_sample(l, wo, rx, ry);
    case LOBE_BSDL_SHEEN: return lobes::sheen_sample(l, wo, rx, ry);
    case LOBE_MX_SPEC: {   // BSDL_WRAP::sample (shading.cpp:94-101)
        BSample s = mx_sample_local(l.mx, l.tf.tolocal(wo), rx, ry, rz);
        return BSample(l.tf.toworld(s.wi), s.weight, s.pdf, s.roughness);
    }
    case LOBE_SPI_THINLAYER: {   // SpiThinLayer::sample (shading.cpp:144-151)
        BSample s = lobes::thin_sample_local(l.thin, l.tf.tolocal(wo), V3(rx, ry, rz));
        return BSample(l.tf.toworld(s.wi), s.weight, s.pdf, s.roughness);
    }
    case LOBE_MX_TRANSLUCENT: {
        V3 wi_l = lobes::bsdl_sample_cos_hemisphere(rx, ry);
        wi_l.z  = -wi_l.z;
        if (wi_l.z >= 0.0f)
            return BSample(l.tf.toworld(V3(0.0f)), V3(0.0f), 0, 0);
        return BSample(l.tf.toworld(wi_l), l.albedo, std::fabs(wi_l.z) * (1 / float(M_PI)), 1.0f);
    }
    }
    return BSample();
}

}  // namespace oslo
