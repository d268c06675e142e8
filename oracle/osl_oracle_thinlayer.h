// osl_oracle_thinlayer.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of spi::ThinLayerLobe, the closure testrender registers as "thinlayer"
// (src/testrender/shading.cpp:119-152, SpiThinLayer) with three RGB channels:
//   spi::ThinFresnel, ThinMicrofacet<GGXDist>, Thinlayer, ThinLayerLobe
//                                    src/libbsdl/include/BSDL/SPI/bsdf_thinlayer_{decl,impl}.h
//   GGXDist::sample (visible normals, Walter's ellipsoid form)   BSDL/microfacet_tools_impl.h:48-70
//   Sample::update, Sample::stretch                               BSDL/bsdf_impl.h:17-46
//   fresnel_dielectric, avg_fresnel_dielectric, sum_max           BSDL/tools.h:159-164, 345-415
// Checked value by value against the reference's own class (oracle/_ref/libref_bsdl.so, lobe 7,
// tests/test_oracle_bsdl.py).  The spi::Thinlayer energy table is data: baked by
// tools/bake_bsdl_luts.cpp, checked against the table in the reference's headers.
// Included from osl_oracle_lobes.h inside namespace oslo::lobes (after the bsdl_* helpers).
#pragma once

inline float thin_fresnel_dielectric(float cosi, float eta)
{
    if (eta == 0.0f)
        return 1.0f;
    if (cosi < 0.0f)
        eta = 1.0f / eta;
    float c = std::fabs(cosi);
    float g = eta * eta - 1 + c * c;
    if (g > 0) {
        g       = sqrtf(g);
        float A = (g - c) / (g + c);
        float B = (c * (g + c) - 1) / (c * (g - c) + 1);
        return 0.5f * A * A * (1 + B * B);
    }
    return 1.0f;
}
inline float thin_avg_fresnel_dielectric(float eta)
{
    if (eta < 1)
        return 0.997118f + eta * (0.1014f + eta * (-0.965241f - eta * 0.130607f));
    return (eta - 1) / (4.08567f + 1.00071f * eta);
}
// ThinFresnel (eta already clamped to [1.001, 5] by thin_setup)
inline float thin_F(float eta, float c)
{
    const float g = sqrtf(eta * eta - 1 + c * c);
    const float A = (g - c) / (g + c);
    const float B = (c * (g + c) - 1) / (c * (g - c) + 1);
    return 0.5f * A * A * (1 + B * B);
}
inline float thin_F_inv(float eta, float c) { return thin_fresnel_dielectric(c, 1 / eta); }
inline float thin_F_avg(float eta) { return (eta - 1) / (4.08567f + 1.00071f * eta); }
inline float thin_F_avg_inv(float eta) { return thin_avg_fresnel_dielectric(1 / eta); }
inline float thin_table_index(float eta)
{
    const float IOR_MIN = 1.001f, IOR_MAX = 5.0f;
    const float seta = mx_clamp(eta < 1 ? 1 / eta : eta, IOR_MIN, IOR_MAX);
    const float x    = (seta - IOR_MIN) * (1 / (IOR_MAX - IOR_MIN));
    return sqrtf(x);
}
// ThinMicrofacet::sum_refl_series / sum_refr_series: the geometric series of the bounces inside the layer
inline float thin_sum_refl(float Rout, float Tin, float Rin, float A)
{
    const float b = mx_sqr(Rin * A);
    return Rout + (1 - b < 1e-4f ? (A < 1 ? 0 : Tin * 0.5f) : Tin * (1 - Rin) * mx_sqr(A) * Rin / (1 - b));
}
inline float thin_sum_refr(float Rout, float Tin, float Rin, float A)
{
    (void)Rout;
    const float b = mx_sqr(Rin * A);
    return 1 - b < 1e-4f ? (A < 1 ? 0 : Tin * 0.5f) : Tin * (1 - Rin) * A / (1 - b);
}
// bsdl::Sample::update: fold another technique (weight ow, pdf opdf, chosen with probability cpdf) into s
inline void thin_update(BSample& s, V3 ow, float opdf, float cpdf)
{
    if (cpdf > 1e-6f) {
        opdf *= cpdf;
        ow = ow * (1 / cpdf);
        float f;
        if (opdf == s.pdf) {
            f        = mx_clamp(0.5f, 0.0f, 1.0f);
            s.weight = (1 - f) * s.weight + f * ow;
        } else if (opdf < s.pdf) {
            f        = mx_clamp(1 / (1 + opdf / s.pdf), 0.0f, 1.0f);
            s.weight = (1 - f) * ow + f * s.weight;
        } else {
            f        = mx_clamp(1 / (1 + s.pdf / opdf), 0.0f, 1.0f);
            s.weight = (1 - f) * s.weight + f * ow;
        }
        s.pdf += opdf;
    }
}
inline float thin_stretch(float x, float min, float length)
{
    return std::min((x - min) / length, 0.999999940395355224609375f);
}
inline float thin_sum_max(float a, float b, float smax)
{
    const float maxab = std::max(a, b), minab = std::min(a, b);
    return maxab + (smax - maxab) * (minab / smax);
}
// GGXDist::sample
inline V3 thin_ggx_sample(const GGXD& d, const V3& wo, float randu, float randv)
{
    const V3 V  = normalized(V3(d.ax * wo.x, d.ay * wo.y, wo.z));
    const V3 T1 = V.z < 0.9999f ? normalized(V3(V.y, -V.x, 0)) : V3(1, 0, 0);
    const V3 T2 = cross(T1, V);
    const V3 p  = bsdl_sample_cos_hemisphere(randu, randv);   // .x, .y = square_to_unit_disc
    const float s   = 0.5f * (1 + V.z);
    const float p2o = s * p.y + (1 - s) * sqrtf(1 - p.x * p.x);
    const float p3  = sqrtf(std::max(1.0f - mx_sqr(p.x) - mx_sqr(p2o), 0.0f));
    const V3 N      = p.x * T1 + p2o * T2 + p3 * V;
    return normalized(V3(d.ax * N.x, d.ay * N.y, std::max(N.z, 0.0f)));
}
inline float thin_fresnel_prob(const ThinSpec& t, float f)
{
    const float safe_prob = 0.2f;
    return mx_lerp(t.prob_clamp, f, mx_clamp(f, safe_prob, 1 - safe_prob));
}
inline float thin_refraction_slope_scale(const ThinSpec& t, float cosNO)
{
    cosNO = std::min(cosNO, 1.0f);
    const float inveta = 1 / t.eta;
    const float sinNI  = inveta * sqrtf(1 - mx_sqr(cosNO));
    if (sinNI > 1.0f)
        return 1;
    const float cosNI = -sqrtf(1 - mx_sqr(sinNI));
    const float refr_jacobian_entry = (1 + inveta * (cosNO / cosNI));
    const float refr_jacobian_exit  = (1 + t.eta * (cosNI / cosNO));
    const float refl_jacobian       = 2;
    const float refr_variance       = mx_sqr(refr_jacobian_entry) + mx_sqr(refr_jacobian_exit);
    const float refl_variance       = mx_sqr(refl_jacobian);
    const float max_scale           = 1 / t.roughness;
    return std::min(sqrtf((refr_variance / refl_variance) / cosNO), max_scale);
}
inline V3 thin_scale_slopes(const V3& m, float s) { return normalized(V3(m.x * s, m.y * s, m.z)); }
inline void thin_attenuation(const ThinSpec& t, const V3& wo, const V3& m, V3* refl, V3* refr)
{
    const V3 wr       = bsdl_refract(wo, m, t.eta);
    const float cosNO = std::min(dot(wo, m), 1.0f);
    const float cosNR = mx_clamp(-wr.z, 0.0f, 1.0f);
    const float d     = t.thickness * mx_lerp(t.roughness, 1 / std::max(cosNR, 1e-6f), 2.2f);
    const float Rout  = thin_F(t.eta, cosNO);
    const float Tin   = 1.0f - Rout;
    const float Rin   = mx_lerp(t.roughness, thin_F_inv(t.eta, cosNR), thin_F_avg_inv(t.eta));
    V3 A(1.0f);
    if (d > 0)
        for (int i = 0; i < 3; ++i)
            A[i] = t.sigma_t[i] > 0 ? fast_exp(-d * t.sigma_t[i]) : 1;
    for (int i = 0; i < 3; ++i) {
        (*refl)[i] = thin_sum_refl(Rout, Tin, Rin, A[i]);
        (*refr)[i] = thin_sum_refr(Rout, Tin, Rin, A[i]);
    }
}
// ThinMicrofacet::eval, the common tail (micronormal m known)
inline BSample thin_eval_m(const ThinSpec& t, const V3& wo, const V3& m, const V3& wi, bool both,
                           float refr_slope_scale, const V3& refl_atten, const V3& refr_atten)
{
    const bool isrefl = wi.z > 0;
    const float R = v3max(refl_atten), T = v3max(refr_atten);
    const float F = R / std::max(R + T, std::numeric_limits<float>::min());
    const float P = both ? thin_fresnel_prob(t, isrefl ? F : 1 - F) : 1;
    if (P < 1e-6f)
        return BSample();
    const V3 wif(wi.x, wi.y, std::fabs(wi.z));
    const float cosNO = wo.z;
    const float cosNM = m.z;
    const float sinNM = sqrtf(1 - std::min(mx_sqr(cosNM), 1.0f));
    const float tmp   = sqrtf(mx_sqr(refr_slope_scale * sinNM) + mx_sqr(cosNM));
    const float J     = isrefl ? 1 : tmp * tmp * tmp / mx_sqr(refr_slope_scale);
    const float D     = t.d.D(m);
    const float G1    = t.d.G1(wo);
    const float out   = t.d.G2_G1(wif, wo) / P;
    const float pdf   = (G1 * D * J * P) / (4.0f * cosNO);
    const V3 w        = (isrefl ? refl_atten : refr_atten) * out;
    return BSample(wi, w, pdf, 0);
}
inline BSample thin_spec_eval(const ThinSpec& t, const V3& wo, const V3& wi, bool doreflect, bool dorefract)
{
    const bool both   = doreflect && dorefract;
    const bool isrefl = wi.z > 0;
    const V3 wif(wi.x, wi.y, std::fabs(wi.z));
    const V3 mt                  = normalized(wo + wif);
    const float refr_slope_scale = thin_refraction_slope_scale(t, wo.z);
    const V3 m                   = isrefl ? mt : thin_scale_slopes(mt, 1 / refr_slope_scale);
    if (wi.z == 0 || wo.z == 0 || dot(m, wo) <= 0 || dot(mt, wo) <= 0)
        return BSample();
    V3 refl_atten, refr_atten;
    thin_attenuation(t, wo, m, &refl_atten, &refr_atten);
    return thin_eval_m(t, wo, m, wi, both, refr_slope_scale, refl_atten, refr_atten);
}
inline BSample thin_spec_sample(const ThinSpec& t, const V3& wo, float randu, float randv, float randw, bool doreflect,
                                bool dorefract)
{
    const bool both = doreflect && dorefract;
    const V3 m      = thin_ggx_sample(t.d, wo, randu, randv);
    if (dot(wo, m) <= 0)
        return BSample();
    V3 refl_atten, refr_atten;
    thin_attenuation(t, wo, m, &refl_atten, &refr_atten);
    const float R = v3max(refl_atten), T = v3max(refr_atten);
    const float F = R / std::max(R + T, std::numeric_limits<float>::min());
    const float P = both ? thin_fresnel_prob(t, F) : (dorefract ? 0 : 1);
    const bool isrefl            = randw < P;
    const float refr_slope_scale = thin_refraction_slope_scale(t, wo.z);
    const V3 mt                  = isrefl ? m : thin_scale_slopes(m, refr_slope_scale);
    if (dot(wo, mt) <= 0)
        return BSample();
    const V3 wif = bsdl_reflect(wo, mt);
    const V3 wi  = isrefl ? wif : V3(wif.x, wif.y, -wif.z);
    if ((isrefl && wi.z <= 0) || (!isrefl && wi.z >= 0))
        return BSample();
    return thin_eval_m(t, wo, m, wi, both, refr_slope_scale, refl_atten, refr_atten);
}
// ThinLayerLobe::get_diff_trans: tints of the diffuse / translucent energy-compensation lobes
inline void thin_diff_trans(const ThinSpec& t, V3* diff_tint, V3* trans_tint)
{
    const float Tf = 1 - thin_F_avg(t.eta), Tb = 1 - thin_F_avg_inv(t.eta);
    const float x    = Tb * mx_sqr(t.eta) / (Tf + Tb * mx_sqr(t.eta));
    const float Tout = Tb * (1 - x), Tin = Tf * x, Rout = 1 - Tin, Rin = 1 - Tout;
    const float avgd = t.thickness * 2.2f;
    for (int i = 0; i < 3; ++i) {
        const float A    = fast_exp(-avgd * t.sigma_t[i]);
        (*diff_tint)[i]  = thin_sum_refl(Rout, Tin, Rin, A);
        (*trans_tint)[i] = thin_sum_refr(Rout, Tin, Rin, A);
    }
}
inline void thin_eval_ec_lobe(const ThinSpec& t, BSample* s, const V3& wi_l, const V3& diff_tint, const V3& trans_tint,
                              float Tprob)
{
    const bool back       = wi_l.z <= 0;
    const float side_prob = back ? Tprob : 1 - Tprob;
    if (side_prob == 0)
        return;
    const float dpdf = t.Eo * std::fabs(wi_l.z) * (1 / float(M_PI));
    thin_update(*s, back ? trans_tint : diff_tint, dpdf, side_prob);
}
// ThinLayerLobe::ThinLayerLobe (everything but the frame)
inline ThinSpec thin_setup(float cosNO, float IOR, float roughness_param, float anisotropy, float thickness,
                           const V3& refl_tint, const V3& refr_tint, const V3& sigma_t, float path_roughness)
{
    ThinSpec t;
    const float r = 1.0f - (1.0f - mx_clamp(roughness_param, 0.0f, 1.0f)) * (1.0f - path_roughness);
    t.d           = GGXD(r, anisotropy);
    t.sigma_t     = sigma_t;
    t.eta         = mx_clamp(IOR, 1.001f, 5.0f);
    t.thickness   = thickness;
    t.roughness   = r;
    t.prob_clamp  = 0;   // "Not exposed": not a registered parameter, the closure leaves it zero
    t.refl_tint   = v3clamped(refl_tint, 0, 1);
    t.refr_tint   = v3clamped(refr_tint, 0, 1);
    EnergyCurve c { bsdl_luts() + LUT_THINLAYER, 32, false, r, thin_table_index(t.eta) };
    t.Eo = c.Emiss_eval(mx_clamp(cosNO, 0.0f, 1.0f));
    return t;
}
// ThinLayerLobe::eval_impl / sample_impl (local frame)
inline BSample thin_eval_local(const ThinSpec& t, const V3& wo, const V3& wi, bool doreflect = true, bool dorefract = true)
{
    const bool both   = doreflect && dorefract;
    const float cosNI = wi.z;
    V3 diff, trans;
    thin_diff_trans(t, &diff, &trans);
    const float R = v3max(diff), T = v3max(trans);
    const float Tprob = T / std::max(R + T, std::numeric_limits<float>::min());
    if ((cosNI > 0 && !doreflect) || (cosNI < 0 && !dorefract) || cosNI == 0)
        return BSample();
    BSample s(wi, V3(0.0f), 0, 0);
    const float PE = t.Eo * (both ? 1 : (dorefract ? Tprob : 1 - Tprob));
    BSample ss     = thin_spec_eval(t, wo, wi, doreflect, dorefract);
    thin_update(s, ss.weight, ss.pdf, 1 - PE);
    thin_eval_ec_lobe(t, &s, wi, diff, trans, Tprob);
    s.weight    = s.weight * (cosNI > 0 ? t.refl_tint : t.refr_tint);
    s.roughness = thin_sum_max(t.roughness, cosNI < 0 ? t.roughness : 0, 1.0f);
    return s;
}
inline BSample thin_sample_local(const ThinSpec& t, const V3& wo, V3 rnd, bool doreflect = true, bool dorefract = true)
{
    const bool both = doreflect && dorefract;
    V3 diff, trans;
    thin_diff_trans(t, &diff, &trans);
    const float R = v3max(diff), T = v3max(trans);
    const float Tprob = T / std::max(R + T, std::numeric_limits<float>::min());
    const float PE    = t.Eo * (both ? 1 : (dorefract ? Tprob : 1 - Tprob));
    BSample s;
    if (rnd.x < 1 - PE) {
        rnd.x      = thin_stretch(rnd.x, 0.0f, 1 - PE);
        BSample ss = thin_spec_sample(t, wo, rnd.x, rnd.y, rnd.z, doreflect, dorefract);
        if (mx_max_abs(ss.wi) == 0)
            return BSample();
        s.wi = ss.wi;
        thin_update(s, ss.weight, ss.pdf, 1 - PE);
        thin_eval_ec_lobe(t, &s, s.wi, diff, trans, Tprob);
    } else {
        rnd.x           = thin_stretch(rnd.x, 1 - PE, PE);
        const bool back = !(!dorefract || (both && rnd.z >= Tprob));
        s.wi            = bsdl_sample_cos_hemisphere(rnd.x, rnd.y);
        if (back)
            s.wi.z = -s.wi.z;
        thin_eval_ec_lobe(t, &s, s.wi, diff, trans, Tprob);
        const BSample ss = thin_spec_eval(t, wo, s.wi, doreflect, dorefract);
        thin_update(s, ss.weight, ss.pdf, 1 - PE);
    }
    const float cosNI = s.wi.z;
    s.weight    = s.weight * (cosNI > 0 ? t.refl_tint : t.refr_tint);
    s.roughness = thin_sum_max(t.roughness, cosNI < 0 ? t.roughness : 0, 1.0f);
    return s;
}
