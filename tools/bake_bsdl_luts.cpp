// bake_bsdl_luts.cpp — bakes the energy-compensation tables of the MaterialX microfacet closures
// (product build tool: writes openshadinglanguage_b200/data/bsdl_luts.bin).
//
// The reference bakes these tables at build time with src/libbsdl/src/genluts.cpp: for every
// (fresnel index, roughness index, cosine) it integrates the lobe's sample weight over 128x128
// stratified, scrambled samples and stores 1 - E.  This is a stand-alone restatement of that
// procedure for the four tables the conductor / dielectric / generalized-Schlick closures read:
//   [0]      spi::MiniMicrofacetGGX     1 x 16 x 16   (microfacet_tools_decl.h:63-101)
//   [256]    mtx::DielectricReflFront  32 x 16 x 16   (MTX/bsdf_dielectric_decl.h:58-94)
//   [8448]   mtx::DielectricBothFront  32 x 16 x 16
//   [16640]  mtx::DielectricBothBack   32 x 16 x 16
// genluts runs with libbsdl's default configuration (std::cos / std::sin, no fast math), and so
// does this tool.  tests/test_oracle_bsdl.py checks the file against the tables the reference's
// own genluts produces (oracle/_ref), entry by entry.
//
// and, into a second file, spi::Thinlayer 32 x 16 x 16 (SPI/bsdf_thinlayer_decl.h) for the thinlayer closure.
//
//   g++ -std=c++17 -O2 tools/bake_bsdl_luts.cpp -lpthread -o /tmp/bake && /tmp/bake bsdl_luts.bin thinlayer_lut.bin
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <thread>
#include <vector>

namespace {

struct V3 {
    float x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline V3 operator-(V3 a, V3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline V3 operator*(V3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline V3 operator*(float s, V3 a) { return { a.x * s, a.y * s, a.z * s }; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
// Imath::Vec3::length / normalized
inline float length(V3 v)
{
    float l2 = dot(v, v);
    if (l2 < 2.0f * FLT_MIN) {
        float ax = std::fabs(v.x), ay = std::fabs(v.y), az = std::fabs(v.z);
        float m = std::max(ax, std::max(ay, az));
        if (m == 0.0f)
            return 0.0f;
        ax /= m; ay /= m; az /= m;
        return m * std::sqrt(ax * ax + ay * ay + az * az);
    }
    return std::sqrt(l2);
}
inline V3 normalized(V3 v)
{
    float l = length(v);
    if (l == 0.0f)
        return { 0, 0, 0 };
    return { v.x / l, v.y / l, v.z / l };
}
inline float SQR(float x) { return x * x; }
inline float CLAMP(float x, float a, float b) { return std::min(std::max(x, a), b); }
inline float LERP(float f, float a, float b)
{
    f = CLAMP(f, 0.0f, 1.0f);
    return (1 - f) * a + f * b;
}
constexpr float PI = float(M_PI);

// tools.h:200-225, 532-552: the concentric square -> disc map with its quadrant polynomials
inline float fast_cos_quadrant(float x)
{
    const float x2 = x * x;
    float c        = 0.01578646f + -0.00029826362f * x2;
    c              = -0.30837047f + c * x2;
    c              = 0.99998736f + c * x2;
    return c;
}
inline float fast_sin_quadrant(float x)
{
    const float x2 = x * x;
    float s        = 0.0024843954015523195266723632812500f + -0.0000341485538228880614042282104492f * x2;
    s              = -0.0807407423853874206542968750000000f + s * x2;
    s              = 0.7853975892066955566406250000000000f + s * x2;
    return s * x;
}
inline void square_to_unit_disc(float rx, float ry, float& x, float& y)
{
    const float a = 2 * rx - 1, qa = std::fabs(a);
    const float b = 2 * ry - 1, qb = std::fabs(b);
    const float rad = qa > qb ? qa : qb;
    const float phi = qa > qb ? qb / qa : ((qa == qb) ? 1.0f : 2 - qa / qb);
    x = copysignf(rad * fast_cos_quadrant(phi), a);
    y = copysignf(rad * fast_sin_quadrant(phi), b);
}
inline V3 reflect(V3 E, V3 N) { return N * (2 * dot(N, E)) - E; }
inline V3 refract(V3 E, V3 N, float eta)
{
    V3 R { 0, 0, 0 };
    if (eta == 0)
        return R;
    V3 Nn;
    float cosi = dot(E, N), neta;
    if (cosi > 0) {
        neta = 1 / eta;
        Nn   = N;
    } else {
        cosi = -cosi;
        neta = eta;
        Nn   = { -N.x, -N.y, -N.z };
    }
    float arg = 1 - (neta * neta * (1 - (cosi * cosi)));
    if (arg >= 0) {
        float dnp = sqrtf(arg);
        float nK  = (neta * cosi) - dnp;
        R         = normalized(E * (-neta) + Nn * nK);
    }
    return R;
}

struct GGX {   // GGXDist (microfacet_tools_impl.h:19-107)
    float ax, ay;
    explicit GGX(float rough)
    {
        ax = ay = SQR(rough);
        ax = std::max(ax * (1 + 0.0f), 1e-5f);
        ay = std::max(ay * (1 - 0.0f), 1e-5f);
    }
    float D(V3 Hr) const
    {
        const float cosPhi2st2 = SQR(Hr.x / ax);
        const float sinPhi2st2 = SQR(Hr.y / ay);
        const float cosThetaM2 = SQR(Hr.z);
        const float sinThetaM2 = cosPhi2st2 + sinPhi2st2;
        return 1.0f / (PI * ax * ay * SQR(cosThetaM2 + sinThetaM2));
    }
    float G1(V3 w) const
    {
        w = { w.x * ax, w.y * ay, w.z };
        return 2.0f * w.z / (w.z + length(w));
    }
    float G2_G1(V3 wi, V3 wo) const
    {
        wi = { wi.x * ax, wi.y * ay, wi.z };
        wo = { wo.x * ax, wo.y * ay, wo.z };
        const float nl = length(wi), nv = length(wo);
        return wi.z * (wo.z + nv) / (wo.z * nl + wi.z * nv);
    }
    V3 sample(V3 wo, float randu, float randv) const   // visible normals (Walter's ellipsoid trick)
    {
        const V3 V  = normalized({ ax * wo.x, ay * wo.y, wo.z });
        const V3 T1 = V.z < 0.9999f ? normalized({ V.y, -V.x, 0 }) : V3 { 1, 0, 0 };
        const V3 T2 = cross(T1, V);
        float px, py;
        square_to_unit_disc(randu, randv, px, py);
        const float s   = 0.5f * (1 + V.z);
        const float p2o = s * py + (1 - s) * sqrtf(1 - px * px);
        const float p3  = sqrtf(std::max(1.0f - SQR(px) - SQR(p2o), 0.0f));
        const V3 N      = px * T1 + p2o * T2 + p3 * V;
        return normalized({ ax * N.x, ay * N.y, std::max(N.z, 0.0f) });
    }
    V3 sample_for_refl(V3 wo, float randu, float randv) const   // bounded VNDF (Eto & Tokuyoshi)
    {
        V3 i_std        = normalized({ wo.x * ax, wo.y * ay, wo.z });
        const float phi = 2.0f * PI * randu;
        const float a   = CLAMP(std::min(ax, ay), 0.0f, 1.0f);
        const float s   = 1 + sqrtf(SQR(wo.x) + SQR(wo.y));
        const float a2 = SQR(a), s2 = SQR(s);
        const float k  = (1 - a2) * s2 / (s2 + a2 * SQR(wo.z));
        const float b  = k * i_std.z;
        const float z  = (1 - randv) * (1 + b) - b;
        const float sinTheta = sqrtf(CLAMP(1 - SQR(z), 0.0f, 1.0f));
        V3 o_std = { sinTheta * std::cos(phi), sinTheta * std::sin(phi), z };
        V3 m_std = i_std + o_std;
        return normalized({ m_std.x * ax, m_std.y * ay, m_std.z });
    }
    float D_refl_D(V3 wo) const
    {
        const float len2 = SQR(wo.x * ax) + SQR(wo.y * ay);
        const float t    = sqrtf(len2 + SQR(wo.z));
        const float a    = CLAMP(std::min(ax, ay), 0.0f, 1.0f);
        const float s    = 1 + sqrtf(SQR(wo.x) + SQR(wo.y));
        const float a2 = SQR(a), s2 = SQR(s);
        const float k  = (1 - a2) * s2 / (s2 + a2 * SQR(wo.z));
        return 2 * wo.z / (k * wo.z + t);
    }
};

// TabulatedEnergyCurve<spi::MiniMicrofacetGGX>::Emiss_eval on the table baked first
struct GGXCurve {
    const float* E;
    float roughness;
    static float get_cosine(int i) { return std::max(SQR(float(i) * (1.0f / 15)), 1e-6f); }
    float interp(int i) const
    {
        float rf = roughness * 15;
        int ra   = static_cast<int>(rf);
        int rb   = std::min(ra + 1, 15);
        rf -= ra;
        return LERP(rf, E[ra * 16 + i], E[rb * 16 + i]);
    }
    float eval(float c) const
    {
        float cos0 = get_cosine(0);
        if (c <= cos0)
            return interp(0);
        for (int i = 1; i < 16; i++) {
            const float cos1 = get_cosine(i);
            if (c < cos1) {
                float q = (c - cos0) / (cos1 - cos0);
                return LERP(q, interp(i - 1), interp(i));
            }
            cos0 = cos1;
        }
        return interp(15);
    }
};

// mtx::DielectricFresnel (MTX/bsdf_dielectric_impl.h:19-66)
struct Fresnel {
    float eta;
    static Fresnel from_table_index(float tx, bool backside)
    {
        const float IOR_MIN = 1.001f, IOR_MAX = 5.0f;
        float e = LERP(SQR(tx), IOR_MIN, IOR_MAX);
        if (backside)
            e = 1 / e;
        Fresnel f;
        f.eta = e >= 1 ? CLAMP(e, IOR_MIN, IOR_MAX) : CLAMP(e, 1 / IOR_MAX, 1 / IOR_MIN);
        return f;
    }
    float eval(float c) const
    {
        float g = (eta - 1.0f) * (eta + 1.0f) + c * c;
        if (g > 0) {
            g       = sqrtf(g);
            float A = (g - c) / (g + c);
            float B = (c * (g + c) - 1) / (c * (g - c) + 1);
            return 0.5f * A * A * (1 + B * B);
        }
        return 1.0f;
    }
};

// the largest channel of the sample weight of each baked BSDF (all channels are equal here)
struct MiniMicrofacetGGX {
    GGX d;
    MiniMicrofacetGGX(float, float rough, float, const float*) : d(rough) {}
    float sample_weight(V3 wo, float ru, float rv, float) const
    {
        const V3 m = d.sample(wo, ru, rv);
        if (dot(m, wo) > 0) {
            const V3 wi = reflect(wo, m);
            if (wi.z > 0)
                return d.G2_G1(wi, wo);
        }
        return 0.0f;
    }
};
template<bool DOREFR, bool BACK> struct Dielectric {   // DielectricBSDF<DielectricFresnel>
    GGX d;
    Fresnel f;
    float E_ms = 0;
    Dielectric(float cosNO, float rough, float fresnel_index, const float* ggx_table)
        : d(rough), f(Fresnel::from_table_index(fresnel_index, BACK))
    {
        if (!DOREFR)
            E_ms = GGXCurve { ggx_table, rough }.eval(cosNO);
    }
    float eval_weight(V3 wo, V3 wi) const
    {
        const float cosNO = wo.z, cosNI = wi.z;
        if (!DOREFR) {   // eval_turquin_microms_reflection
            if (cosNI <= 0 || cosNO <= 0)
                return 0.0f;
            V3 m            = normalized(wo + wi);
            float cosMO     = dot(m, wo);
            float D_refl_D  = d.D_refl_D(wo);
            const float G1  = d.G1(wo);
            const float out = d.G2_G1(wi, wo) * G1 / D_refl_D;
            const float F   = f.eval(cosMO);
            const float msf = E_ms / std::max(0.01f, 1 - E_ms);
            return out * F * (1.0f + F * msf);
        }
        if (cosNI > 0) {
            const V3 m        = normalized(wo + wi);
            const float cosMO = dot(m, wo);
            if (cosMO <= 0)
                return 0.0f;
            const float G1 = d.G1(wo);
            const float F  = f.eval(cosMO);
            if (F <= 0)
                return 0.0f;
            const float D_refl_D = d.D_refl_D(wo);
            return F * (d.G2_G1(wi, wo) * G1 / (D_refl_D * F));
        } else if (cosNI < 0) {
            const V3 Ht       = normalized(f.eta * wi + wo) * ((f.eta > 1) ? -1.0f : 1.0f);
            const float cosHO = dot(Ht, wo), cosHI = dot(Ht, wi);
            if (cosHO <= 0 || cosHI >= 0)
                return 0.0f;
            const float Ft = 1.0f - f.eval(cosHO);
            if (Ht.z <= 0 || Ft <= 0)
                return 0.0f;
            const float G1       = d.G1(wo);
            const float D_refl_D = d.D_refl_D(wo);
            return Ft * (d.G2_G1({ wi.x, wi.y, -wi.z }, wo) * G1 / (D_refl_D * Ft));
        }
        return 0.0f;
    }
    float sample_weight(V3 wo, float ru, float rv, float rw) const
    {
        if (!DOREFR) {   // sample_turquin_microms_reflection
            if (wo.z <= 0)
                return 0.0f;
            V3 m  = d.sample_for_refl(wo, ru, rv);
            V3 wi = reflect(wo, m);
            if (wi.z <= 0)
                return 0.0f;
            return eval_weight(wo, wi);
        }
        V3 m              = d.sample_for_refl(wo, ru, rv);
        const float cosMO = dot(wo, m);
        if (cosMO <= 0)
            return 0.0f;
        const float F       = f.eval(cosMO);
        bool choose_reflect = rw < F;
        const V3 wi         = choose_reflect ? reflect(wo, m) : refract(wo, m, f.eta);
        if ((choose_reflect && wi.z <= 0) || (!choose_reflect && wi.z >= 0))
            return 0.0f;
        return eval_weight(wo, wi);
    }
};

// spi::Thinlayer as genluts builds it: Thinlayer(0, roughness_index, fresnel_index), i.e. thickness 0,
// no extinction, prob_clamp 0, isotropic (SPI/bsdf_thinlayer_impl.h:114-127, 135-178, 187-310)
inline float fresnel_dielectric(float cosi, float eta)
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
struct ThinlayerBake {
    GGX d;
    float eta, roughness;
    ThinlayerBake(float, float rough, float fresnel_index, const float*) : d(rough), roughness(rough)
    {
        eta = CLAMP(LERP(SQR(fresnel_index), 1.001f, 5.0f), 1.001f, 5.0f);
    }
    float F(float c) const
    {
        const float g = sqrtf(eta * eta - 1 + c * c);
        const float A = (g - c) / (g + c);
        const float B = (c * (g + c) - 1) / (c * (g - c) + 1);
        return 0.5f * A * A * (1 + B * B);
    }
    float avg_invf() const
    {
        const float e = 1 / eta;
        if (e < 1)
            return 0.997118f + e * (0.1014f + e * (-0.965241f - e * 0.130607f));
        return (e - 1) / (4.08567f + 1.00071f * e);
    }
    float slope_scale(float cosNO) const
    {
        cosNO = std::min(cosNO, 1.0f);
        const float inveta = 1 / eta;
        const float sinNI  = inveta * sqrtf(1 - SQR(cosNO));
        if (sinNI > 1.0f)
            return 1;
        const float cosNI = -sqrtf(1 - SQR(sinNI));
        const float je = (1 + inveta * (cosNO / cosNI)), jx = (1 + eta * (cosNI / cosNO));
        const float refr_variance = SQR(je) + SQR(jx), refl_variance = SQR(2.0f);
        return std::min(sqrtf((refr_variance / refl_variance) / cosNO), 1 / roughness);
    }
    float sample_weight(V3 wo, float ru, float rv, float rw) const
    {
        const V3 m = d.sample(wo, ru, rv);
        if (dot(wo, m) <= 0)
            return 0.0f;
        // attenuation(): thickness 0 -> A = 1
        const V3 wr       = refract(wo, m, eta);
        const float cosNO = std::min(dot(wo, m), 1.0f);
        const float cosNR = CLAMP(-wr.z, 0.0f, 1.0f);
        const float Rout = F(cosNO), Tin = 1.0f - Rout;
        const float Rin  = LERP(roughness, fresnel_dielectric(cosNR, 1 / eta), avg_invf());
        const float A = 1, b = SQR(Rin * A);
        const float R = Rout + (1 - b < 1e-4f ? (A < 1 ? 0 : Tin * 0.5f) : Tin * (1 - Rin) * SQR(A) * Rin / (1 - b));
        const float T = 1 - b < 1e-4f ? (A < 1 ? 0 : Tin * 0.5f) : Tin * (1 - Rin) * A / (1 - b);
        const float Fp = R / std::max(R + T, FLT_MIN);
        auto prob = [](float f) { return LERP(0.0f, f, CLAMP(f, 0.2f, 0.8f)); };
        const bool isrefl = rw < prob(Fp);
        const float sc    = slope_scale(wo.z);
        const V3 mt       = isrefl ? m : normalized({ m.x * sc, m.y * sc, m.z });
        if (dot(wo, mt) <= 0)
            return 0.0f;
        const V3 wif = reflect(wo, mt);
        const V3 wi  = isrefl ? wif : V3 { wif.x, wif.y, -wif.z };
        if ((isrefl && wi.z <= 0) || (!isrefl && wi.z >= 0))
            return 0.0f;
        const float P = prob(isrefl ? Fp : 1 - Fp);
        if (P < 1e-6f)
            return 0.0f;
        const float out = d.G2_G1({ wi.x, wi.y, std::fabs(wi.z) }, wo) / P;
        return (isrefl ? R : T) * out;
    }
};

// ---- genluts.cpp: the stratified scrambled sample set and the running mean -------------------
inline uint32_t ri_LP(uint32_t i)
{
    uint32_t r = 0;
    for (uint32_t v = 1U << 31; i; i >>= 1, v |= v >> 1)
        if (i & 1)
            r ^= v;
    return r;
}
inline uint32_t ri_LP_inv(uint32_t i)
{
    uint32_t r = 0;
    for (uint32_t v = 3U << 30; i; i >>= 1, v >>= 1)
        if (i & 1)
            r ^= v;
    return r;
}
inline void get_sample(int si, int AA, uint32_t sx, uint32_t sy, uint32_t sz, float& vx, float& vy, float& vz)
{
    const uint32_t ex = si % AA, ey = si / AA;
    const uint32_t upper   = (ex ^ (sx >> 16)) << 16;
    const uint32_t lpUpper = ri_LP(upper) ^ sy;
    const uint32_t delta   = (ey << 16) ^ (lpUpper & 0xFFFF0000u);
    const uint32_t lower   = ri_LP_inv(delta);
    const uint32_t index   = upper | lower;
    const uint32_t x       = index ^ sx;
    const uint32_t y       = lpUpper ^ delta;
    const float jx = (x & 65535) * (1 / 65536.0f), jy = (y & 65535) * (1 / 65536.0f);
    uint32_t rz = sz, ii = index;
    for (uint64_t v2 = uint64_t(3) << 62; ii; ii >>= 1, v2 ^= v2 >> 1)
        if (ii & 1)
            rz ^= uint32_t(v2 >> 31);
    vx = (ex + jx) / AA;
    vy = (ey + jy) / AA;
    vz = rz * 2.3283063e-10f;
}
inline uint64_t fasthash64_mix(uint64_t h)
{
    h ^= h >> 23;
    h *= 0x2127599bf4325c37ULL;
    h ^= h >> 47;
    return h;
}
inline uint32_t randhash3(uint32_t x, uint32_t y, uint32_t z)
{
    const uint64_t m = 0x880355f21e6d1965ULL;
    const uint64_t buf[2] = { (uint64_t(x) << 32) + y, uint64_t(z) };
    uint64_t h = (2 * sizeof(uint64_t)) * m;
    for (uint64_t v : buf) {
        h ^= fasthash64_mix(v);
        h *= m;
    }
    return (uint32_t)fasthash64_mix(h);
}
template<class BSDF> float compute_E(float cos_theta, const BSDF& bsdf, uint32_t fi, uint32_t ri)
{
    const int AA = 128, NUM_SAMPLES = AA * AA;
    const V3 wo  = { sqrtf(1 - SQR(cos_theta)), 0, cos_theta };
    float E      = 0;
    const uint32_t seedx = randhash3(fi, ri, 0), seedy = randhash3(fi, ri, 1), seedz = randhash3(fi, ri, 2);
    for (int i = 0; i < NUM_SAMPLES; i++) {
        float rx, ry, rz;
        get_sample(i, AA, seedx, seedy, seedz, rx, ry, rz);
        float out = bsdf.sample_weight(wo, rx, ry, rz);
        E         = LERP(1.0f / (1.0f + i), E, out);
    }
    return std::min(E, 1.0f);
}
template<class BSDF> void bake(float* stored, int Nf, bool sqr_cosines, const float* ggx_table)
{
    const int Nr = 16, Nc = 16;
    auto get_cosine = [&](int i) {
        const float x = float(i) * (1.0f / (Nc - 1));
        return std::max(sqr_cosines ? SQR(x) : x, 1e-6f);
    };
    std::vector<std::thread> th;
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([&, t] {
            for (int f = (int)t; f < Nf; f += (int)nt) {
                const float fresnel_index = float(f) * (1.0f / std::max(1, Nf - 1));
                for (int r = 0; r < Nr; r++) {
                    const float roughness_index = float(r) * (1.0f / (Nr - 1));
                    for (int c = 0; c < Nc; c++) {
                        const BSDF bsdf(get_cosine(c), roughness_index, fresnel_index, ggx_table);
                        stored[f * Nr * Nc + r * Nc + c] = 1 - compute_E(get_cosine(c), bsdf, (uint32_t)f, (uint32_t)r);
                    }
                }
            }
        });
    for (auto& x : th)
        x.join();
}

}  // namespace

int
main(int argc, char** argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s bsdl_luts.bin [thinlayer_lut.bin]\n", argv[0]);
        return 2;
    }
    std::vector<float> luts(256 + 3 * 8192, 0.0f);
    bake<MiniMicrofacetGGX>(luts.data(), 1, true, nullptr);
    bake<Dielectric<false, false>>(luts.data() + 256, 32, false, luts.data());
    bake<Dielectric<true, false>>(luts.data() + 256 + 8192, 32, false, luts.data());
    bake<Dielectric<true, true>>(luts.data() + 256 + 2 * 8192, 32, false, luts.data());
    FILE* f = fopen(argv[1], "wb");
    if (!f || fwrite(luts.data(), sizeof(float), luts.size(), f) != luts.size()) {
        fprintf(stderr, "cannot write %s\n", argv[1]);
        return 1;
    }
    fclose(f);
    printf("wrote %zu floats to %s\n", luts.size(), argv[1]);
    if (argc > 2) {   // spi::Thinlayer 32 x 16 x 16 (SPI/bsdf_thinlayer_decl.h:37-41), its own file
        std::vector<float> thin(8192, 0.0f);
        bake<ThinlayerBake>(thin.data(), 32, false, nullptr);
        f = fopen(argv[2], "wb");
        if (!f || fwrite(thin.data(), sizeof(float), thin.size(), f) != thin.size()) {
            fprintf(stderr, "cannot write %s\n", argv[2]);
            return 1;
        }
        fclose(f);
        printf("wrote %zu floats to %s\n", thin.size(), argv[2]);
    }
    return 0;
}
