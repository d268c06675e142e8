// ref_bsdl.cpp — C-ABI window onto the REFERENCE's own libbsdl lobes (TEST INFRASTRUCTURE).
//
// oracle/build_ref.py compiles this file against the reference's headers where they lie
// (/root/reference/src/libbsdl/include), the Imath stand-in in oracle/ref_shim and the LUT
// headers produced by the reference's own genluts (built and run by the same script), into
// oracle/_ref/libref_bsdl.so.  Nothing of the reference is copied into the repo.  The tests
// use it to check the restated lobes of oracle/osl_oracle_lobes.h (and the LUTs baked by
// tools/bake_bsdl_luts.py) value by value: eval / sample / albedo / filter_o on random inputs.
//
// The lobes are instantiated the way testrender does (BSDL_WRAP, src/testrender/shading.cpp:
// 73-115): BsdfGlobals(wo, N, N, backfacing, path_roughness, outer_ior 1, lambda_0 0), three
// RGB channels, and Fast:: math = OIIO fast_* (src/testrender/bsdl_config.h) - here the
// oracle's restatement of those polynomials, so that a difference can only come from the lobe
// arithmetic itself.
#include <cstdint>
#include <cstring>

#include "osl_oracle_ops.h"   // oslo::fast_* (restated OIIO fmath)

#define BSDL_INLINE        static inline
#define BSDL_INLINE_METHOD inline
#define BSDL_DECL
#define BSDL_UNROLL()
#define BSDL_STRHASH(str) ((uintptr_t)0)

#include <BSDL/config.h>

struct BSDLConfig : public bsdl::BSDLDefaultConfig {
    static constexpr int HERO_WAVELENGTH_CHANNELS = 3;
    struct Fast {
        static float cosf(float x) { return oslo::fast_cos(x); }
        static float sinf(float x) { return oslo::fast_sin(x); }
        static float asinf(float x) { return oslo::fast_asin(x); }
        static float acosf(float x) { return oslo::fast_acos(x); }
        static float atan2f(float y, float x) { return oslo::fast_atan2(y, x); }
        static void sincosf(float x, float* s, float* c) { oslo::fast_sincos(x, s, c); }
        // (sinpi / cospi / log1p are only used by the hair and volume lobes, which are not checked here)
        static float sinpif(float x) { return std::sin(x * float(M_PI)); }
        static float cospif(float x) { return std::cos(x * float(M_PI)); }
        static float expf(float x) { return oslo::fast_exp(x); }
        static float exp2f(float x) { return oslo::fast_exp2(x); }
        static float logf(float x) { return oslo::fast_log(x); }
        static float log2f(float x) { return oslo::fast_log2(x); }
        static float log1pf(float x) { return std::log1p(x); }
        static float powf(float x, float y) { return oslo::fast_safe_pow(x, y); }
    };
    static ColorSpaceTag current_color_space() { return ColorSpaceTag::sRGB; }
    static const JakobHanikaLut* get_jakobhanika_lut(ColorSpaceTag) { return nullptr; }
};

#include <BSDL/MTX/bsdf_burley_diffuse_impl.h>
#include <BSDL/MTX/bsdf_conductor_impl.h>
#include <BSDL/MTX/bsdf_dielectric_impl.h>
#include <BSDL/MTX/bsdf_oren_nayar_diffuse_impl.h>
#include <BSDL/MTX/bsdf_schlick_impl.h>
#include <BSDL/MTX/bsdf_sheen_impl.h>
#include <BSDL/MTX/bsdf_translucent_impl.h>
#include <BSDL/SPI/bsdf_thinlayer_impl.h>
#include <BSDL/spectrum_impl.h>

namespace {

using Imath::C3f;
using Imath::V3f;

// what BSDL asks of its root class (testrender's BSDLLobe, shading.cpp:52-68)
struct Root {
    template<typename L> Root(L*, float rough, float, bool) : m_roughness(rough) {}
    void set_roughness(float r) { m_roughness = r; }
    float roughness() const { return m_roughness; }
    float m_roughness;
};

V3f v3(const float* p) { return V3f(p[0], p[1], p[2]); }
C3f c3(const float* p) { return C3f(p[0], p[1], p[2]); }

template<class LOBE> struct Wrap : public LOBE {
    using Data = typename LOBE::Data;
    Wrap(const Data& d, const V3f& wo, bool backfacing, float path_roughness)
        : LOBE(this, bsdl::BsdfGlobals(wo, d.N, d.N, backfacing, path_roughness, 1.0f, 0), d)
    {
    }
};

void put(float* out, const bsdl::Sample& s, const V3f& wi)
{
    C3f w = s.weight.toRGB(0);
    out[0] = wi.x; out[1] = wi.y; out[2] = wi.z;
    out[3] = w.x; out[4] = w.y; out[5] = w.z;
    out[6] = s.pdf; out[7] = s.roughness;
}

// mode 0: eval(wo, wi = arg)   1: sample(wo, rnd = arg)   2: albedo   3: filter_o (layering)
template<class W, bool HAS_FILTER_O>
int run(const typename W::Data& d, const float* wo_, int backfacing, float path_roughness, int mode, const float* arg,
        float* out)
{
    const V3f wo = v3(wo_);
    W lobe(d, wo, backfacing != 0, path_roughness);
    if (mode == 0) {
        const V3f wi = v3(arg);
        put(out, lobe.eval_impl(lobe.frame.local(wo), lobe.frame.local(wi)), wi);
    } else if (mode == 1) {
        bsdl::Sample s = lobe.sample_impl(lobe.frame.local(wo), v3(arg));
        put(out, s, lobe.frame.world(s.wi));
    } else if (mode == 2) {
        C3f a = lobe.albedo_impl().toRGB(0);
        out[0] = a.x; out[1] = a.y; out[2] = a.z;
    } else if (mode == 3) {
        if constexpr (HAS_FILTER_O) {
            C3f a = lobe.filter_o(wo).toRGB(0);
            out[0] = a.x; out[1] = a.y; out[2] = a.z;
        } else
            return 2;
    } else
        return 1;
    return 0;
}

}  // namespace

// lobe: 0 conductor  1 dielectric  2 generalized schlick  3 translucent  4 sheen  5 oren-nayar diffuse
//       6 burley diffuse  7 spi thinlayer
// p: the closure's parameters in registration order (strings skipped), see the Data structs
extern "C" int
ref_bsdl(int lobe, const float* p, const float* wo, int backfacing, float path_roughness, int mode,
         const float* arg, float* out)
{
    using namespace bsdl;
    switch (lobe) {
    case 0: {
        using W = Wrap<mtx::ConductorLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.U = v3(p + 3); d.roughness_x = p[6]; d.roughness_y = p[7];
        d.IOR = c3(p + 8); d.extinction = c3(p + 11);
        return run<W, false>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 1: {
        using W = Wrap<mtx::DielectricLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.U = v3(p + 3); d.refl_tint = c3(p + 6); d.refr_tint = c3(p + 9);
        d.roughness_x = p[12]; d.roughness_y = p[13]; d.IOR = p[14];
        d.thinfilm_thickness = p[15]; d.thinfilm_ior = p[16]; d.absorption = c3(p + 17); d.dispersion = p[20];
        return run<W, true>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 2: {
        using W = Wrap<mtx::SchlickLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.U = v3(p + 3); d.refl_tint = c3(p + 6); d.refr_tint = c3(p + 9);
        d.roughness_x = p[12]; d.roughness_y = p[13]; d.F0 = c3(p + 14); d.F90 = c3(p + 17); d.exponent = p[20];
        return run<W, true>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 3: {
        using W = Wrap<mtx::TranslucentLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.albedo = c3(p + 3);
        return run<W, false>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 4: {
        using W = Wrap<mtx::SheenLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.albedo = c3(p + 3); d.roughness = p[6]; d.mode = (int)p[7];
        return run<W, true>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 5: {
        using W = Wrap<mtx::OrenNayarDiffuseLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.albedo = c3(p + 3); d.roughness = p[6]; d.energy_compensation = (int)p[7];
        return run<W, false>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 6: {
        using W = Wrap<mtx::BurleyDiffuseLobe<Root>>;
        W::Data d {};
        d.N = v3(p); d.albedo = c3(p + 3); d.roughness = p[6];
        return run<W, false>(d, wo, backfacing, path_roughness, mode, arg, out);
    }
    case 7: {   // SpiThinLayer (shading.cpp:119-152): backfacing false, both lobes on, no albedo_impl / filter_o
        using L = spi::ThinLayerLobe<Root>;
        struct W : public L {
            W(const L::Data& d, const V3f& wo, float path_roughness)
                : L(this, bsdl::BsdfGlobals(wo, d.N, d.N, false, path_roughness, 1.0f, 0), d)
            {
            }
        };
        L::Data d {};
        d.N = v3(p); d.T = v3(p + 3); d.IOR = p[6]; d.roughness = p[7]; d.anisotropy = p[8]; d.thickness = p[9];
        d.prob_clamp = 0;
        d.refl_tint = c3(p + 10); d.refr_tint = c3(p + 13); d.sigma_t = c3(p + 16);
        const V3f wov = v3(wo);
        W lobe(d, wov, path_roughness);
        if (mode == 0) {
            const V3f wi = v3(arg);
            put(out, lobe.eval_impl(lobe.frame.local(wov), lobe.frame.local(wi), true, true), wi);
        } else if (mode == 1) {
            bsdl::Sample s = lobe.sample_impl(lobe.frame.local(wov), v3(arg), true, true);
            put(out, s, lobe.frame.world(s.wi));
        } else if (mode == 2) {
            out[0] = out[1] = out[2] = 1.0f;   // BSDF::get_albedo default (shading.h:287)
        } else
            return 2;
        return 0;
    }
    default: return 1;
    }
}

// raw LUT access: table 0 MiniMicrofacetGGX, 1 DielectricReflFront, 2 DielectricBothFront,
// 3 DielectricBothBack, 4 ZeltnerBurleySheen, 5 ContyKullaSheen, 6 Thinlayer, 7 Zeltner LTC coefficients
extern "C" const float*
ref_bsdl_lut(int table, int* count)
{
    using namespace bsdl;
#define T(n, type) case n: *count = (int)(sizeof(type::get_energy().data) / sizeof(float)); return type::get_energy().data;
    switch (table) {
        T(0, spi::MiniMicrofacetGGX)
        T(1, mtx::DielectricReflFront)
        T(2, mtx::DielectricBothFront)
        T(3, mtx::DielectricBothBack)
        T(4, mtx::ZeltnerBurleySheen)
        T(5, mtx::ContyKullaSheen)
        T(6, spi::Thinlayer)
    }
#undef T
    if (table == 7) {   // the 32x32 (A, B, R) coefficients of the Zeltner-Burley sheen LTC
        *count = 32 * 32 * 3;
        return &bsdl::mtx::ZeltnerBurleySheen::param_ptr()[0][0].x;
    }
    *count = 0;
    return nullptr;
}
