// oracle_bsdl_check.cpp — CPU ORACLE (test infrastructure, NOT product code).
// The restated libbsdl lobes of osl_oracle_lobes.h / osl_oracle_mxlobes.h behind the same C
// signature as oracle/ref_bsdl.cpp (the reference's own classes), so that
// tests/test_oracle_bsdl.py can compare the two value by value.
#include "osl_oracle_render.h"

using namespace oslo;

extern "C" void oracle_set_bsdl_luts(const float* p) { bsdl_luts() = p; }

// lobe: 0 conductor  1 dielectric  2 generalized schlick  3 translucent  4 sheen  5 oren-nayar diffuse
//       6 burley diffuse  7 spi thinlayer;  p: closure parameters in registration order, strings skipped (see ref_bsdl.cpp)
// mode 0: eval(wo, wi = arg)   1: sample(wo, rnd = arg)   2: albedo   3: filter_o
extern "C" int
oracle_bsdl(int lobe, const float* p, const float* wo_, int backfacing, float path_roughness, int mode, const float* arg,
            float* out)
{
    // a closure component laid out the way the generated code does (one word for the distribution string)
    float store[64] = { 0 };
    ClosComp* comp  = reinterpret_cast<ClosComp*>(store);
    comp->w         = V3(1.0f);
    float* q        = comp->params;
    SG sg;
    std::memset((void*)&sg, 0, sizeof sg);
    const V3 wo(wo_[0], wo_[1], wo_[2]);
    sg.I.val      = -wo;
    sg.backfacing = backfacing;
    Lobe l;
    bool schlick = false;
    switch (lobe) {
    case 0:
        comp->id = MX_CONDUCTOR_ID;
        std::memcpy(q, p, 14 * 4);
        mx_from_component(l, comp, sg, path_roughness);
        break;
    case 1:
        comp->id = MX_DIELECTRIC_ID;
        std::memcpy(q, p, 15 * 4);          // N U refl refr rx ry ior
        q[15] = 0;                          // distribution
        q[16] = p[15]; q[17] = p[16];       // thinfilm
        q[18] = p[17]; q[19] = p[18]; q[20] = p[19];   // absorption
        q[21] = p[20];                      // dispersion
        mx_from_component(l, comp, sg, path_roughness);
        break;
    case 2:
        comp->id = MX_GENERALIZED_SCHLICK_ID;
        std::memcpy(q, p, 21 * 4);
        mx_from_component(l, comp, sg, path_roughness);
        schlick = true;
        break;
    case 3:
        l.type   = LOBE_MX_TRANSLUCENT;
        l.N      = V3(p[0], p[1], p[2]);
        l.albedo = V3(p[3], p[4], p[5]);
        l.tf     = TangentFrame::from_normal(lobes::bsdl_visible_normal(wo, l.N, l.N));
        break;
    case 4:
        comp->id = MX_SHEEN_ID;
        std::memcpy(q, p, 7 * 4);
        { int m = (int)p[7]; std::memcpy(q + 7, &m, 4); }
        if (!sheen_from_component(l, comp, sg, path_roughness))
            return 3;
        break;
    case 5:
    case 6:
        l.type   = lobe == 6 ? LOBE_BSDL_BURLEY : LOBE_BSDL_OREN_NAYAR;
        l.N      = V3(p[0], p[1], p[2]);
        l.albedo = V3(p[3], p[4], p[5]);
        l.ax     = lobes::bsdl_clamp(p[6], 0.0f, 1.0f);
        l.energy_compensation = lobe == 5 ? (int)p[7] : 0;
        l.tf     = TangentFrame::from_normal(lobes::bsdl_visible_normal(wo, l.N, l.N));
        break;
    case 7: {   // thinlayer: N T IOR roughness anisotropy thickness refl_tint refr_tint sigma_t
        l.type     = LOBE_SPI_THINLAYER;
        l.N        = V3(p[0], p[1], p[2]);
        const V3 Z = lobes::bsdl_visible_normal(wo, l.N, l.N);
        l.tf       = lobes::bsdl_frame_zx(Z, V3(p[3], p[4], p[5]));
        l.thin     = lobes::thin_setup(dot(wo, Z), p[6], p[7], p[8], p[9], V3(p[10], p[11], p[12]), V3(p[13], p[14], p[15]),
                                       V3(p[16], p[17], p[18]), path_roughness);
        break;
    }
    default: return 1;
    }
    BSample s;
    if (mode == 0)
        s = l.eval(wo, V3(arg[0], arg[1], arg[2]));
    else if (mode == 1)
        s = l.sample(wo, arg[0], arg[1], arg[2]);
    else if (mode == 2) {
        V3 a = l.get_albedo(wo);
        out[0] = a.x; out[1] = a.y; out[2] = a.z;
        return 0;
    } else if (mode == 3) {
        V3 a;
        if (l.type == LOBE_MX_SPEC && !l.mx.conductor)
            a = mx_filter_o(l.mx, schlick);
        else if (l.type == LOBE_BSDL_SHEEN)
            a = V3(l.emiss);
        else
            return 2;
        out[0] = a.x; out[1] = a.y; out[2] = a.z;
        return 0;
    } else
        return 1;
    out[0] = s.wi.x; out[1] = s.wi.y; out[2] = s.wi.z;
    out[3] = s.weight.x; out[4] = s.weight.y; out[5] = s.weight.z;
    out[6] = s.pdf; out[7] = s.roughness;
    return 0;
}
