// osl_oracle_ops.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Scalar restatement of the osl_* shadeop runtime that generated layer code
// calls (reference: src/liboslexec/llvm_ops.cpp, src/include/OSL/dual.h,
// dual_vec.h, and the per-op IR emitters in src/liboslexec/llvm_gen.cpp).
// Everything per-component is written once over S in {float, Df}; the layer
// code produced by oracle/oso2cpp.py pulls components with getc()/setc()
// exactly the way llvm_gen loads "component i, derivative d" of an operand
// (llvm_gen.cpp llvm_load_value(sym, deriv, component, cast)).
//
// OIIO fast_* forms are restated from OpenImageIO's published fmath.h
// (OIIO >= 3.0; not in /root/reference): value-level PARITY UNPINNED.
#pragma once
#include "osl_oracle.h"
#include "osl_oracle_simplex.h"

namespace oslo {

// ---------------------------------------------------------------------------
// OIIO fmath.h restatements (madd is a*b+c, unfused in the default JIT mode)
// ---------------------------------------------------------------------------
inline float madd(float a, float b, float c) { return a * b + c; }
inline int fast_rint(float x) { return (int)std::rint(x); }
inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

inline void fast_sincos(float x, float* sine, float* cosine)
{
    int q    = fast_rint(x * float(M_1_PI));
    float qf = float(q);
    x        = madd(qf, -0.78515625f * 4, x);
    x        = madd(qf, -0.00024187564849853515625f * 4, x);
    x        = madd(qf, -3.7747668102383613586e-08f * 4, x);
    x        = madd(qf, -1.2816720341285448015e-12f * 4, x);
    x        = float(M_PI_2) - (float(M_PI_2) - x);
    float s  = x * x;
    if ((q & 1) != 0)
        x = -x;
    float su = 2.6083159809786593541503e-06f;
    su       = madd(su, s, -0.0001981069071916863322258f);
    su       = madd(su, s, +0.00833307858556509017944336f);
    su       = madd(su, s, -0.166666597127914428710938f);
    su       = madd(s, su * x, x);
    float cu = -2.71811842367242206819355e-07f;
    cu       = madd(cu, s, +2.47990446951007470488548e-05f);
    cu       = madd(cu, s, -0.00138888787478208541870117f);
    cu       = madd(cu, s, +0.0416666641831398010253906f);
    cu       = madd(cu, s, -0.5f);
    cu       = madd(cu, s, +1.0f);
    if ((q & 1) != 0)
        cu = -cu;
    if (std::fabs(su) > 1.0f)
        su = 0.0f;
    if (std::fabs(cu) > 1.0f)
        cu = 0.0f;
    *sine   = su;
    *cosine = cu;
}
inline float fast_sin(float x)
{
    float s, c;
    fast_sincos(x, &s, &c);
    return s;
}
inline float fast_cos(float x)
{
    float s, c;
    fast_sincos(x, &s, &c);
    return c;
}
inline float fast_tan(float x)
{
    int q    = fast_rint(x * float(2 * M_1_PI));
    float qf = float(q);
    x        = madd(qf, -0.78515625f * 2, x);
    x        = madd(qf, -0.00024187564849853515625f * 2, x);
    x        = madd(qf, -3.7747668102383613586e-08f * 2, x);
    x        = madd(qf, -1.2816720341285448015e-12f * 2, x);
    if ((q & 1) == 0)
        x = float(M_PI_4) - (float(M_PI_4) - x);
    float s = x * x;
    float u = 0.00927245803177356719970703f;
    u       = madd(u, s, 0.00331984995864331722259521f);
    u       = madd(u, s, 0.0242998078465461730957031f);
    u       = madd(u, s, 0.0534495301544666290283203f);
    u       = madd(u, s, 0.133383005857467651367188f);
    u       = madd(u, s, 0.333331853151321411132812f);
    u       = madd(s, u * x, x);
    if ((q & 1) != 0)
        u = -1.0f / u;
    return u;
}
inline float fast_acos(float x)
{
    const float f = std::fabs(x);
    const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
    const float a = std::sqrt(1.0f - m)
                    * (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
    return x < 0 ? float(M_PI) - a : a;
}
inline float fast_asin(float x)
{
    const float f = std::fabs(x);
    const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
    const float a = float(M_PI_2)
                    - std::sqrt(1.0f - m)
                          * (1.5707963267f
                             + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
    return std::copysign(a, x);
}
inline float fast_atan(float x)
{
    const float a = std::fabs(x);
    const float k = a > 1.0f ? 1 / a : a;
    const float s = 1.0f - (1.0f - k);
    const float t = s * s;
    float r = s * madd(0.43157974f, t, 1.0f) / madd(madd(0.05831938f, t, 0.76443945f), t, 1.0f);
    if (a > 1.0f)
        r = 1.570796326794896557998982f - r;
    return std::copysign(r, x);
}
inline float fast_atan2(float y, float x)
{
    const float a = std::fabs(x);
    const float b = std::fabs(y);
    const float k = (b == 0) ? 0.0f : ((a == b) ? 1.0f : (b > a ? a / b : b / a));
    const float s = 1.0f - (1.0f - k);
    const float t = s * s;
    float r = s * madd(0.43157974f, t, 1.0f) / madd(madd(0.05831938f, t, 0.76443945f), t, 1.0f);
    if (b > a)
        r = 1.570796326794896557998982f - r;
    if (f2u(x) & 0x80000000u)
        r = float(M_PI) - r;
    return std::copysign(r, y);
}
inline float fast_log2(float x)
{
    x = clampf(x, std::numeric_limits<float>::min(), std::numeric_limits<float>::max());
    unsigned bits = f2u(x);
    int exponent  = int(bits >> 23) - 127;
    float f       = u2f((bits & 0x007FFFFF) | 0x3f800000) - 1.0f;
    float f2      = f * f;
    float f4      = f2 * f2;
    float hi      = madd(f, -0.00931049621349f, 0.05206469089414f);
    float lo      = madd(f, 0.47868480909345f, -0.72116591947498f);
    hi            = madd(f, hi, -0.13753123777116f);
    hi            = madd(f, hi, 0.24187369696082f);
    hi            = madd(f, hi, -0.34730547155299f);
    lo            = madd(f, lo, 1.442689881667200f);
    return ((f4 * hi) + (f * lo)) + exponent;
}
inline float fast_log(float x) { return fast_log2(x) * float(M_LN2); }
inline float fast_log10(float x) { return fast_log2(x) * float(M_LN2 / M_LN10); }
inline float fast_logb(float x)
{
    x = std::fabs(x);
    if (x < std::numeric_limits<float>::min())
        x = std::numeric_limits<float>::min();
    if (x > std::numeric_limits<float>::max())
        x = std::numeric_limits<float>::max();
    return float(int(f2u(x) >> 23) - 127);
}
inline float fast_exp2(float x)
{
    if (x < -126.0f)
        x = -126.0f;
    if (x > 126.0f)
        x = 126.0f;
    int m = int(x);
    x -= m;
    x       = 1.0f - (1.0f - x);
    float r = 1.33336498402e-3f;
    r       = madd(x, r, 9.810352697968e-3f);
    r       = madd(x, r, 5.551834031939e-2f);
    r       = madd(x, r, 0.2401793301105f);
    r       = madd(x, r, 0.693144857883f);
    r       = madd(x, r, 1.0f);
    return u2f(f2u(r) + ((unsigned)m << 23));
}
inline float fast_exp(float x) { return fast_exp2(x * float(1 / M_LN2)); }
inline float fast_expm1(float x)
{
    if (std::fabs(x) < 0.03f) {
        float y = 1.0f - (1.0f - x);
        return std::copysign(madd(0.5f, y * y, y), x);
    }
    return fast_exp(x) - 1.0f;
}
inline float fast_sinh(float x)
{
    float a = std::fabs(x);
    if (a > 1.0f) {
        float e = fast_exp(a);
        return std::copysign(0.5f * e - 0.5f / e, x);
    }
    a        = 1.0f - (1.0f - a);
    float a2 = a * a;
    float r  = 2.03945513931e-4f;
    r        = madd(r, a2, 8.32990277558e-3f);
    r        = madd(r, a2, 0.1666673421859f);
    r        = madd(r * a, a2, a);
    return std::copysign(r, x);
}
inline float fast_cosh(float x)
{
    float e = fast_exp(std::fabs(x));
    return 0.5f * e + 0.5f / e;
}
inline float fast_tanh(float x)
{
    float e = fast_exp(2.0f * std::fabs(x));
    return std::copysign(1 - 2 / (1 + e), x);
}
inline float fast_safe_pow(float x, float y)
{
    if (y == 0)
        return 1.0f;
    if (x == 0)
        return 0.0f;
    if (y == 1.0f)
        return x;
    if (y == 2.0f)
        return std::min(x * x, std::numeric_limits<float>::max());
    float sign = 1.0f;
    if (x < 0) {
        int ybits = (int)f2u(y) & 0x7fffffff;
        if (ybits >= 0x4b800000) {
            // always an even int
        } else if (ybits >= 0x3f800000) {
            int k = (ybits >> 23) - 127;
            int j = ybits >> (23 - k);
            if ((j << (23 - k)) == ybits)
                sign = u2f(0x3f800000u | ((unsigned)j << 31));
            else
                return 0.0f;
        } else {
            return 0.0f;
        }
    }
    return sign * fast_exp2(y * fast_log2(std::fabs(x)));
}
inline float fast_erf(float x)
{
    const float a1 = 0.0705230784f, a2 = 0.0422820123f, a3 = 0.0092705272f,
                a4 = 0.0001520143f, a5 = 0.0002765672f, a6 = 0.0000430638f;
    const float a = std::fabs(x);
    const float b = 1.0f - (1.0f - a);
    const float r = madd(madd(madd(madd(madd(madd(a6, b, a5), b, a4), b, a3), b, a2), b, a1), b, 1.0f);
    const float s = r * r;
    const float t = s * s;
    const float u = t * t;
    const float v = u * u;
    return std::copysign(1.0f - 1.0f / v, x);
}
inline float fast_erfc(float x) { return 1.0f - fast_erf(x); }
inline float fast_ierf(float x)
{
    float a = std::fabs(x);
    if (a > 0.99999994f)
        a = 0.99999994f;
    float w = -fast_log((1.0f - a) * (1.0f + a)), p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = madd(p, w, 3.43273939e-07f);
        p = madd(p, w, -3.5233877e-06f);
        p = madd(p, w, -4.39150654e-06f);
        p = madd(p, w, 0.00021858087f);
        p = madd(p, w, -0.00125372503f);
        p = madd(p, w, -0.00417768164f);
        p = madd(p, w, 0.246640727f);
        p = madd(p, w, 1.50140941f);
    } else {
        w = std::sqrt(w) - 3.0f;
        p = -0.000200214257f;
        p = madd(p, w, 0.000100950558f);
        p = madd(p, w, 0.00134934322f);
        p = madd(p, w, -0.00367342844f);
        p = madd(p, w, 0.00573950773f);
        p = madd(p, w, -0.0076224613f);
        p = madd(p, w, 0.00943887047f);
        p = madd(p, w, 1.00167406f);
        p = madd(p, w, 2.83297682f);
    }
    return p * x;
}
inline float fast_cbrt(float x)
{
    float x0 = std::fabs(x);
    float a  = u2f(0x2a5137a0 + f2u(x0) / 3);
    a        = 0.333333333f * (2.0f * a + x0 / (a * a));
    a        = 0.333333333f * (2.0f * a + x0 / (a * a));
    a        = (x0 == 0) ? 0 : a;
    return std::copysign(a, x);
}
inline float safe_sqrt(float x) { return x >= 0.0f ? std::sqrt(x) : 0.0f; }
inline float safe_inversesqrt(float x) { return x > 0.0f ? 1.0f / std::sqrt(x) : 0.0f; }
inline float safe_fmod(float a, float b)
{
    if (b != 0.0f) {
        int N = (int)(a / b);
        return a - N * b;
    }
    return 0.0f;
}
// llvm_ops.cpp:684-700
inline float safe_div(float a, float b)
{
    float q = a / b;
    return std::isfinite(q) ? q : 0.0f;
}

// ---------------------------------------------------------------------------
// per-component scalar ops, S in {float, Df}.  A float argument converts to
// Df implicitly (zero derivs), mirroring the *_dffdf / *_dfdff entry points.
// ---------------------------------------------------------------------------
inline float nd(float a) { return a; }
inline float nd(const Df& a) { return a.val; }
inline V3 nd(const V3& a) { return a; }
inline V3 nd(const Dv& a) { return a.val; }
inline int nd(int a) { return a; }

inline float o_add(float a, float b) { return a + b; }
inline Df o_add(const Df& a, const Df& b) { return a + b; }
inline int o_add(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
inline float o_sub(float a, float b) { return a - b; }
inline Df o_sub(const Df& a, const Df& b) { return a - b; }
inline int o_sub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
inline float o_mul(float a, float b) { return a * b; }
inline Df o_mul(const Df& a, const Df& b) { return a * b; }
inline int o_mul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
// llvm_gen_div (llvm_gen.cpp): safe_div unless the divisor is a nonzero const
inline float o_div(float a, float b) { return safe_div(a, b); }
inline Df o_div(const Df& a, const Df& b)
{
    float q    = safe_div(a.val, b.val);
    float binv = safe_div(1.0f, b.val);
    return Df(q, binv * (a.dx - q * b.dx), binv * (a.dy - q * b.dy));
}
inline int o_div(int a, int b) { return b != 0 ? a / b : 0; }
inline float o_divc(float a, float b) { return a / b; }  // divisor = nonzero constant
inline Df o_divc(const Df& a, const Df& b)
{
    float q    = a.val / b.val;
    float binv = 1.0f / b.val;
    return Df(q, binv * (a.dx - q * b.dx), binv * (a.dy - q * b.dy));
}
inline int o_divc(int a, int b) { return a / b; }
inline int o_mod(int a, int b) { return b != 0 ? a % b : 0; }
inline float o_neg(float a) { return -a; }
inline Df o_neg(const Df& a) { return -a; }
inline int o_neg(int a) { return -a; }

#define OSLO_UNARY(name, ffun, dfun)                  \
    inline float o_##name(float a) { return ffun(a); } \
    inline Df o_##name(const Df& a) { return dfun(a); }

inline Df d_sin(const Df& a) { float s, c; fast_sincos(a.val, &s, &c); return dualfunc(a, s, c); }
inline Df d_cos(const Df& a) { float s, c; fast_sincos(a.val, &s, &c); return dualfunc(a, c, -s); }
inline Df d_tan(const Df& a)
{
    float t = fast_tan(a.val), c = fast_cos(a.val);
    return dualfunc(a, t, 1 / (c * c));
}
inline Df d_asin(const Df& a)
{
    float f  = fast_asin(a.val);
    float df = std::fabs(a.val) < 1.0f ? 1.0f / std::sqrt(1.0f - a.val * a.val) : 0.0f;
    return dualfunc(a, f, df);
}
inline Df d_acos(const Df& a)
{
    float f  = fast_acos(a.val);
    float df = std::fabs(a.val) < 1.0f ? -1.0f / std::sqrt(1.0f - a.val * a.val) : 0.0f;
    return dualfunc(a, f, df);
}
inline Df d_atan(const Df& a) { return dualfunc(a, fast_atan(a.val), 1.0f / (1.0f + a.val * a.val)); }
inline Df d_sinh(const Df& a) { return dualfunc(a, fast_sinh(a.val), fast_cosh(a.val)); }
inline Df d_cosh(const Df& a) { return dualfunc(a, fast_cosh(a.val), fast_sinh(a.val)); }
inline Df d_tanh(const Df& a)
{
    float t = fast_tanh(a.val), c = fast_cosh(a.val);
    return dualfunc(a, t, 1.0f / (c * c));
}
inline Df d_log(const Df& a)
{
    float df = a.val < std::numeric_limits<float>::min() ? 0.0f : 1.0f / a.val;
    return dualfunc(a, fast_log(a.val), df);
}
inline Df d_log2(const Df& a)
{
    float aln2 = a.val * float(M_LN2);
    float df   = aln2 < std::numeric_limits<float>::min() ? 0.0f : 1.0f / aln2;
    return dualfunc(a, fast_log2(a.val), df);
}
inline Df d_log10(const Df& a)
{
    float al = a.val * float(M_LN10);
    float df = al < std::numeric_limits<float>::min() ? 0.0f : 1.0f / al;
    return dualfunc(a, fast_log10(a.val), df);
}
inline Df d_exp(const Df& a) { float f = fast_exp(a.val); return dualfunc(a, f, f); }
inline Df d_exp2(const Df& a) { float f = fast_exp2(a.val); return dualfunc(a, f, f * float(M_LN2)); }
inline Df d_expm1(const Df& a) { return dualfunc(a, fast_expm1(a.val), fast_exp(a.val)); }
inline Df d_erf(const Df& a)
{
    return dualfunc(a, fast_erf(a.val), fast_exp(-a.val * a.val) * 1.128379167095512573896158903f);
}
inline Df d_erfc(const Df& a)
{
    return dualfunc(a, fast_erfc(a.val), fast_exp(-a.val * a.val) * -1.128379167095512573896158903f);
}
inline Df d_cbrt(const Df& a)
{
    if (a.val != 0.0f) {
        float f = fast_cbrt(a.val);
        return dualfunc(a, f, 1.0f / (3.0f * f * f));
    }
    return Df(0.0f);
}
inline Df d_sqrt(const Df& a)
{
    if (a.val > 0.0f) {
        float f = std::sqrt(a.val);
        return dualfunc(a, f, 0.5f / f);
    }
    return Df(0.0f);
}
inline Df d_inversesqrt(const Df& a)
{
    if (a.val > 0.0f) {
        float f = 1.0f / std::sqrt(a.val);
        return dualfunc(a, f, -0.5f * f / a.val);
    }
    return Df(0.0f);
}
inline Df d_fabs(const Df& a) { return a.val >= 0 ? a : -a; }
inline float sign_f(float x) { return x < 0.0f ? -1.0f : (x == 0.0f ? 0.0f : 1.0f); }
// value-only functions: result carries no derivatives (llvm_gen zeroes them)
inline Df d_floor(const Df& a) { return Df(std::floor(a.val)); }
inline Df d_ceil(const Df& a) { return Df(std::ceil(a.val)); }
inline Df d_round(const Df& a) { return Df(std::round(a.val)); }
inline Df d_trunc(const Df& a) { return Df(std::trunc(a.val)); }
inline Df d_sign(const Df& a) { return Df(sign_f(a.val)); }
inline Df d_logb(const Df& a) { return Df(fast_logb(a.val)); }

OSLO_UNARY(sin, fast_sin, d_sin)
OSLO_UNARY(cos, fast_cos, d_cos)
OSLO_UNARY(tan, fast_tan, d_tan)
OSLO_UNARY(asin, fast_asin, d_asin)
OSLO_UNARY(acos, fast_acos, d_acos)
OSLO_UNARY(atan, fast_atan, d_atan)
OSLO_UNARY(sinh, fast_sinh, d_sinh)
OSLO_UNARY(cosh, fast_cosh, d_cosh)
OSLO_UNARY(tanh, fast_tanh, d_tanh)
OSLO_UNARY(log, fast_log, d_log)
OSLO_UNARY(log2, fast_log2, d_log2)
OSLO_UNARY(log10, fast_log10, d_log10)
OSLO_UNARY(exp, fast_exp, d_exp)
OSLO_UNARY(exp2, fast_exp2, d_exp2)
OSLO_UNARY(expm1, fast_expm1, d_expm1)
OSLO_UNARY(erf, fast_erf, d_erf)
OSLO_UNARY(erfc, fast_erfc, d_erfc)
OSLO_UNARY(cbrt, fast_cbrt, d_cbrt)
OSLO_UNARY(sqrt, safe_sqrt, d_sqrt)
OSLO_UNARY(inversesqrt, safe_inversesqrt, d_inversesqrt)
OSLO_UNARY(abs, std::fabs, d_fabs)
OSLO_UNARY(fabs, std::fabs, d_fabs)
OSLO_UNARY(floor, std::floor, d_floor)
OSLO_UNARY(ceil, std::ceil, d_ceil)
OSLO_UNARY(round, std::round, d_round)
OSLO_UNARY(trunc, std::trunc, d_trunc)
OSLO_UNARY(sign, sign_f, d_sign)
OSLO_UNARY(logb, fast_logb, d_logb)
inline int o_abs(int a) { return std::abs(a); }
inline int o_fabs(int a) { return std::abs(a); }

inline float o_atan2(float y, float x) { return fast_atan2(y, x); }
inline Df o_atan2(const Df& y, const Df& x)
{
    float f     = fast_atan2(y.val, x.val);
    float denom = (x.val == 0 && y.val == 0) ? 0.0f : 1.0f / (x.val * x.val + y.val * y.val);
    return dualfunc(y, x, f, -x.val * denom, y.val * denom);
}
inline float o_pow(float x, float y) { return fast_safe_pow(x, y); }
inline Df o_pow(const Df& u, const Df& v)
{
    float powuvm1 = fast_safe_pow(u.val, v.val - 1.0f);
    float powuv   = powuvm1 * u.val;
    float logu    = u.val > 0 ? fast_log(u.val) : 0.0f;
    return dualfunc(u, v, powuv, v.val * powuvm1, logu * powuv);
}
inline float o_fmod(float a, float b) { return safe_fmod(a, b); }
inline Df o_fmod(const Df& a, const Df& b) { return Df(safe_fmod(a.val, b.val), a.dx, a.dy); }
inline int o_fmod(int a, int b) { return o_mod(a, b); }
inline float o_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
// step has no derivative form (osl_step_fff only): the result's derivatives are zero
inline float o_step(const Df& edge, const Df& x) { return o_step(edge.val, x.val); }
inline float o_step(float edge, const Df& x) { return o_step(edge, x.val); }
inline float o_step(const Df& edge, float x) { return o_step(edge.val, x); }
// llvm_gen_minmax: min = select(x<=y, x, y), max = select(x>y, x, y)
inline float o_min(float a, float b) { return a <= b ? a : b; }
inline float o_max(float a, float b) { return a > b ? a : b; }
inline Df o_min(const Df& a, const Df& b) { return a.val <= b.val ? a : b; }
inline Df o_max(const Df& a, const Df& b) { return a.val > b.val ? a : b; }
inline int o_min(int a, int b) { return a <= b ? a : b; }
inline int o_max(int a, int b) { return a > b ? a : b; }
// llvm_gen_mix: r = a*(1-x) + b*x ; rx = ((ax*(1-x) - a*xx) + b*xx) + bx*x
inline float o_mix(float a, float b, float x) { return a * (1.0f - x) + b * x; }
inline Df o_mix(const Df& a, const Df& b, const Df& x)
{
    float omx = 1.0f - x.val;
    float r   = a.val * omx + b.val * x.val;
    float rx  = ((a.dx * omx - a.val * x.dx) + b.val * x.dx) + b.dx * x.val;
    float ry  = ((a.dy * omx - a.val * x.dy) + b.val * x.dy) + b.dy * x.val;
    return Df(r, rx, ry);
}
inline float o_smoothstep(float e0, float e1, float x)
{
    if (x < e0)
        return 0.0f;
    else if (x >= e1)
        return 1.0f;
    float t = (x - e0) / (e1 - e0);
    return (3.0f - 2.0f * t) * (t * t);
}
inline Df o_smoothstep(const Df& e0, const Df& e1, const Df& x)
{
    if (x.val < e0.val)
        return Df(0.0f);
    else if (x.val >= e1.val)
        return Df(1.0f);
    Df t = (x - e0) / (e1 - e0);
    return (3.0f - 2.0f * t) * t * t;
}
// llvm_gen_clamp: min(max(x,lo),hi) with selects
inline float o_clamp(float x, float lo, float hi) { float t = x < lo ? lo : x; return t > hi ? hi : t; }
inline Df o_clamp(const Df& x, const Df& lo, const Df& hi)
{
    Df t = x.val < lo.val ? lo : x;
    return t.val > hi.val ? hi : t;
}
inline int o_clamp(int x, int lo, int hi) { int t = x < lo ? lo : x; return t > hi ? hi : t; }
inline float o_select(float a, float b, float c) { return c != 0.0f ? b : a; }
inline Df o_select(const Df& a, const Df& b, const Df& c) { return c.val != 0.0f ? b : a; }

// ---------------------------------------------------------------------------
// component access: the generator's analogue of llvm_load_value/store_value
// ---------------------------------------------------------------------------
inline float getc(float a, int) { return a; }
inline float getc(int a, int) { return (float)a; }
inline Df getc(const Df& a, int) { return a; }
inline float getc(const V3& a, int c) { return a[c]; }
inline Df getc(const Dv& a, int c) { return Df(a.val[c], a.dx[c], a.dy[c]); }
inline void setc(float& d, int, float v) { d = v; }
inline void setc(float& d, int, const Df& v) { d = v.val; }
inline void setc(Df& d, int, float v) { d = Df(v); }
inline void setc(Df& d, int, const Df& v) { d = v; }
inline void setc(V3& d, int c, float v) { d[c] = v; }
inline void setc(V3& d, int c, const Df& v) { d[c] = v.val; }
inline void setc(Dv& d, int c, float v) { d.val[c] = v; d.dx[c] = 0; d.dy[c] = 0; }
inline void setc(Dv& d, int c, const Df& v) { d.val[c] = v.val; d.dx[c] = v.dx; d.dy[c] = v.dy; }
inline void setc(int& d, int, int v) { d = v; }
inline void setc(int& d, int, float v) { d = (int)v; }

// whole-value conversion used by `assign`
inline void assign(float& d, float s) { d = s; }
inline void assign(float& d, int s) { d = (float)s; }
inline void assign(float& d, const Df& s) { d = s.val; }
inline void assign(Df& d, float s) { d = Df(s); }
inline void assign(Df& d, int s) { d = Df((float)s); }
inline void assign(Df& d, const Df& s) { d = s; }
inline void assign(int& d, int s) { d = s; }
inline void assign(int& d, float s) { d = (int)s; }
inline void assign(int& d, const Df& s) { d = (int)s.val; }
inline void assign(V3& d, float s) { d = V3(s); }
inline void assign(V3& d, int s) { d = V3((float)s); }
inline void assign(V3& d, const Df& s) { d = V3(s.val); }
inline void assign(V3& d, const V3& s) { d = s; }
inline void assign(V3& d, const Dv& s) { d = s.val; }
inline void assign(Dv& d, float s) { d = Dv(V3(s)); }
inline void assign(Dv& d, int s) { d = Dv(V3((float)s)); }
inline void assign(Dv& d, const Df& s) { d = Dv(V3(s.val), V3(s.dx), V3(s.dy)); }
inline void assign(Dv& d, const V3& s) { d = Dv(s); }
inline void assign(Dv& d, const Dv& s) { d = s; }
inline void assign(const char*& d, const char* s) { d = s; }

// ---------------------------------------------------------------------------
// vector functions over component scalars (dual_vec.h:405-560, Imath Vec3)
// ---------------------------------------------------------------------------
inline float imath_length(const V3& v)
{
    float l2 = v.x * v.x + v.y * v.y + v.z * v.z;
    if (l2 < 2.0f * std::numeric_limits<float>::min()) {
        float ax = std::fabs(v.x), ay = std::fabs(v.y), az = std::fabs(v.z);
        float m = ax;
        if (m < ay) m = ay;
        if (m < az) m = az;
        if (m == 0.0f)
            return 0.0f;
        ax /= m; ay /= m; az /= m;
        return m * std::sqrt(ax * ax + ay * ay + az * az);
    }
    return std::sqrt(l2);
}
inline float o_dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Df o_dot(const Dv& a, const Dv& b)
{
    return comp(a, 0) * comp(b, 0) + comp(a, 1) * comp(b, 1) + comp(a, 2) * comp(b, 2);
}
inline Df o_dot(const Dv& a, const V3& b) { return o_dot(a, Dv(b)); }
inline Df o_dot(const V3& a, const Dv& b) { return o_dot(Dv(a), b); }
inline V3 o_cross(const V3& a, const V3& b) { return cross(a, b); }
inline Dv o_cross(const Dv& a, const Dv& b)
{
    Df ax = comp(a, 0), ay = comp(a, 1), az = comp(a, 2);
    Df bx = comp(b, 0), by = comp(b, 1), bz = comp(b, 2);
    return make_dv(ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx);
}
inline Dv o_cross(const Dv& a, const V3& b) { return o_cross(a, Dv(b)); }
inline Dv o_cross(const V3& a, const Dv& b) { return o_cross(Dv(a), b); }
inline float o_length(const V3& a) { return imath_length(a); }
inline Df o_length(const Dv& a)
{
    Df ax = comp(a, 0), ay = comp(a, 1), az = comp(a, 2);
    return d_sqrt(ax * ax + ay * ay + az * az);
}
inline float o_distance(const V3& a, const V3& b)
{
    float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    return std::sqrt(x * x + y * y + z * z);
}
inline Df o_distance(const Dv& a, const Dv& b) { return o_length(a - b); }
inline Df o_distance(const Dv& a, const V3& b) { return o_length(a - Dv(b)); }
inline Df o_distance(const V3& a, const Dv& b) { return o_length(Dv(a) - b); }
inline V3 o_normalize(const V3& a)
{
    V3 v      = a;
    float len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (len > 0.0f) {
        float inv = 1.0f / len;
        v.x *= inv; v.y *= inv; v.z *= inv;
    } else
        v = V3(0.0f);
    return v;
}
inline Dv o_normalize(const Dv& a)
{
    Df ax = comp(a, 0), ay = comp(a, 1), az = comp(a, 2);
    Df len = d_sqrt(ax * ax + ay * ay + az * az);
    if (len.val > 0.0f) {
        Df inv = 1.0f / len;
        return make_dv(ax * inv, ay * inv, az * inv);
    }
    return Dv(V3(0.0f));
}
inline float filter_width(float dx, float dy) { return std::sqrt(dx * dx + dy * dy); }
inline float o_filterwidth(const Df& x) { return filter_width(x.dx, x.dy); }
inline float o_filterwidth(float) { return 0.0f; }
inline V3 o_filterwidth(const Dv& x)
{
    return V3(filter_width(x.dx.x, x.dy.x), filter_width(x.dx.y, x.dy.y),
              filter_width(x.dx.z, x.dy.z));
}
inline V3 o_filterwidth(const V3&) { return V3(0.0f); }
inline V3 o_calculatenormal(const Dv& P, bool flip) { return flip ? cross(P.dy, P.dx) : cross(P.dx, P.dy); }
inline float o_area(const Dv& P) { return imath_length(cross(P.dx, P.dy)); }
// opcolor.cpp osl_luminance: Rec.709 weights of the default "Rec709" colorspace
inline float o_luminance(const V3& c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }
inline Df o_luminance(const Dv& c)
{
    return Df(o_luminance(c.val), o_luminance(c.dx), o_luminance(c.dy));
}

// Dx/Dy (llvm_gen_DxDy): move a partial into the value slot, zero derivs
inline float o_Dx(const Df& a) { return a.dx; }
inline float o_Dy(const Df& a) { return a.dy; }
inline V3 o_Dx(const Dv& a) { return a.dx; }
inline V3 o_Dy(const Dv& a) { return a.dy; }
inline float o_Dx(float) { return 0.0f; }
inline float o_Dy(float) { return 0.0f; }
inline V3 o_Dx(const V3&) { return V3(0.0f); }
inline V3 o_Dy(const V3&) { return V3(0.0f); }

// ---------------------------------------------------------------------------
// noise front ends (opnoise.cpp:71-272, 276-470; llvm_gen.cpp:3117-3299).
// KIND: 0 noise(uperlin) 1 snoise(perlin) 2 cellnoise 3 hashnoise
// ---------------------------------------------------------------------------
enum { N_NOISE = 0, N_SNOISE = 1, N_CELL = 2, N_HASH = 3, N_SIMPLEX = 4, N_USIMPLEX = 5 };

template<int KIND, class S, int NC> inline void noise_core(S* out, int dim, const S* in)
{
    if (KIND == N_SIMPLEX)
        simplex_nd<NC, false>(out, dim, in);
    else if (KIND == N_USIMPLEX)
        simplex_nd<NC, true>(out, dim, in);
    else if (KIND == N_NOISE)
        perlin_nd<S, NC, false>(out, dim, in, nullptr);
    else
        perlin_nd<S, NC, true>(out, dim, in, nullptr);
}
template<int KIND, int NC> inline void ihnoise_core(float* out, int dim, const float* in)
{
    const int K = KIND == N_CELL ? 0 : 1;
    if (NC == 1) {
        switch (dim) {
        case 1: out[0] = ihnoise_f<K>(in[0]); break;
        case 2: out[0] = ihnoise_f<K>(in[0], in[1]); break;
        case 3: out[0] = ihnoise_f<K>(V3(in[0], in[1], in[2])); break;
        default: out[0] = ihnoise_f<K>(V3(in[0], in[1], in[2]), in[3]); break;
        }
    } else {
        V3 r;
        switch (dim) {
        case 1: r = ihnoise_v<K>(in[0]); break;
        case 2: r = ihnoise_v<K>(in[0], in[1]); break;
        case 3: r = ihnoise_v<K>(V3(in[0], in[1], in[2])); break;
        default: r = ihnoise_v<K>(V3(in[0], in[1], in[2]), in[3]); break;
        }
        out[0] = r.x; out[1] = r.y; out[2] = r.z;
    }
}

// ---------------------------------------------------------------------------
// spline / splineinverse (src/liboslexec/splineimpl.h:15-295, opspline.cpp)
// ---------------------------------------------------------------------------
struct SplineBasis {
    int step;
    float b[4][4];
};
enum { SPL_CATMULLROM, SPL_BEZIER, SPL_BSPLINE, SPL_HERMITE, SPL_LINEAR, SPL_CONSTANT };
static const SplineBasis g_spline_basis[6] = {
    { 1, { { (-1.0f / 2.0f), (3.0f / 2.0f), (-3.0f / 2.0f), (1.0f / 2.0f) },
           { (2.0f / 2.0f), (-5.0f / 2.0f), (4.0f / 2.0f), (-1.0f / 2.0f) },
           { (-1.0f / 2.0f), (0.0f / 2.0f), (1.0f / 2.0f), (0.0f / 2.0f) },
           { (0.0f / 2.0f), (2.0f / 2.0f), (0.0f / 2.0f), (0.0f / 2.0f) } } },
    { 3, { { -1, 3, -3, 1 }, { 3, -6, 3, 0 }, { -3, 3, 0, 0 }, { 1, 0, 0, 0 } } },
    { 1, { { (-1.0f / 6.0f), (3.0f / 6.0f), (-3.0f / 6.0f), (1.0f / 6.0f) },
           { (3.0f / 6.0f), (-6.0f / 6.0f), (3.0f / 6.0f), (0.0f / 6.0f) },
           { (-3.0f / 6.0f), (0.0f / 6.0f), (3.0f / 6.0f), (0.0f / 6.0f) },
           { (1.0f / 6.0f), (4.0f / 6.0f), (1.0f / 6.0f), (0.0f / 6.0f) } } },
    { 2, { { 2, 1, -2, 1 }, { -3, -2, 3, -1 }, { 0, 1, 0, 0 }, { 1, 0, 0, 0 } } },
    { 1, { { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, -1, 1, 0 }, { 0, 1, 0, 0 } } },
    { 1, { { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 } } }
};
inline int spline_type(const char* n)
{
    if (!n) return SPL_LINEAR;
    if (!std::strcmp(n, "catmull-rom") || !std::strcmp(n, "catmullrom")) return SPL_CATMULLROM;
    if (!std::strcmp(n, "bezier")) return SPL_BEZIER;
    if (!std::strcmp(n, "bspline")) return SPL_BSPLINE;
    if (!std::strcmp(n, "hermite")) return SPL_HERMITE;
    if (!std::strcmp(n, "constant")) return SPL_CONSTANT;
    return SPL_LINEAR;
}
inline Dv operator*(const V3& a, const Df& b) { return Dv(a * b.val, a * b.dx, a * b.dy); }
inline Dv operator+(const Dv& a, const V3& b) { return Dv(a.val + b, a.dx, a.dy); }
inline float sclamp01(float a) { return (a >= 0.0f) ? ((a <= 1.0f) ? a : 1.0f) : 0.0f; }
inline Df sclamp01(const Df& a) { return (a.val >= 0.0f) ? ((a.val <= 1.0f) ? a : Df(1.0f)) : Df(0.0f); }
// R result, X abscissa (float|Df), K knot element type (float|Df|V3|Dv)
template<class R, class X, class K>
inline void spline_eval(R& result, const X& xval, const K* knots, int knot_count, int type)
{
    const SplineBasis& sp = g_spline_basis[type];
    X x        = sclamp01(xval);
    int nsegs  = ((knot_count - 4) / sp.step) + 1;
    x          = x * (float)nsegs;
    float segx = nd(x);
    int segnum = (int)segx;
    if (segnum < 0) segnum = 0;
    if (segnum > (nsegs - 1)) segnum = nsegs - 1;
    if (type == SPL_CONSTANT) {
        assign(result, nd(knots[segnum + 1]));
        return;
    }
    x     = x - float(segnum);
    int s = segnum * sp.step;
    K P[4];
    for (int k = 0; k < 4; k++)
        P[k] = knots[s + k];
    K tk[4];
    for (int k = 0; k < 4; k++)
        tk[k] = sp.b[k][0] * P[0] + sp.b[k][1] * P[1] + sp.b[k][2] * P[2] + sp.b[k][3] * P[3];
    auto t = (tk[0] * x + tk[1]);
    auto t2 = (t * x + tk[2]);
    auto t3 = (t2 * x + tk[3]);
    assign(result, t3);
}
// splineinverse: OIIO::invert (regula falsi / bisection hybrid, 32 iterations, eps 1e-6)
template<class F> inline float oiio_invert(F& func, float y, float xmin, float xmax, int maxiters, float eps, bool* brack)
{
    float v0 = func(xmin), v1 = func(xmax);
    float x = xmin, v = v0;
    bool increasing = (v0 < v1);
    float vmin = increasing ? v0 : v1, vmax = increasing ? v1 : v0;
    bool bracketed = (y >= vmin && y <= vmax);
    if (brack) *brack = bracketed;
    if (!bracketed)
        return ((y < vmin) == increasing) ? xmin : xmax;
    if (std::fabs(v0 - v1) < eps)
        return x;
    int rfiters = (3 * maxiters) / 4;
    for (int iters = 0; iters < maxiters; ++iters) {
        float t;
        if (iters < rfiters) {
            t = (y - v0) / (v1 - v0);
            if (t <= 0.0f || t >= 1.0f)
                t = 0.5f;
        } else {
            t = 0.5f;
        }
        x = xmin * (1.0f - t) + xmax * t;   // OIIO::lerp
        v = func(x);
        if ((v < y) == increasing) {
            xmin = x;
            v0   = v;
        } else {
            xmax = x;
            v1   = v;
        }
        if (std::fabs(xmax - xmin) < eps || std::fabs(v - y) < eps)
            return x;
    }
    return x;
}
inline float spline_inverse(float y, const float* knots, int knot_count, int type)
{
    const SplineBasis& sp = g_spline_basis[type];
    int lowindex  = sp.step == 1 ? 1 : 0;
    int highindex = sp.step == 1 ? knot_count - 2 : knot_count - 1;
    bool increasing = knots[1] < knots[knot_count - 2];
    if (increasing) {
        if (y <= knots[lowindex]) return 0.0f;
        if (y >= knots[highindex]) return 1.0f;
    } else {
        if (y >= knots[lowindex]) return 0.0f;
        if (y <= knots[highindex]) return 1.0f;
    }
    auto S = [&](float x) { float v; spline_eval(v, x, knots, knot_count, type); return v; };
    int nsegs     = (knot_count - 4) / sp.step + 1;
    float nseginv = 1.0f / nsegs;
    float r0 = 0.0f, x = 0.0f;
    for (int s = 0; s < nsegs; ++s) {
        float r1 = nseginv * (s + 1);
        bool brack;
        x = oiio_invert(S, y, r0, r1, 32, 1.0e-6f, &brack);
        if (brack)
            return x;
        r0 = r1;
    }
    return x;
}
// osl_splineinverse_dfdff: the same search instantiated on Dual2<float> - comparisons look at
// the values, the arithmetic carries the derivatives of y through the regula falsi steps
template<class F> inline Df oiio_invert(F& func, const Df& y, Df xmin, Df xmax, int maxiters, float eps, bool* brack)
{
    Df v0 = func(xmin), v1 = func(xmax);
    Df x = xmin, v = v0;
    bool increasing = (v0.val < v1.val);
    Df vmin = increasing ? v0 : v1, vmax = increasing ? v1 : v0;
    bool bracketed = (y.val >= vmin.val && y.val <= vmax.val);
    if (brack) *brack = bracketed;
    if (!bracketed)
        return ((y.val < vmin.val) == increasing) ? xmin : xmax;
    if (std::fabs((v0 - v1).val) < eps)
        return x;
    int rfiters = (3 * maxiters) / 4;
    for (int iters = 0; iters < maxiters; ++iters) {
        Df t;
        if (iters < rfiters) {
            t = (y - v0) / (v1 - v0);
            if (t.val <= 0.0f || t.val >= 1.0f)
                t = Df(0.5f);
        } else {
            t = Df(0.5f);
        }
        x = lerp(xmin, xmax, t);
        v = func(x);
        if ((v.val < y.val) == increasing) {
            xmin = x;
            v0   = v;
        } else {
            xmax = x;
            v1   = v;
        }
        if (std::fabs((xmax - xmin).val) < eps || std::fabs((v - y).val) < eps)
            return x;
    }
    return x;
}
inline Df spline_inverse(const Df& y, const float* knots, int knot_count, int type)
{
    const SplineBasis& sp = g_spline_basis[type];
    int lowindex  = sp.step == 1 ? 1 : 0;
    int highindex = sp.step == 1 ? knot_count - 2 : knot_count - 1;
    bool increasing = knots[1] < knots[knot_count - 2];
    if (increasing) {
        if (y.val <= knots[lowindex]) return Df(0.0f);
        if (y.val >= knots[highindex]) return Df(1.0f);
    } else {
        if (y.val >= knots[lowindex]) return Df(0.0f);
        if (y.val <= knots[highindex]) return Df(1.0f);
    }
    auto S = [&](const Df& x) { Df v; spline_eval(v, x, knots, knot_count, type); return v; };
    int nsegs     = (knot_count - 4) / sp.step + 1;
    float nseginv = 1.0f / nsegs;
    Df r0(0.0f), x(0.0f);
    for (int s = 0; s < nsegs; ++s) {
        Df r1(nseginv * (s + 1));
        bool brack;
        x = oiio_invert(S, y, r0, r1, 32, 1.0e-6f, &brack);
        if (brack)
            return x;
        r0 = r1;
    }
    return x;
}

}  // namespace oslo
