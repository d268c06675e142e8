// osl_b200_color.cuh — colour shadeops on the device.
//
//   osl_blackbody_vf, osl_wavelength_color_vf, osl_luminance_fv/_dfdv,
//   osl_prepend_color_from, osl_transformc        src/liboslexec/opcolor.cpp:434-518
//   ColorSystem::to_rgb / from_rgb / transformc   src/liboslexec/opcolor.cpp:287-376
//   hsv/hsl/YIQ/xyY/sRGB conversions, blackbody
//   table lookup, wavelength_color_XYZ            src/liboslexec/opcolor_impl.h:103-575
//
// The colour system is a constant array baked into the generated module by the
// host (csrc/host/osl_b200_color.cpp; layout: XYZ2RGB[9] RGB2XYZ[9]
// luminance[3] blackbody[317][3]); `cs` below points at it.  Colour-space
// names are resolved to the OSLD_CS_* codes by the code generator (constant
// names) or by an id compare chain (run-time names), so no string handling
// happens here.  Included from generated code only when a colour op is used.
#pragma once
#define OSLB200_CIE_DECL static __device__
#include "osl_b200_cie1931.cuh"

namespace osld {

enum { OSLD_CS_UNKNOWN = 0, OSLD_CS_RGB, OSLD_CS_HSV, OSLD_CS_HSL, OSLD_CS_YIQ, OSLD_CS_XYZ, OSLD_CS_XYY, OSLD_CS_SRGB };

OSLD float cvof(float a) { return a; }
OSLD float cvof(Df a) { return a.val; }
OSLD void cset(float& d, float v) { d = v; }
OSLD void cset(Df& d, float v) { d = mkd(v); }
OSLD Df operator/(Df a, float b)
{
    float binv = 1.0f / b;
    return mkd(a.val / b, binv * a.dx, binv * a.dy);
}
template<class S> OSLD S cmin(S l, S r) { return (cvof(r) > cvof(l)) ? l : r; }
template<class S> OSLD S cmax(S l, S r) { return (cvof(r) > cvof(l)) ? r : l; }

// row vector * 3x3, M row-major
template<class S> OSLD void cs_mul33(S& x, S& y, S& z, const float* M)
{
    S a = x * M[0] + y * M[3] + z * M[6];
    S b = x * M[1] + y * M[4] + z * M[7];
    S c = x * M[2] + y * M[5] + z * M[8];
    x = a;
    y = b;
    z = c;
}
template<class S> OSLD void hsv_to_rgb(S& x, S& y, S& z)
{
    S h = x, s = y, v = z;
    if (cvof(s) < 0.0001f) {
        x = v; y = v; z = v;
        return;
    }
    h      = 6.0f * (h - floorf(cvof(h)));
    int hi = (int)floorf(cvof(h));
    S f    = h - (float)hi;
    S p    = v * (1.0f - s);
    S q    = v * (1.0f - s * f);
    S t    = v * (1.0f - s * (1.0f - f));
    // select instead of a switch: the sextant differs lane to lane
    x = hi == 0 ? v : hi == 1 ? q : hi == 2 ? p : hi == 3 ? p : hi == 4 ? t : v;
    y = hi == 0 ? t : hi == 1 ? v : hi == 2 ? v : hi == 3 ? q : hi == 4 ? p : p;
    z = hi == 0 ? p : hi == 1 ? p : hi == 2 ? t : hi == 3 ? v : hi == 4 ? v : q;
}
template<class S> OSLD void rgb_to_hsv(S& x, S& y, S& z)
{
    S r = x, g = y, b = z;
    S mincomp = cmin(r, cmin(g, b));
    S maxcomp = cmax(r, cmax(g, b));
    S delta   = maxcomp - mincomp;
    S v       = maxcomp;
    S s;
    cset(s, 0.0f);
    if (cvof(maxcomp) > 0.0f)
        s = delta / maxcomp;
    S h;
    cset(h, 0.0f);
    if (cvof(s) > 0.0f) {
        float k;
        S xx, yy;
        if (cvof(r) >= cvof(maxcomp)) {
            k  = 0.0f / 6.0f;
            xx = g;
            yy = b;
        } else if (cvof(g) >= cvof(maxcomp)) {
            k  = 2.0f / 6.0f;
            xx = b;
            yy = r;
        } else {
            k  = 4.0f / 6.0f;
            xx = r;
            yy = g;
        }
        h = k + (xx - yy) / (6.0f * delta);
        if (cvof(h) < 0.0f)
            h = h + 1.0f;
    }
    x = h; y = s; z = v;
}
template<class S> OSLD void hsl_to_rgb(S& x, S& y, S& z)
{
    S h = x, s = y, l = z;
    S v = (cvof(l) <= 0.5f) ? (l * (1.0f + s)) : (l * (1.0f - s) + s);
    if (cvof(v) <= 0.0f) {
        cset(x, 0.0f); cset(y, 0.0f); cset(z, 0.0f);
    } else {
        S mn = 2.0f * l - v;
        s    = (v - mn) / v;
        x = h; y = s; z = v;
        hsv_to_rgb(x, y, z);
    }
}
template<class S> OSLD void rgb_to_hsl(S& x, S& y, S& z)
{
    S minval = cmin(x, cmin(y, z));
    rgb_to_hsv(x, y, z);
    S maxval = z;
    S h = x, s, l = (minval + maxval) / 2.0f;
    if (cvof(minval) == cvof(maxval))
        cset(s, 0.0f);
    else {
        S min2max = (maxval - minval);
        S divisor = (cvof(l) <= 0.5f) ? (maxval + minval) : (2.0f - maxval - minval);
        s         = min2max / divisor;
    }
    x = h; y = s; z = l;
}
template<class S> OSLD void XYZ_to_xyY(S& x, S& y, S& z)
{
    S n = (x + y + z);
    S n_inv;
    if (cvof(n) >= 1.0e-6f)
        n_inv = 1.0f / n;
    else
        cset(n_inv, 0.0f);
    S X = x, Y = y;
    x = X * n_inv; y = Y * n_inv; z = Y;
}
template<class S> OSLD void xyY_to_XYZ(S& x, S& y, S& z)
{
    S Y = z, Y_y;
    if (cvof(y) > 1.0e-6f)
        Y_y = Y / y;
    else
        cset(Y_y, 0.0f);
    S X = Y_y * x;
    S Z = Y_y * (1.0f - x - y);
    x = X; y = Y; z = Z;
}
// OIIO::safe_pow: guards around powf (the reference calls std::pow here, not fast_*)
OSLD float c_safe_pow(float x, float y)
{
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return 0.0f;
    if ((x < 0.0f) && (y != floorf(y))) return 0.0f;
    float r = powf(x, y);
    return fminf(fmaxf(r, -OSLD_FLT_MAX), OSLD_FLT_MAX);
}
OSLD float cpow(float x, float y) { return c_safe_pow(x, y); }
OSLD Df cpow(Df u, float v)
{
    // dual.h:1058-1068 (v has no derivatives here)
    float powuvm1 = c_safe_pow(u.val, v - 1.0f);
    float powuv   = powuvm1 * u.val;
    float logu    = u.val > 0 ? logf(u.val) : 0.0f;
    float fu = v * powuvm1, fv = logu * powuv;
    return mkd(powuv, fu * u.dx + fv * 0.0f, fu * u.dy + fv * 0.0f);
}
template<class S> OSLD S srgb_to_linear1(S x)
{
    return (cvof(x) <= 0.04045f) ? (x * (1.0f / 12.92f)) : cpow((x + 0.055f) * (1.0f / 1.055f), 2.4f);
}
template<class S> OSLD S linear_to_srgb1(S x)
{
    return (cvof(x) <= 0.0031308f) ? (12.92f * x) : (1.055f * cpow(x, 1.f / 2.4f) - 0.055f);
}

// ColorSystem::transformc on three scalars; from/to are OSLD_CS_* codes
template<class S> OSLD void color_transform(const float* cs, int from, int to, S& x, S& y, S& z)
{
    const float YIQ2RGB[9] = { 1.0000f, 1.0000f, 1.0000f, 0.9557f, -0.2716f, -1.1082f, 0.6199f, -0.6469f, 1.7051f };
    const float RGB2YIQ[9] = { 0.299f, 0.596f, 0.212f, 0.587f, -0.275f, -0.523f, 0.114f, -0.321f, 0.311f };
    if (from == OSLD_CS_UNKNOWN || to == OSLD_CS_UNKNOWN)
        return;  // would be an OpenColorIO transform; the colour passes through
    switch (from) {
    case OSLD_CS_HSV: hsv_to_rgb(x, y, z); break;
    case OSLD_CS_HSL: hsl_to_rgb(x, y, z); break;
    case OSLD_CS_YIQ: cs_mul33(x, y, z, YIQ2RGB); break;
    case OSLD_CS_XYZ: cs_mul33(x, y, z, cs); break;
    case OSLD_CS_XYY: xyY_to_XYZ(x, y, z); cs_mul33(x, y, z, cs); break;
    case OSLD_CS_SRGB: x = srgb_to_linear1(x); y = srgb_to_linear1(y); z = srgb_to_linear1(z); break;
    default: break;
    }
    switch (to) {
    case OSLD_CS_HSV: rgb_to_hsv(x, y, z); break;
    case OSLD_CS_HSL: rgb_to_hsl(x, y, z); break;
    case OSLD_CS_YIQ: cs_mul33(x, y, z, RGB2YIQ); break;
    case OSLD_CS_XYZ: cs_mul33(x, y, z, cs + 9); break;
    case OSLD_CS_XYY: cs_mul33(x, y, z, cs + 9); XYZ_to_xyY(x, y, z); break;
    case OSLD_CS_SRGB: x = linear_to_srgb1(x); y = linear_to_srgb1(y); z = linear_to_srgb1(z); break;
    default: break;
    }
}
OSLD V3 color_transformc(const float* cs, int from, int to, V3 C)
{
    color_transform(cs, from, to, C.x, C.y, C.z);
    return C;
}
OSLD Dv color_transformc(const float* cs, int from, int to, const Dv& C)
{
    Df x = getc(C, 0), y = getc(C, 1), z = getc(C, 2);
    color_transform(cs, from, to, x, y, z);
    Dv r;
    setc(r, 0, x);
    setc(r, 1, y);
    setc(r, 2, z);
    return r;
}
OSLD float color_luminance(const float* cs, V3 c) { return c.x * cs[18] + c.y * cs[19] + c.z * cs[20]; }
OSLD Df color_luminance(const float* cs, const Dv& c)
{
    return mkd(color_luminance(cs, c.val), color_luminance(cs, c.dx), color_luminance(cs, c.dy));
}
// Planck spectrum against the CIE observer, for temperatures above the table
OSLD V3 blackbody_XYZ(float temp)
{
    float X = 0, Y = 0, Z = 0;
    const float dlambda = (float)(5.0f * 1e-9);
    for (int i = 0; i < 81; ++i) {
        float lambda   = 380.0f + 5.0f * i;
        float wlm      = lambda * 1e-9f;
        const float c1 = 3.74183e-16f, c2 = 1.4388e-2f;
        float wlm2 = wlm * wlm, wlm4 = wlm2 * wlm2, wlm5 = wlm4 * wlm;
        float inv5 = 1.0f / wlm5;
        float Me   = ((c1 * inv5) / fast_expm1(c2 / (wlm * temp))) * dlambda;
        X += Me * cie_xbar[i];
        Y += Me * cie_ybar[i];
        Z += Me * cie_zbar[i];
    }
    return mkv(X, Y, Z);
}
OSLD V3 color_blackbody(const float* cs, float T)
{
    if (T < 12000.0f) {
        if (T < 800.0f)
            return mkv(1.0e-6f, 0.0f, 0.0f);
        float t  = (T - 800.0f) / 2.0f;
        float ic = fast_cbrt(t);
        t        = ic * ic;
        int ti   = (int)t;
        float r  = t - ti;
        const float* e = cs + 21 + 3 * ti;
        V3 a = mkv(e[0], e[1], e[2]), b = mkv(e[3], e[4], e[5]);
        V3 rgb  = a * mkv(1.0f - r) + b * mkv(r);
        V3 rgb2 = rgb * rgb;
        V3 rgb4 = rgb2 * rgb2;
        return rgb4 * rgb;
    }
    V3 xyz = blackbody_XYZ(T);
    cs_mul33(xyz.x, xyz.y, xyz.z, cs);
    return mkv(xyz.x < 0.0f ? 0.0f : xyz.x, xyz.y < 0.0f ? 0.0f : xyz.y, xyz.z < 0.0f ? 0.0f : xyz.z);
}
OSLD V3 color_wavelength(const float* cs, float lambda_nm)
{
    V3 XYZ   = mkv(0.0f);
    float ii = (lambda_nm - 380.0f) / 5.0f;
    int i    = (int)ii;
    if (!((i < 0) | (i >= 80))) {
        float r = ii - i;
        V3 a = mkv(cie_xbar[i], cie_ybar[i], cie_zbar[i]), b = mkv(cie_xbar[i + 1], cie_ybar[i + 1], cie_zbar[i + 1]);
        XYZ  = a * mkv(1.0f - r) + b * mkv(r);
    }
    cs_mul33(XYZ.x, XYZ.y, XYZ.z, cs);
    V3 rgb = XYZ * (float)(1.0 / 2.52);
    return mkv(rgb.x < 0.0f ? 0.0f : rgb.x, rgb.y < 0.0f ? 0.0f : rgb.y, rgb.z < 0.0f ? 0.0f : rgb.z);
}

}  // namespace osld
