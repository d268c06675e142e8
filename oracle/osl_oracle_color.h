// osl_oracle_color.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of OSL's colour shadeops:
//   ColorSystem::set_colorspace, chromaticity table   src/liboslexec/opcolor.cpp:26-236
//   to_rgb / from_rgb / transformc                    src/liboslexec/opcolor.cpp:287-395
//   osl_blackbody_vf, osl_wavelength_color_vf,
//   osl_luminance_*, osl_prepend_color_from,
//   osl_transformc                                    src/liboslexec/opcolor.cpp:434-518
//   wavelength_color_XYZ, bb_spectrum, spectrum_to_XYZ,
//   hsv/hsl/YIQ/xyY/sRGB conversions, blackbody table  src/liboslexec/opcolor_impl.h:103-575
//   XYZ_to_RGB / RGB_to_XYZ / luminance               src/liboslexec/opcolor.h:47-62
// Third-party arithmetic restated from its published form: Imath 3.1
// Matrix33::inverse (cofactors / determinant with the singular guard),
// Vec3 * Matrix33 (row vector), OIIO safe_pow (guards around std::pow),
// fast_expm1 / fast_cbrt (osl_oracle_ops.h).  OCIO transforms are outside the
// path: an unknown space leaves the colour unchanged, as the reference does
// after reporting the error (opcolor.cpp:243-270).
#pragma once
#include "cie1931_5nm.h"

namespace oslo {

struct ColorSystem {
    std::string colorspace;
    float XYZ2RGB[3][3], RGB2XYZ[3][3];
    V3 luminance_scale;
    V3 blackbody_table[317];
};

namespace color_impl {

const float BB_DRAPER = 800.0f, BB_MAX_TABLE_RANGE = 12000.0f, BB_TABLE_YPOWER = 5.0f, BB_TABLE_SPACING = 2.0f;

inline float BB_TABLE_MAP(float i)
{
    float is = std::sqrt(i);
    float ip = is * is * is;
    return ip * BB_TABLE_SPACING + BB_DRAPER;
}
inline float BB_TABLE_UNMAP(float T)
{
    float t  = (T - BB_DRAPER) / BB_TABLE_SPACING;
    float ic = fast_cbrt(t);
    return ic * ic;
}
inline float bb_spectrum(float temp, float wavelength_nm)
{
    float wlm      = wavelength_nm * 1e-9f;
    const float c1 = 3.74183e-16f;
    const float c2 = 1.4388e-2f;
    const float wlm2 = wlm * wlm, wlm4 = wlm2 * wlm2, wlm5 = wlm4 * wlm;
    const float inverse_of_wlm5 = 1.0f / wlm5;
    return float((c1 * inverse_of_wlm5) / fast_expm1(c2 / (wlm * temp)));
}
inline V3 blackbody_XYZ(float temp)
{
    float X = 0, Y = 0, Z = 0;
    const float dlambda = 5.0f * 1e-9;
    for (int i = 0; i < 81; ++i) {
        float lambda = 380.0f + 5.0f * i;
        float Me     = bb_spectrum(temp, lambda) * dlambda;
        X += Me * cie_xbar[i];
        Y += Me * cie_ybar[i];
        Z += Me * cie_zbar[i];
    }
    return V3(X, Y, Z);
}
inline void clamp_zero(V3& c)
{
    if (c.x < 0.0f) c.x = 0.0f;
    if (c.y < 0.0f) c.y = 0.0f;
    if (c.z < 0.0f) c.z = 0.0f;
}
// row vector * 3x3 (Imath Vec3 * Matrix33; dual_vec.h:305-316 for duals)
template<class S> inline void mul_m33(S& x, S& y, S& z, const float M[3][3])
{
    S a = x * M[0][0] + y * M[1][0] + z * M[2][0];
    S b = x * M[0][1] + y * M[1][1] + z * M[2][1];
    S c = x * M[0][2] + y * M[1][2] + z * M[2][2];
    x = a;
    y = b;
    z = c;
}
inline V3 mul_m33(const V3& v, const float M[3][3])
{
    float x = v.x, y = v.y, z = v.z;
    mul_m33(x, y, z, M);
    return V3(x, y, z);
}
inline bool m33_inverse(const float x[3][3], float out[3][3])
{
    float s[3][3] = {
        { x[1][1] * x[2][2] - x[2][1] * x[1][2], x[2][1] * x[0][2] - x[0][1] * x[2][2], x[0][1] * x[1][2] - x[1][1] * x[0][2] },
        { x[2][0] * x[1][2] - x[1][0] * x[2][2], x[0][0] * x[2][2] - x[2][0] * x[0][2], x[1][0] * x[0][2] - x[0][0] * x[1][2] },
        { x[1][0] * x[2][1] - x[2][0] * x[1][1], x[2][0] * x[0][1] - x[0][0] * x[2][1], x[0][0] * x[1][1] - x[1][0] * x[0][1] }
    };
    float r = x[0][0] * s[0][0] + x[0][1] * s[1][0] + x[0][2] * s[2][0];
    bool ok = true;
    if (std::fabs(r) >= 1) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s[i][j] /= r;
    } else {
        float mr = std::fabs(r) / std::numeric_limits<float>::min();
        for (int i = 0; i < 3 && ok; ++i)
            for (int j = 0; j < 3 && ok; ++j) {
                if (mr > std::fabs(s[i][j]))
                    s[i][j] /= r;
                else
                    ok = false;
            }
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            out[i][j] = ok ? s[i][j] : (i == j ? 1.0f : 0.0f);
    return ok;
}

struct Chroma {
    const char* name;
    float xRed, yRed, xGreen, yGreen, xBlue, yBlue, xWhite, yWhite;
};
#define OSLO_ILL_C 0.3101, 0.3162
#define OSLO_ILL_D65 0.3127, 0.3291
#define OSLO_ILL_E 0.33333333, 0.33333333
#define OSLO_ILL_ACES 0.32168, 0.33767
static const Chroma k_color_systems[13] = {
    { "Rec709", 0.64, 0.33, 0.30, 0.60, 0.15, 0.06, OSLO_ILL_D65 },
    { "sRGB", 0.64, 0.33, 0.30, 0.60, 0.15, 0.06, OSLO_ILL_D65 },
    { "NTSC", 0.67, 0.33, 0.21, 0.71, 0.14, 0.08, OSLO_ILL_C },
    { "EBU", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, OSLO_ILL_D65 },
    { "PAL", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, OSLO_ILL_D65 },
    { "SECAM", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, OSLO_ILL_D65 },
    { "SMPTE", 0.630, 0.340, 0.310, 0.595, 0.155, 0.070, OSLO_ILL_D65 },
    { "HDTV", 0.670, 0.330, 0.210, 0.710, 0.150, 0.060, OSLO_ILL_D65 },
    { "CIE", 0.7355, 0.2645, 0.2658, 0.7243, 0.1669, 0.0085, OSLO_ILL_E },
    { "AdobeRGB", 0.64, 0.33, 0.21, 0.71, 0.15, 0.06, OSLO_ILL_D65 },
    { "XYZ", 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, OSLO_ILL_E },
    { "ACES2065-1", 0.7347, 0.2653, 0.0, 1.0, 0.0001, -0.077, OSLO_ILL_ACES },
    { "ACEScg", 0.713, 0.293, 0.165, 0.83, 0.128, 0.044, OSLO_ILL_ACES },
};

// ---- per-colour conversions, S = float or Df ------------------------------------------
template<class S> inline S min_val(const S& l, const S& r) { return (val_of(r) > val_of(l)) ? l : r; }
template<class S> inline S max_val(const S& l, const S& r) { return (val_of(r) > val_of(l)) ? r : l; }

template<class S> inline void hsv_to_rgb(S& x, S& y, S& z)
{
    S h = x, s = y, v = z;
    if (val_of(s) < 0.0001f) {
        x = v; y = v; z = v;
        return;
    }
    h    = 6.0f * (h - std::floor(val_of(h)));
    int hi = (int)std::floor(val_of(h));
    S f  = h - S(float(hi));
    S p  = v * (1.0f - s);
    S q  = v * (1.0f - s * f);
    S t  = v * (1.0f - s * (1.0f - f));
    switch (hi) {
    case 0: x = v; y = t; z = p; break;
    case 1: x = q; y = v; z = p; break;
    case 2: x = p; y = v; z = t; break;
    case 3: x = p; y = q; z = v; break;
    case 4: x = t; y = p; z = v; break;
    default: x = v; y = p; z = q; break;
    }
}
template<class S> inline void rgb_to_hsv(S& x, S& y, S& z)
{
    S r = x, g = y, b = z;
    S mincomp = min_val(r, min_val(g, b));
    S maxcomp = max_val(r, max_val(g, b));
    S delta   = maxcomp - mincomp;
    S v       = maxcomp;
    S s       = S(0.0f);
    if (val_of(maxcomp) > 0.0f)
        s = delta / maxcomp;
    S h = S(0.0f);
    if (val_of(s) > 0.0f) {
        float k;
        S xx, yy;
        if (val_of(r) >= val_of(maxcomp)) {
            k  = 0.0f / 6.0f;
            xx = g;
            yy = b;
        } else if (val_of(g) >= val_of(maxcomp)) {
            k  = 2.0f / 6.0f;
            xx = b;
            yy = r;
        } else {
            k  = 4.0f / 6.0f;
            xx = r;
            yy = g;
        }
        h = k + (xx - yy) / (6.0f * delta);
        if (val_of(h) < 0.0f)
            h = h + 1.0f;
    }
    x = h; y = s; z = v;
}
template<class S> inline void hsl_to_rgb(S& x, S& y, S& z)
{
    S h = x, s = y, l = z;
    S v = (val_of(l) <= 0.5f) ? (l * (1.0f + s)) : (l * (1.0f - s) + s);
    if (val_of(v) <= 0.0f) {
        x = S(0.0f); y = S(0.0f); z = S(0.0f);
    } else {
        S mn = 2.0f * l - v;
        s    = (v - mn) / v;
        x = h; y = s; z = v;
        hsv_to_rgb(x, y, z);
    }
}
template<class S> inline void rgb_to_hsl(S& x, S& y, S& z)
{
    S minval = min_val(x, min_val(y, z));
    rgb_to_hsv(x, y, z);
    S maxval = z;
    S h = x, s, l = (minval + maxval) / 2.0f;
    if (val_of(minval) == val_of(maxval))
        s = S(0.0f);
    else {
        S min2max = (maxval - minval);
        S divisor;
        if (val_of(l) <= 0.5f)
            divisor = (maxval + minval);
        else
            divisor = (2.0f - maxval - minval);
        s = min2max / divisor;
    }
    x = h; y = s; z = l;
}
static const float M_YIQ2RGB[3][3] = { { 1.0000, 1.0000, 1.0000 }, { 0.9557, -0.2716, -1.1082 }, { 0.6199, -0.6469, 1.7051 } };
static const float M_RGB2YIQ[3][3] = { { 0.299, 0.596, 0.212 }, { 0.587, -0.275, -0.523 }, { 0.114, -0.321, 0.311 } };
template<class S> inline void XYZ_to_xyY(S& x, S& y, S& z)
{
    S n     = (x + y + z);
    S n_inv = (val_of(n) >= 1.0e-6f ? 1.0f / n : S(0.0f));
    S X = x, Y = y;
    x = X * n_inv; y = Y * n_inv; z = Y;
}
template<class S> inline void xyY_to_XYZ(S& x, S& y, S& z)
{
    S Y   = z;
    S Y_y = (val_of(y) > 1.0e-6f ? Y / y : S(0.0f));
    S X   = Y_y * x;
    S Z   = Y_y * (1.0f - x - y);
    x = X; y = Y; z = Z;
}
// OIIO::safe_pow (fmath.h) around std::pow; dual.h:1058-1068 for duals
inline float oiio_safe_pow(float x, float y)
{
    if (y == 0.0f) return 1.0f;
    if (x == 0.0f) return 0.0f;
    if ((x < 0.0f) && (y != std::floor(y))) return 0.0f;
    float r         = std::pow(x, y);
    const float big = std::numeric_limits<float>::max();
    return r < -big ? -big : (r > big ? big : r);
}
inline float oiio_safe_log(float x) { return x <= 0.0f ? -std::numeric_limits<float>::max() : std::log(x); }
inline float cpow(float x, float y) { return oiio_safe_pow(x, y); }
inline Df cpow(const Df& u, float yv)
{
    Df v(yv);
    float powuvm1 = oiio_safe_pow(u.val, v.val - 1.0f);
    float powuv   = powuvm1 * u.val;
    float logu    = u.val > 0 ? oiio_safe_log(u.val) : 0.0f;
    return dualfunc(u, v, powuv, v.val * powuvm1, logu * powuv);
}
template<class S> inline S srgb_to_linear1(const S& x)
{
    return (val_of(x) <= 0.04045f) ? (x * (1.0f / 12.92f)) : cpow((x + 0.055f) * (1.0f / 1.055f), 2.4f);
}
template<class S> inline S linear_to_srgb1(const S& x)
{
    return (val_of(x) <= 0.0031308f) ? (12.92f * x) : (1.055f * cpow(x, 1.f / 2.4f) - 0.055f);
}

}  // namespace color_impl

// ColorSystem::set_colorspace (opcolor.cpp:131-236)
inline bool colorsystem_setup(ColorSystem& cs, const std::string& colorspace)
{
    using namespace color_impl;
    const Chroma* chroma = nullptr;
    for (const Chroma& c : k_color_systems)
        if (colorspace == c.name)
            chroma = &c;
    if (!chroma)
        return false;
    cs.colorspace = colorspace;
    V3 R(chroma->xRed, chroma->yRed, 0.0f), G(chroma->xGreen, chroma->yGreen, 0.0f), B(chroma->xBlue, chroma->yBlue, 0.0f),
        W(chroma->xWhite, chroma->yWhite, 0.0f);
    R.z = 1.0f - (R.x + R.y);
    G.z = 1.0f - (G.x + G.y);
    B.z = 1.0f - (B.x + B.y);
    W.z = 1.0f - (W.x + W.y);
    V3 r(G.y * B.z - B.y * G.z, B.x * G.z - G.x * B.z, G.x * B.y - B.x * G.y);
    V3 g(B.y * R.z - R.y * B.z, R.x * B.z - B.x * R.z, B.x * R.y - R.x * B.y);
    V3 b(R.y * G.z - G.y * R.z, G.x * R.z - R.x * G.z, R.x * G.y - G.x * R.y);
    V3 w(dot(r, W), dot(g, W), dot(b, W));
    if (W.y != 0.0f)
        w = w * (1.0f / W.y);
    r = r / w.x;
    g = g / w.y;
    b = b / w.z;
    float m[3][3] = { { r.x, g.x, b.x }, { r.y, g.y, b.y }, { r.z, g.z, b.z } };
    std::memcpy(cs.XYZ2RGB, m, sizeof(m));
    m33_inverse(cs.XYZ2RGB, cs.RGB2XYZ);
    cs.luminance_scale = V3(cs.RGB2XYZ[0][1], cs.RGB2XYZ[1][1], cs.RGB2XYZ[2][1]);
    float lum2         = (1.0f - cs.luminance_scale.x - cs.luminance_scale.y);
    if (std::fabs(lum2 - cs.luminance_scale.z) < 0.001f)
        cs.luminance_scale.z = lum2;
    float lastT = 0;
    for (int i = 0; lastT <= BB_MAX_TABLE_RANGE; ++i) {
        float T = BB_TABLE_MAP(float(i));
        lastT   = T;
        V3 rgb  = mul_m33(blackbody_XYZ(T), cs.XYZ2RGB);
        clamp_zero(rgb);
        rgb = V3(powf(rgb.x, 1.0f / BB_TABLE_YPOWER), powf(rgb.y, 1.0f / BB_TABLE_YPOWER), powf(rgb.z, 1.0f / BB_TABLE_YPOWER));
        cs.blackbody_table[i] = rgb;
    }
    return true;
}

inline float color_luminance(const ColorSystem& cs, const V3& c) { return dot(c, cs.luminance_scale); }
inline Df color_luminance(const ColorSystem& cs, const Dv& c)
{
    return Df(color_luminance(cs, c.val), color_luminance(cs, c.dx), color_luminance(cs, c.dy));
}
inline V3 color_blackbody(const ColorSystem& cs, float T)
{
    using namespace color_impl;
    if (T < BB_MAX_TABLE_RANGE) {
        if (T < BB_DRAPER)
            return V3(1.0e-6f, 0.0f, 0.0f);
        float t         = BB_TABLE_UNMAP(T);
        int ti          = static_cast<int>(t);
        float remainder = t - ti;
        V3 rgb  = lerp(cs.blackbody_table[ti], cs.blackbody_table[ti + 1], V3(remainder));
        V3 rgb2 = rgb * rgb;
        V3 rgb4 = rgb2 * rgb2;
        return rgb4 * rgb;
    }
    V3 rgb = mul_m33(blackbody_XYZ(T), cs.XYZ2RGB);
    clamp_zero(rgb);
    return rgb;
}
inline V3 color_wavelength(const ColorSystem& cs, float lambda_nm)
{
    using namespace color_impl;
    V3 XYZ(0.0f);
    float ii = (lambda_nm - 380.0f) / 5.0f;
    int i    = (int)ii;
    if (!((i < 0) | (i >= 80))) {
        float remainder = ii - i;
        XYZ = lerp(V3(cie_xbar[i], cie_ybar[i], cie_zbar[i]), V3(cie_xbar[i + 1], cie_ybar[i + 1], cie_zbar[i + 1]),
                   V3(remainder));
    }
    V3 rgb = mul_m33(XYZ, cs.XYZ2RGB);
    // `rgb *= 1.0 / 2.52` on a float colour: the double constant is narrowed first
    rgb = rgb * float(1.0 / 2.52);
    clamp_zero(rgb);
    return rgb;
}

// ColorSystem::transformc (opcolor.cpp:320-376) on three scalars (float or Df)
template<class S> inline void color_transform(const ColorSystem& cs, const char* from, const char* to, S& x, S& y, S& z)
{
    using namespace color_impl;
    auto is = [](const char* a, const char* b) { return a && !std::strcmp(a, b); };
    const S ix = x, iy = y, iz = z;
    bool use_colorconfig = false;
    if (is(from, "RGB") || is(from, "rgb") || is(from, "linear") || cs.colorspace == (from ? from : "")) {
    } else if (is(from, "hsv"))
        hsv_to_rgb(x, y, z);
    else if (is(from, "hsl"))
        hsl_to_rgb(x, y, z);
    else if (is(from, "YIQ"))
        mul_m33(x, y, z, M_YIQ2RGB);
    else if (is(from, "XYZ"))
        mul_m33(x, y, z, cs.XYZ2RGB);
    else if (is(from, "xyY")) {
        xyY_to_XYZ(x, y, z);
        mul_m33(x, y, z, cs.XYZ2RGB);
    } else if (is(from, "sRGB")) {
        x = srgb_to_linear1(x);
        y = srgb_to_linear1(y);
        z = srgb_to_linear1(z);
    } else
        use_colorconfig = true;
    if (use_colorconfig) {
    } else if (is(to, "RGB") || is(to, "rgb") || is(to, "linear") || cs.colorspace == (to ? to : "")) {
    } else if (is(to, "hsv"))
        rgb_to_hsv(x, y, z);
    else if (is(to, "hsl"))
        rgb_to_hsl(x, y, z);
    else if (is(to, "YIQ"))
        mul_m33(x, y, z, M_RGB2YIQ);
    else if (is(to, "XYZ"))
        mul_m33(x, y, z, cs.RGB2XYZ);
    else if (is(to, "xyY")) {
        mul_m33(x, y, z, cs.RGB2XYZ);
        XYZ_to_xyY(x, y, z);
    } else if (is(to, "sRGB")) {
        x = linear_to_srgb1(x);
        y = linear_to_srgb1(y);
        z = linear_to_srgb1(z);
    } else
        use_colorconfig = true;
    if (use_colorconfig) {
        // no OpenColorIO on this path: the colour passes through unchanged
        x = ix; y = iy; z = iz;
    }
}
// ColorSystem::to_rgb (opcolor.cpp:287-305): used by color("space", ...) constructors;
// note: no "linear"/"sRGB" clauses here, unlike transformc
inline V3 color_to_rgb(const ColorSystem& cs, const char* from, const V3& C)
{
    auto is = [](const char* a, const char* b) { return a && !std::strcmp(a, b); };
    if (is(from, "linear") || is(from, "sRGB"))
        return C;
    float x = C.x, y = C.y, z = C.z;
    color_transform(cs, from, "rgb", x, y, z);
    return V3(x, y, z);
}
inline V3 color_transformc(const ColorSystem& cs, const char* from, const char* to, const V3& C)
{
    float x = C.x, y = C.y, z = C.z;
    color_transform(cs, from, to, x, y, z);
    return V3(x, y, z);
}
inline Dv color_transformc(const ColorSystem& cs, const char* from, const char* to, const Dv& C)
{
    Df x = comp(C, 0), y = comp(C, 1), z = comp(C, 2);
    color_transform(cs, from, to, x, y, z);
    return make_dv(x, y, z);
}

}  // namespace oslo
