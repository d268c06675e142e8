// osl_b200_color.cpp — host side of the colour shadeops (product code).
//
// Builds the uniform colour state the reference keeps in
// ShadingStateUniform::m_colorsystem (ColorSystem::set_colorspace,
// src/liboslexec/opcolor.cpp:131-236): XYZ<->RGB matrices of the working
// space, the luminance weights and the 317-entry blackbody table
// (opcolor_impl.h:134-160).  On this back end the state is not a run-time
// object: the code generator bakes it into the generated CUDA module as a
// constant array (layout below), so matrix entries fold into immediates and
// no per-launch upload or pointer is needed.
#include "osl_b200_group.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../device/osl_b200_cie1931.cuh"

namespace oslb200 {

namespace {

// OIIO fast_exp2 / fast_exp / fast_expm1 (fmath.h), host restatement: the table
// must hold exactly what the reference's host computes
inline float
h_fast_exp2(float x)
{
    if (x < -126.0f)
        x = -126.0f;
    if (x > 126.0f)
        x = 126.0f;
    int m = (int)x;
    x -= (float)m;
    x       = 1.0f - (1.0f - x);
    float r = 1.33336498402e-3f;
    r       = x * r + 9.810352697968e-3f;
    r       = x * r + 5.551834031939e-2f;
    r       = x * r + 0.2401793301105f;
    r       = x * r + 0.693144857883f;
    r       = x * r + 1.0f;
    uint32_t b;
    std::memcpy(&b, &r, 4);
    b += (uint32_t)m << 23;
    std::memcpy(&r, &b, 4);
    return r;
}
inline float
h_fast_expm1(float x)
{
    if (std::fabs(x) < 0.03f) {
        float y = 1.0f - (1.0f - x);
        return std::copysign(0.5f * (y * y) + y, x);
    }
    return h_fast_exp2(x * (float)(1.0 / M_LN2)) - 1.0f;
}

struct C3 {
    float x, y, z;
};

// Planck spectrum integrated against the CIE observer (bb_spectrum, spectrum_to_XYZ)
C3
blackbody_XYZ(float temp)
{
    float X = 0, Y = 0, Z = 0;
    const float dlambda = 5.0f * 1e-9;
    for (int i = 0; i < 81; ++i) {
        float lambda   = 380.0f + 5.0f * i;
        float wlm      = lambda * 1e-9f;
        const float c1 = 3.74183e-16f, c2 = 1.4388e-2f;
        float wlm2 = wlm * wlm, wlm4 = wlm2 * wlm2, wlm5 = wlm4 * wlm;
        float inv5 = 1.0f / wlm5;
        float Me   = float((c1 * inv5) / h_fast_expm1(c2 / (wlm * temp))) * dlambda;
        X += Me * cie_xbar[i];
        Y += Me * cie_ybar[i];
        Z += Me * cie_zbar[i];
    }
    return { X, Y, Z };
}

C3
mul33(C3 v, const float M[3][3])
{
    return { v.x * M[0][0] + v.y * M[1][0] + v.z * M[2][0], v.x * M[0][1] + v.y * M[1][1] + v.z * M[2][1],
             v.x * M[0][2] + v.y * M[1][2] + v.z * M[2][2] };
}

struct Chroma {
    const char* name;
    float xr, yr, xg, yg, xb, yb, xw, yw;
};
// chromaticities of the primaries and white points (opcolor.cpp:26-50; published
// values of the respective standards)
const Chroma systems[] = {
    { "Rec709", 0.64, 0.33, 0.30, 0.60, 0.15, 0.06, 0.3127, 0.3291 },
    { "sRGB", 0.64, 0.33, 0.30, 0.60, 0.15, 0.06, 0.3127, 0.3291 },
    { "NTSC", 0.67, 0.33, 0.21, 0.71, 0.14, 0.08, 0.3101, 0.3162 },
    { "EBU", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, 0.3127, 0.3291 },
    { "PAL", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, 0.3127, 0.3291 },
    { "SECAM", 0.64, 0.33, 0.29, 0.60, 0.15, 0.06, 0.3127, 0.3291 },
    { "SMPTE", 0.630, 0.340, 0.310, 0.595, 0.155, 0.070, 0.3127, 0.3291 },
    { "HDTV", 0.670, 0.330, 0.210, 0.710, 0.150, 0.060, 0.3127, 0.3291 },
    { "CIE", 0.7355, 0.2645, 0.2658, 0.7243, 0.1669, 0.0085, 0.33333333, 0.33333333 },
    { "AdobeRGB", 0.64, 0.33, 0.21, 0.71, 0.15, 0.06, 0.3127, 0.3291 },
    { "XYZ", 1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.33333333, 0.33333333 },
    { "ACES2065-1", 0.7347, 0.2653, 0.0, 1.0, 0.0001, -0.077, 0.32168, 0.33767 },
    { "ACEScg", 0.713, 0.293, 0.165, 0.83, 0.128, 0.044, 0.32168, 0.33767 },
};

}  // namespace

// Layout (floats): [0..8] XYZ2RGB row-major, [9..17] RGB2XYZ, [18..20] luminance
// weights, [21 ..] blackbody table, 317 RGB entries holding value^(1/5).
bool
colorsystem_table(const std::string& colorspace, std::vector<float>& out)
{
    const Chroma* c = nullptr;
    for (const Chroma& s : systems)
        if (colorspace == s.name)
            c = &s;
    if (!c)
        return false;
    C3 R { c->xr, c->yr, 0 }, G { c->xg, c->yg, 0 }, B { c->xb, c->yb, 0 }, W { c->xw, c->yw, 0 };
    R.z = 1.0f - (R.x + R.y);
    G.z = 1.0f - (G.x + G.y);
    B.z = 1.0f - (B.x + B.y);
    W.z = 1.0f - (W.x + W.y);
    C3 r { G.y * B.z - B.y * G.z, B.x * G.z - G.x * B.z, G.x * B.y - B.x * G.y };
    C3 g { B.y * R.z - R.y * B.z, R.x * B.z - B.x * R.z, B.x * R.y - R.x * B.y };
    C3 b { R.y * G.z - G.y * R.z, G.x * R.z - R.x * G.z, R.x * G.y - G.x * R.y };
    auto dot = [](C3 a, C3 q) { return a.x * q.x + a.y * q.y + a.z * q.z; };
    C3 w { dot(r, W), dot(g, W), dot(b, W) };
    if (W.y != 0.0f) {
        float s = 1.0f / W.y;
        w       = { w.x * s, w.y * s, w.z * s };
    }
    r = { r.x / w.x, r.y / w.x, r.z / w.x };
    g = { g.x / w.y, g.y / w.y, g.z / w.y };
    b = { b.x / w.z, b.y / w.z, b.z / w.z };
    float x[3][3] = { { r.x, g.x, b.x }, { r.y, g.y, b.y }, { r.z, g.z, b.z } };
    // Imath Matrix33::inverse(): cofactors over the determinant, identity when singular
    float s[3][3] = {
        { x[1][1] * x[2][2] - x[2][1] * x[1][2], x[2][1] * x[0][2] - x[0][1] * x[2][2], x[0][1] * x[1][2] - x[1][1] * x[0][2] },
        { x[2][0] * x[1][2] - x[1][0] * x[2][2], x[0][0] * x[2][2] - x[2][0] * x[0][2], x[1][0] * x[0][2] - x[0][0] * x[1][2] },
        { x[1][0] * x[2][1] - x[2][0] * x[1][1], x[2][0] * x[0][1] - x[0][0] * x[2][1], x[0][0] * x[1][1] - x[1][0] * x[0][1] }
    };
    float det = x[0][0] * s[0][0] + x[0][1] * s[1][0] + x[0][2] * s[2][0];
    bool ok   = true;
    if (std::fabs(det) >= 1) {
        for (auto& row : s)
            for (float& e : row)
                e /= det;
    } else {
        float mr = std::fabs(det) / std::numeric_limits<float>::min();
        for (int i = 0; i < 3 && ok; ++i)
            for (int j = 0; j < 3 && ok; ++j) {
                if (mr > std::fabs(s[i][j]))
                    s[i][j] /= det;
                else
                    ok = false;
            }
    }
    if (!ok)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s[i][j] = i == j ? 1.0f : 0.0f;
    out.clear();
    for (auto& row : x)
        for (float e : row)
            out.push_back(e);
    for (auto& row : s)
        for (float e : row)
            out.push_back(e);
    float lum[3] = { s[0][1], s[1][1], s[2][1] };
    float lum2   = (1.0f - lum[0] - lum[1]);
    if (std::fabs(lum2 - lum[2]) < 0.001f)
        lum[2] = lum2;
    for (float e : lum)
        out.push_back(e);
    float lastT = 0;
    for (int i = 0; lastT <= 12000.0f; ++i) {
        float is = std::sqrt(float(i));
        float T  = is * is * is * 2.0f + 800.0f;
        lastT    = T;
        C3 rgb   = mul33(blackbody_XYZ(T), x);
        rgb.x    = rgb.x < 0.0f ? 0.0f : rgb.x;
        rgb.y    = rgb.y < 0.0f ? 0.0f : rgb.y;
        rgb.z    = rgb.z < 0.0f ? 0.0f : rgb.z;
        out.push_back(powf(rgb.x, 1.0f / 5.0f));
        out.push_back(powf(rgb.y, 1.0f / 5.0f));
        out.push_back(powf(rgb.z, 1.0f / 5.0f));
    }
    out.resize(21 + 3 * 317, 0.0f);
    return true;
}

// The table as a CUDA definition; %.9g round-trips every float exactly.
std::string
colorsystem_cuda_definition(const std::string& colorspace)
{
    std::vector<float> t;
    if (!colorsystem_table(colorspace, t))
        throw std::runtime_error("B200 back end: unknown colorspace \"" + colorspace + "\"");
    std::ostringstream o;
    o << "// colour system '" << colorspace << "': XYZ2RGB[9] RGB2XYZ[9] luminance[3] blackbody^(1/5)[317][3]\n";
    o << "static __device__ const float osl_cs_[" << t.size() << "] = {";
    char buf[32];
    for (size_t i = 0; i < t.size(); ++i) {
        if (i % 8 == 0)
            o << "\n    ";
        snprintf(buf, sizeof buf, "%.9g", (double)t[i]);
        std::string lit = buf;
        if (lit.find_first_of(".eEn") == std::string::npos)
            lit += ".0";
        o << lit << "f, ";
    }
    o << "\n};\n";
    return o.str();
}

}  // namespace oslb200
