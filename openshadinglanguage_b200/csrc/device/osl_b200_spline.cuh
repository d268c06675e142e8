// osl_b200_spline.cuh — spline() / splineinverse() for sm_100a (product code;
// included by osl_b200_device.cuh).
//
// Replaces osl_spline_* / osl_splineinverse_* (src/liboslexec/opspline.cpp)
// and Spline::SplineInterp::evaluate / inverse (src/liboslexec/splineimpl.h:15-295).
// One template over the abscissa type X (float | Df) and knot type K
// (float | Df | V3 | Dv) covers the reference's eight type-code variants.
#pragma once

namespace osld {

enum { SPL_CATMULLROM, SPL_BEZIER, SPL_BSPLINE, SPL_HERMITE, SPL_LINEAR, SPL_CONSTANT };

// basis matrices as constant-folded selects (no constant-bank table walk under divergence)
OSLD float spline_coeff(int type, int k, int j)
{
    const float CR[16] = { -0.5f, 1.5f, -1.5f, 0.5f, 1.0f, -2.5f, 2.0f, -0.5f, -0.5f, 0.0f, 0.5f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f };
    const float BZ[16] = { -1, 3, -3, 1, 3, -6, 3, 0, -3, 3, 0, 0, 1, 0, 0, 0 };
    const float BS[16] = { (-1.0f / 6.0f), (3.0f / 6.0f), (-3.0f / 6.0f), (1.0f / 6.0f), (3.0f / 6.0f), (-6.0f / 6.0f),
                           (3.0f / 6.0f),  (0.0f / 6.0f), (-3.0f / 6.0f), (0.0f / 6.0f), (3.0f / 6.0f), (0.0f / 6.0f),
                           (1.0f / 6.0f),  (4.0f / 6.0f), (1.0f / 6.0f),  (0.0f / 6.0f) };
    const float HM[16] = { 2, 1, -2, 1, -3, -2, 3, -1, 0, 1, 0, 0, 1, 0, 0, 0 };
    const float LN[16] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, -1, 1, 0, 0, 1, 0, 0 };
    int i = 4 * k + j;
    switch (type) {
    case SPL_CATMULLROM: return CR[i];
    case SPL_BEZIER: return BZ[i];
    case SPL_BSPLINE: return BS[i];
    case SPL_HERMITE: return HM[i];
    case SPL_LINEAR: return LN[i];
    default: return 0.0f;
    }
}
OSLD int spline_step(int type) { return type == SPL_BEZIER ? 3 : (type == SPL_HERMITE ? 2 : 1); }

// Dual2<Vec3> arithmetic needed by the polynomial evaluation
OSLD Dv operator+(const Dv& a, const Dv& b) { return mkdv(a.val + b.val, a.dx + b.dx, a.dy + b.dy); }
OSLD Dv operator+(const Dv& a, V3 b) { return mkdv(a.val + b, a.dx, a.dy); }
OSLD Dv operator*(float b, const Dv& a) { return mkdv(a.val * b, a.dx * b, a.dy * b); }
OSLD Dv operator*(const Dv& a, float b) { return mkdv(a.val * b, a.dx * b, a.dy * b); }
OSLD Dv operator*(V3 a, Df b) { return mkdv(a * b.val, a * b.dx, a * b.dy); }
OSLD Dv operator*(const Dv& a, Df b)
{
    return mkdv(a.val * b.val, a.val * b.dx + a.dx * b.val, a.val * b.dy + a.dy * b.val);
}
OSLD float sclamp01(float a) { return (a >= 0.0f) ? ((a <= 1.0f) ? a : 1.0f) : 0.0f; }
OSLD Df sclamp01(Df a) { return (a.val >= 0.0f) ? ((a.val <= 1.0f) ? a : mkd(1.0f)) : mkd(0.0f); }

template<class R, class X, class K>
OSLD void spline_eval(R& result, X xval, const K* knots, int knot_count, int type)
{
    const int step = spline_step(type);
    X x        = sclamp01(xval);
    int nsegs  = ((knot_count - 4) / step) + 1;
    x          = x * (float)nsegs;
    float segx = nd(x);
    int segnum = (int)segx;
    if (segnum < 0) segnum = 0;
    if (segnum > (nsegs - 1)) segnum = nsegs - 1;
    if (type == SPL_CONSTANT) {
        assign(result, nd(knots[segnum + 1]));
        return;
    }
    x     = x - (float)segnum;
    int s = segnum * step;
    K P0 = knots[s], P1 = knots[s + 1], P2 = knots[s + 2], P3 = knots[s + 3];
    K tk[4];
#pragma unroll
    for (int k = 0; k < 4; k++)
        tk[k] = spline_coeff(type, k, 0) * P0 + spline_coeff(type, k, 1) * P1 + spline_coeff(type, k, 2) * P2
                + spline_coeff(type, k, 3) * P3;
    auto t1 = (tk[0] * x + tk[1]);
    auto t2 = (t1 * x + tk[2]);
    auto t3 = (t2 * x + tk[3]);
    assign(result, t3);
}

// OIIO::invert (regula falsi, then bisection) over one spline segment range
OSLD float spline_inverse(float y, const float* knots, int knot_count, int type)
{
    const int step = spline_step(type);
    int lowindex   = step == 1 ? 1 : 0;
    int highindex  = step == 1 ? knot_count - 2 : knot_count - 1;
    bool incr      = knots[1] < knots[knot_count - 2];
    if (incr) {
        if (y <= knots[lowindex]) return 0.0f;
        if (y >= knots[highindex]) return 1.0f;
    } else {
        if (y >= knots[lowindex]) return 0.0f;
        if (y <= knots[highindex]) return 1.0f;
    }
    int nsegs     = (knot_count - 4) / step + 1;
    float nseginv = 1.0f / (float)nsegs;
    float r0 = 0.0f, x = 0.0f;
    for (int sg = 0; sg < nsegs; ++sg) {
        float r1 = nseginv * (float)(sg + 1);
        float xmin = r0, xmax = r1, v0, v1;
        spline_eval(v0, xmin, knots, knot_count, type);
        spline_eval(v1, xmax, knots, knot_count, type);
        x = xmin;
        float v = v0;
        bool increasing = (v0 < v1);
        float vmin = increasing ? v0 : v1, vmax = increasing ? v1 : v0;
        bool bracketed = (y >= vmin && y <= vmax);
        if (bracketed) {
            if (fabsf(v0 - v1) < 1.0e-6f)
                return x;
            for (int it = 0; it < 32; ++it) {
                float t;
                if (it < 24) {
                    t = (y - v0) / (v1 - v0);
                    if (t <= 0.0f || t >= 1.0f)
                        t = 0.5f;
                } else
                    t = 0.5f;
                x = xmin * (1.0f - t) + xmax * t;
                spline_eval(v, x, knots, knot_count, type);
                if ((v < y) == increasing) {
                    xmin = x;
                    v0   = v;
                } else {
                    xmax = x;
                    v1   = v;
                }
                if (fabsf(xmax - xmin) < 1.0e-6f || fabsf(v - y) < 1.0e-6f)
                    return x;
            }
            return x;
        }
        x  = ((y < vmin) == increasing) ? xmin : xmax;
        r0 = r1;
    }
    return x;
}

// osl_splineinverse_dfdff: the same search on Dual2<float> - comparisons look at the values,
// the arithmetic carries the derivatives of y through the regula falsi steps
OSLD Df spline_inverse(Df y, const float* knots, int knot_count, int type)
{
    const int step = spline_step(type);
    int lowindex   = step == 1 ? 1 : 0;
    int highindex  = step == 1 ? knot_count - 2 : knot_count - 1;
    bool incr      = knots[1] < knots[knot_count - 2];
    if (incr) {
        if (y.val <= knots[lowindex]) return mkd(0.0f);
        if (y.val >= knots[highindex]) return mkd(1.0f);
    } else {
        if (y.val >= knots[lowindex]) return mkd(0.0f);
        if (y.val <= knots[highindex]) return mkd(1.0f);
    }
    int nsegs     = (knot_count - 4) / step + 1;
    float nseginv = 1.0f / (float)nsegs;
    Df r0 = mkd(0.0f), x = mkd(0.0f);
    for (int sg = 0; sg < nsegs; ++sg) {
        Df r1 = mkd(nseginv * (float)(sg + 1));
        Df xmin = r0, xmax = r1, v0, v1;
        spline_eval(v0, xmin, knots, knot_count, type);
        spline_eval(v1, xmax, knots, knot_count, type);
        x = xmin;
        Df v = v0;
        bool increasing = (v0.val < v1.val);
        Df vmin = increasing ? v0 : v1, vmax = increasing ? v1 : v0;
        bool bracketed = (y.val >= vmin.val && y.val <= vmax.val);
        if (bracketed) {
            if (fabsf((v0 - v1).val) < 1.0e-6f)
                return x;
            for (int it = 0; it < 32; ++it) {
                Df t;
                if (it < 24) {
                    t = (y - v0) / (v1 - v0);
                    if (t.val <= 0.0f || t.val >= 1.0f)
                        t = mkd(0.5f);
                } else
                    t = mkd(0.5f);
                x = xmin * (mkd(1.0f) - t) + xmax * t;
                spline_eval(v, x, knots, knot_count, type);
                if ((v.val < y.val) == increasing) {
                    xmin = x;
                    v0   = v;
                } else {
                    xmax = x;
                    v1   = v;
                }
                if (fabsf((xmax - xmin).val) < 1.0e-6f || fabsf((v - y).val) < 1.0e-6f)
                    return x;
            }
            return x;
        }
        x  = ((y.val < vmin.val) == increasing) ? xmin : xmax;
        r0 = r1;
    }
    return x;
}

}  // namespace osld
