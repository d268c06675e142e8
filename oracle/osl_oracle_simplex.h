// osl_oracle_simplex.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Scalar restatement of simplex noise with analytic derivatives:
//   src/liboslnoise/simplexnoise.cpp:58-760 (scramble, gradient LUTs,
//   simplexnoise1..4) and the SimplexNoise / USimplexNoise functors,
//   src/include/OSL/oslnoise.h:2765-3250 (seed 0,1,2 per colour channel,
//   chain rule onto Dual2 inputs, unsigned = 0.5*(n+1) with halved gradients).
// Gradient tables are the published Gustavson simplex tables the reference uses.
#pragma once
#include "osl_oracle.h"

namespace oslo {

inline uint32_t scramble(uint32_t v0, uint32_t v1 = 0, uint32_t v2 = 0)
{
    return bjfinal(v0, v1, v2 ^ 0xdeadbeefu);
}

static const float s_zero[4]      = { 0, 0, 0, 0 };
static const float s_grad2[8][2]  = { { -1, -1 }, { 1, 0 }, { -1, 0 }, { 1, 1 },
                                      { -1, 1 },  { 0, -1 }, { 0, 1 }, { 1, -1 } };
static const float s_grad3[16][3] = { { 1, 0, 1 },  { 0, 1, 1 },   { -1, 0, 1 },  { 0, -1, 1 },
                                      { 1, 0, -1 }, { 0, 1, -1 },  { -1, 0, -1 }, { 0, -1, -1 },
                                      { 1, -1, 0 }, { 1, 1, 0 },   { -1, 1, 0 },  { -1, -1, 0 },
                                      { 1, 0, 1 },  { -1, 0, 1 },  { 0, 1, -1 },  { 0, -1, -1 } };
static const float s_grad4[32][4]
    = { { 0, 1, 1, 1 },   { 0, 1, 1, -1 },   { 0, 1, -1, 1 },   { 0, 1, -1, -1 },  { 0, -1, 1, 1 },
        { 0, -1, 1, -1 }, { 0, -1, -1, 1 },  { 0, -1, -1, -1 }, { 1, 0, 1, 1 },    { 1, 0, 1, -1 },
        { 1, 0, -1, 1 },  { 1, 0, -1, -1 },  { -1, 0, 1, 1 },   { -1, 0, 1, -1 },  { -1, 0, -1, 1 },
        { -1, 0, -1, -1 }, { 1, 1, 0, 1 },   { 1, 1, 0, -1 },   { 1, -1, 0, 1 },   { 1, -1, 0, -1 },
        { -1, 1, 0, 1 },  { -1, 1, 0, -1 },  { -1, -1, 0, 1 },  { -1, -1, 0, -1 }, { 1, 1, 1, 0 },
        { 1, 1, -1, 0 },  { 1, -1, 1, 0 },   { 1, -1, -1, 0 },  { -1, 1, 1, 0 },   { -1, 1, -1, 0 },
        { -1, -1, 1, 0 }, { -1, -1, -1, 0 } };
static const unsigned char s_simplex[64][4]
    = { { 0, 1, 2, 3 }, { 0, 1, 3, 2 }, { 0, 0, 0, 0 }, { 0, 2, 3, 1 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 0, 0, 0, 0 }, { 1, 2, 3, 0 }, { 0, 2, 1, 3 }, { 0, 0, 0, 0 }, { 0, 3, 1, 2 }, { 0, 3, 2, 1 },
        { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 1, 3, 2, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 1, 2, 0, 3 }, { 0, 0, 0, 0 }, { 1, 3, 0, 2 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 2, 3, 0, 1 }, { 2, 3, 1, 0 }, { 1, 0, 2, 3 }, { 1, 0, 3, 2 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 0, 0, 0, 0 }, { 2, 0, 3, 1 }, { 0, 0, 0, 0 }, { 2, 1, 3, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 2, 0, 1, 3 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 3, 0, 1, 2 }, { 3, 0, 2, 1 },
        { 0, 0, 0, 0 }, { 3, 1, 2, 0 }, { 2, 1, 0, 3 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 }, { 0, 0, 0, 0 },
        { 3, 1, 0, 2 }, { 0, 0, 0, 0 }, { 3, 2, 0, 1 }, { 3, 2, 1, 0 } };

inline float sgrad1(int i, int seed)
{
    int h   = (int)scramble((uint32_t)i, (uint32_t)seed);
    float g = 1.0f + (h & 7);
    if (h & 8)
        g = -g;
    return g;
}
inline const float* sgradN(int D, const int* c, int seed)
{
    if (D == 2)
        return s_grad2[scramble((uint32_t)c[0], (uint32_t)c[1], (uint32_t)seed) & 7];
    if (D == 3)
        return s_grad3[scramble((uint32_t)c[0], (uint32_t)c[1],
                                scramble((uint32_t)c[2], (uint32_t)seed)) & 15];
    return s_grad4[scramble((uint32_t)c[0], (uint32_t)c[1],
                            scramble((uint32_t)c[2], (uint32_t)c[3], (uint32_t)seed)) & 31];
}

inline float simplexnoise1(float x, int seed, float* dn)
{
    int i0 = ifloor(x), i1 = i0 + 1;
    float x0 = x - i0, x1 = x0 - 1.0f;
    float x20 = x0 * x0, t0 = 1.0f - x20, t20 = t0 * t0, t40 = t20 * t20;
    float gx0 = sgrad1(i0, seed);
    float n0  = t40 * gx0 * x0;
    float x21 = x1 * x1, t1 = 1.0f - x21, t21 = t1 * t1, t41 = t21 * t21;
    float gx1 = sgrad1(i1, seed);
    float n1  = t41 * gx1 * x1;
    const float scale = 0.36f;
    if (dn) {
        float d = t20 * t0 * gx0 * x20;
        d += t21 * t1 * gx1 * x21;
        d *= -8.0f;
        d += t40 * gx0 + t41 * gx1;
        d *= scale;
        *dn = d;
    }
    return scale * (n0 + n1);
}

// D = 2, 3, 4.  dn: D floats or nullptr.
template<int D> inline float simplexnoiseN(const float* x, int seed, float* dn)
{
    const float F = D == 2 ? 0.366025403f : (D == 3 ? 0.333333333f : 0.309016994f);
    const float G = D == 2 ? 0.211324865f : (D == 3 ? 0.166666667f : 0.138196601f);
    const float scale = D == 2 ? 64.0f : (D == 3 ? 68.0f : 54.0f);
    float sum = x[0];
    for (int d = 1; d < D; ++d)
        sum = sum + x[d];
    float s = sum * F;
    int ic[4];
    int isum = 0;
    for (int d = 0; d < D; ++d) {
        ic[d] = ifloor(x[d] + s);
        isum += ic[d];
    }
    float t = (float)isum * G;
    float xc[5][4];  // corner-relative coordinates
    for (int d = 0; d < D; ++d) {
        float X0 = ic[d] - t;
        xc[0][d] = x[d] - X0;
    }
    int off[5][4] = {};  // integer offsets of the corners
    const float* p = xc[0];
    if (D == 2) {
        if (p[0] > p[1]) { off[1][0] = 1; off[1][1] = 0; }
        else { off[1][0] = 0; off[1][1] = 1; }
    } else if (D == 3) {
        int i1, j1, k1, i2, j2, k2;
        float x0 = p[0], y0 = p[1], z0 = p[2];
        if (x0 >= y0) {
            if (y0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
            else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
            else { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
        } else {
            if (y0 < z0) { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
            else if (x0 < z0) { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
            else { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        }
        off[1][0] = i1; off[1][1] = j1; off[1][2] = k1;
        off[2][0] = i2; off[2][1] = j2; off[2][2] = k2;
    } else {
        float x0 = p[0], y0 = p[1], z0 = p[2], w0 = p[3];
        int c = ((x0 > y0) ? 32 : 0) | ((x0 > z0) ? 16 : 0) | ((y0 > z0) ? 8 : 0)
                | ((x0 > w0) ? 4 : 0) | ((y0 > w0) ? 2 : 0) | ((z0 > w0) ? 1 : 0);
        for (int d = 0; d < 4; ++d) {
            off[1][d] = s_simplex[c][d] >= 3 ? 1 : 0;
            off[2][d] = s_simplex[c][d] >= 2 ? 1 : 0;
            off[3][d] = s_simplex[c][d] >= 1 ? 1 : 0;
        }
    }
    for (int d = 0; d < D; ++d)
        off[D][d] = 1;
    for (int c = 1; c < D; ++c)
        for (int d = 0; d < D; ++d)
            xc[c][d] = xc[0][d] - off[c][d] + (c == 1 ? G : (float)c * G);
    for (int d = 0; d < D; ++d)
        xc[D][d] = xc[0][d] - 1.0f + (float)D * G;
    float tt[5], t2[5], t4[5], n[5], dots[5];
    const float* g[5];
    for (int c = 0; c <= D; ++c) {
        float tc = 0.5f;
        for (int d = 0; d < D; ++d)
            tc = tc - xc[c][d] * xc[c][d];
        tt[c] = tc;
        g[c]  = s_zero;
        t2[c] = t4[c] = n[c] = 0.0f;
        if (tc >= 0.0f) {
            int corner[4];
            for (int d = 0; d < D; ++d)
                corner[d] = ic[d] + off[c][d];
            g[c]  = sgradN(D, corner, seed);
            t2[c] = tc * tc;
            t4[c] = t2[c] * t2[c];
        }
        float dot = g[c][0] * xc[c][0];
        for (int d = 1; d < D; ++d)
            dot = dot + g[c][d] * xc[c][d];
        dots[c] = dot;
        if (tc >= 0.0f)
            n[c] = t4[c] * dot;
    }
    float nsum = n[0];
    for (int c = 1; c <= D; ++c)
        nsum = nsum + n[c];
    if (dn) {
        for (int d = 0; d < D; ++d)
            dn[d] = 0.0f;
        for (int c = 0; c <= D; ++c) {
            float temp = t2[c] * tt[c] * dots[c];
            for (int d = 0; d < D; ++d) {
                if (c == 0)
                    dn[d] = temp * xc[c][d];
                else
                    dn[d] += temp * xc[c][d];
            }
        }
        for (int d = 0; d < D; ++d) {
            dn[d] *= -8.0f;
            float gs = t4[0] * g[0][d];
            for (int c = 1; c <= D; ++c)
                gs = gs + t4[c] * g[c][d];
            dn[d] += gs;
            dn[d] *= scale;
        }
    }
    return scale * nsum;
}

inline float simplex_eval(int dim, const float* x, int seed, float* dn)
{
    switch (dim) {
    case 1: return simplexnoise1(x[0], seed, dn);
    case 2: return simplexnoiseN<2>(x, seed, dn);
    case 3: return simplexnoiseN<3>(x, seed, dn);
    default: return simplexnoiseN<4>(x, seed, dn);
    }
}

// NC = 1|3 results; UNSIGNED: usimplex.  float inputs.
template<int NC, bool UNSIGNED> inline void simplex_nd(float* out, int dim, const float* in)
{
    for (int c = 0; c < NC; ++c) {
        float r = simplex_eval(dim, in, c, nullptr);
        out[c]  = UNSIGNED ? 0.5f * (r + 1.0f) : r;
    }
}
// Dual inputs: chain rule exactly as the functors write it
template<int NC, bool UNSIGNED> inline void simplex_nd(Df* out, int dim, const Df* in)
{
    float x[4];
    for (int d = 0; d < dim; ++d)
        x[d] = in[d].val;
    for (int c = 0; c < NC; ++c) {
        float dn[4];
        float r = simplex_eval(dim, x, c, dn);
        if (UNSIGNED) {
            r = 0.5f * (r + 1.0f);
            for (int d = 0; d < dim; ++d)
                dn[d] *= 0.5f;
        }
        float dx = dn[0] * in[0].dx, dy = dn[0] * in[0].dy;
        for (int d = 1; d < dim; ++d) {
            dx = dx + dn[d] * in[d].dx;
            dy = dy + dn[d] * in[d].dy;
        }
        out[c] = Df(r, dx, dy);
    }
}

}  // namespace oslo
