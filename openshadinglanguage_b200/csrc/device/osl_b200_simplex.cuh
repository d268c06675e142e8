// osl_b200_simplex.cuh — simplex noise 1-4D with analytic derivatives for
// sm_100a (product code; included by osl_b200_device.cuh).
//
// Replaces src/liboslnoise/simplexnoise.cpp:58-760 and the SimplexNoise /
// USimplexNoise functors (src/include/OSL/oslnoise.h:2765-3250).  Same
// operation order as the reference so strict mode (--fmad=false) is bit-exact.
// The gradient tables are 2-bit packed into immediates and decoded with
// shifts, so divergent hash values never serialise on a constant-bank lookup.
#pragma once

namespace osld {

OSLD u32 scramble(u32 v0, u32 v1 = 0u, u32 v2 = 0u) { return bjfinal(v0, v1, v2 ^ 0xdeadbeefu); }

// component tables, 2 bits per entry: 0 -> -1, 1 -> 0, 2 -> +1
OSLD float unpack2(unsigned long long tab, u32 idx) { return (float)((int)((tab >> (2u * idx)) & 3ull) - 1); }

OSLD void sgrad2(float* g, int i, int j, int seed)
{
    u32 h = scramble((u32)i, (u32)j, (u32)seed) & 7u;
    // { -1,-1 },{ 1,0 },{ -1,0 },{ 1,1 },{ -1,1 },{ 0,-1 },{ 0,1 },{ 1,-1 }
    const unsigned long long gx = 0x2ull << 2 | 0x0ull | 0x0ull << 4 | 0x2ull << 6 | 0x0ull << 8 | 0x1ull << 10 | 0x1ull << 12 | 0x2ull << 14;
    const unsigned long long gy = 0x0ull | 0x1ull << 2 | 0x1ull << 4 | 0x2ull << 6 | 0x2ull << 8 | 0x0ull << 10 | 0x2ull << 12 | 0x0ull << 14;
    g[0] = unpack2(gx, h);
    g[1] = unpack2(gy, h);
}
// encode helper for the larger tables: E(a) maps {-1,0,1} -> {0,1,2}
#define OSLD_E(a) ((unsigned long long)((a) + 1))
#define OSLD_P8(s, a, b, c, d, e, f, g, h)                                                            \
    (OSLD_E(a) << (2 * (s)) | OSLD_E(b) << (2 * (s) + 2) | OSLD_E(c) << (2 * (s) + 4) | OSLD_E(d) << (2 * (s) + 6) \
     | OSLD_E(e) << (2 * (s) + 8) | OSLD_E(f) << (2 * (s) + 10) | OSLD_E(g) << (2 * (s) + 12) | OSLD_E(h) << (2 * (s) + 14))
OSLD void sgrad3(float* g, int i, int j, int k, int seed)
{
    u32 h = scramble((u32)i, (u32)j, scramble((u32)k, (u32)seed)) & 15u;
    const unsigned long long gx = OSLD_P8(0, 1, 0, -1, 0, 1, 0, -1, 0) | OSLD_P8(8, 1, 1, -1, -1, 1, -1, 0, 0);
    const unsigned long long gy = OSLD_P8(0, 0, 1, 0, -1, 0, 1, 0, -1) | OSLD_P8(8, -1, 1, 1, -1, 0, 0, 1, -1);
    const unsigned long long gz = OSLD_P8(0, 1, 1, 1, 1, -1, -1, -1, -1) | OSLD_P8(8, 0, 0, 0, 0, 1, 1, -1, -1);
    g[0] = unpack2(gx, h);
    g[1] = unpack2(gy, h);
    g[2] = unpack2(gz, h);
}
OSLD void sgrad4(float* g, int i, int j, int k, int l, int seed)
{
    u32 h = scramble((u32)i, (u32)j, scramble((u32)k, (u32)l, (u32)seed)) & 31u;
    const unsigned long long gx = OSLD_P8(0, 0, 0, 0, 0, 0, 0, 0, 0) | OSLD_P8(8, 1, 1, 1, 1, -1, -1, -1, -1)
                                  | OSLD_P8(16, 1, 1, 1, 1, -1, -1, -1, -1) | OSLD_P8(24, 1, 1, 1, 1, -1, -1, -1, -1);
    const unsigned long long gy = OSLD_P8(0, 1, 1, 1, 1, -1, -1, -1, -1) | OSLD_P8(8, 0, 0, 0, 0, 0, 0, 0, 0)
                                  | OSLD_P8(16, 1, 1, -1, -1, 1, 1, -1, -1) | OSLD_P8(24, 1, 1, -1, -1, 1, 1, -1, -1);
    const unsigned long long gz = OSLD_P8(0, 1, 1, -1, -1, 1, 1, -1, -1) | OSLD_P8(8, 1, 1, -1, -1, 1, 1, -1, -1)
                                  | OSLD_P8(16, 0, 0, 0, 0, 0, 0, 0, 0) | OSLD_P8(24, 1, -1, 1, -1, 1, -1, 1, -1);
    const unsigned long long gw = OSLD_P8(0, 1, -1, 1, -1, 1, -1, 1, -1) | OSLD_P8(8, 1, -1, 1, -1, 1, -1, 1, -1)
                                  | OSLD_P8(16, 1, -1, 1, -1, 1, -1, 1, -1) | OSLD_P8(24, 0, 0, 0, 0, 0, 0, 0, 0);
    g[0] = unpack2(gx, h);
    g[1] = unpack2(gy, h);
    g[2] = unpack2(gz, h);
    g[3] = unpack2(gw, h);
}
OSLD float sgrad1(int i, int seed)
{
    u32 h   = scramble((u32)i, (u32)seed);
    float g = 1.0f + (float)(h & 7u);
    return (h & 8u) ? -g : g;
}

OSLD float simplex1(float x, int seed, float* dn)
{
    int i0;
    float x0  = ffrac(x, &i0);
    float x1  = x0 - 1.0f;
    float x20 = x0 * x0, t0 = 1.0f - x20, t20 = t0 * t0, t40 = t20 * t20;
    float gx0 = sgrad1(i0, seed);
    float n0  = t40 * gx0 * x0;
    float x21 = x1 * x1, t1 = 1.0f - x21, t21 = t1 * t1, t41 = t21 * t21;
    float gx1 = sgrad1(i0 + 1, seed);
    float n1  = t41 * gx1 * x1;
    if (dn) {
        float d = t20 * t0 * gx0 * x20;
        d += t21 * t1 * gx1 * x21;
        d *= -8.0f;
        d += t40 * gx0 + t41 * gx1;
        d *= 0.36f;
        *dn = d;
    }
    return 0.36f * (n0 + n1);
}

// rank of coordinate d among the 4 (the reference's simplex[64][4] table):
// rank = number of other coordinates that are strictly smaller, with the
// table's tie-breaking (earlier coordinate wins ties).
template<int D, bool DERIV> OSLD float simplexN(const float* x, int seed, float* dn)
{
    const float F     = D == 2 ? 0.366025403f : (D == 3 ? 0.333333333f : 0.309016994f);
    const float G     = D == 2 ? 0.211324865f : (D == 3 ? 0.166666667f : 0.138196601f);
    const float scale = D == 2 ? 64.0f : (D == 3 ? 68.0f : 54.0f);
    float sum = x[0];
#pragma unroll
    for (int d = 1; d < D; ++d)
        sum = sum + x[d];
    float s = sum * F;
    int ic[D];
    int isum = 0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        ic[d] = ifloor(x[d] + s);
        isum += ic[d];
    }
    float t = (float)isum * G;
    float xc[D + 1][D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
        float X0 = (float)ic[d] - t;
        xc[0][d] = x[d] - X0;
    }
    int off[D + 1][D];
#pragma unroll
    for (int c = 0; c <= D; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d)
            off[c][d] = (c == D) ? 1 : 0;
    if (D == 2) {
        bool a    = xc[0][0] > xc[0][1];
        off[1][0] = a ? 1 : 0;
        off[1][1] = a ? 0 : 1;
    } else if (D == 3) {
        float x0 = xc[0][0], y0 = xc[0][1], z0 = xc[0][2 % D];
        int i1, j1, k1, i2, j2, k2;
        if (x0 >= y0) {
            if (y0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
            else if (x0 >= z0) { i1 = 1; j1 = 0; k1 = 0; i2 = 1; j2 = 0; k2 = 1; }
            else { i1 = 0; j1 = 0; k1 = 1; i2 = 1; j2 = 0; k2 = 1; }
        } else {
            if (y0 < z0) { i1 = 0; j1 = 0; k1 = 1; i2 = 0; j2 = 1; k2 = 1; }
            else if (x0 < z0) { i1 = 0; j1 = 1; k1 = 0; i2 = 0; j2 = 1; k2 = 1; }
            else { i1 = 0; j1 = 1; k1 = 0; i2 = 1; j2 = 1; k2 = 0; }
        }
        off[1][0] = i1; off[1][1] = j1; off[1][2 % D] = k1;
        off[2][0] = i2; off[2][1] = j2; off[2][2 % D] = k2;
    } else {
        // the simplex[c] row is the rank vector of (x0,y0,z0,w0) under the six
        // strict comparisons the reference makes; compute the ranks directly
        float x0 = xc[0][0], y0 = xc[0][1], z0 = xc[0][2 % D], w0 = xc[0][3 % D];
        int c1 = x0 > y0, c2 = x0 > z0, c3 = y0 > z0, c4 = x0 > w0, c5 = y0 > w0, c6 = z0 > w0;
        int rk[4];
        rk[0] = c1 + c2 + c4;
        rk[1] = (1 - c1) + c3 + c5;
        rk[2] = (1 - c2) + (1 - c3) + c6;
        rk[3] = (1 - c4) + (1 - c5) + (1 - c6);
#pragma unroll
        for (int d = 0; d < D; ++d) {
            off[1][d] = rk[d] >= 3 ? 1 : 0;
            off[2][d] = rk[d] >= 2 ? 1 : 0;
            off[3 % (D + 1)][d] = rk[d] >= 1 ? 1 : 0;
        }
    }
#pragma unroll
    for (int c = 1; c < D; ++c)
#pragma unroll
        for (int d = 0; d < D; ++d)
            xc[c][d] = xc[0][d] - (float)off[c][d] + (c == 1 ? G : (float)c * G);
#pragma unroll
    for (int d = 0; d < D; ++d)
        xc[D][d] = xc[0][d] - 1.0f + (float)D * G;
    float tt[D + 1], t2[D + 1], t4[D + 1], n[D + 1], dots[D + 1], g[D + 1][D];
#pragma unroll
    for (int c = 0; c <= D; ++c) {
        float tc = 0.5f;
#pragma unroll
        for (int d = 0; d < D; ++d)
            tc = tc - xc[c][d] * xc[c][d];
        tt[c] = tc;
#pragma unroll
        for (int d = 0; d < D; ++d)
            g[c][d] = 0.0f;
        t2[c] = t4[c] = n[c] = 0.0f;
        if (tc >= 0.0f) {
            if (D == 2) sgrad2(g[c], ic[0] + off[c][0], ic[1] + off[c][1], seed);
            else if (D == 3) sgrad3(g[c], ic[0] + off[c][0], ic[1] + off[c][1], ic[2 % D] + off[c][2 % D], seed);
            else sgrad4(g[c], ic[0] + off[c][0], ic[1] + off[c][1], ic[2 % D] + off[c][2 % D], ic[3 % D] + off[c][3 % D], seed);
            t2[c] = tc * tc;
            t4[c] = t2[c] * t2[c];
        }
        float dot = g[c][0] * xc[c][0];
#pragma unroll
        for (int d = 1; d < D; ++d)
            dot = dot + g[c][d] * xc[c][d];
        dots[c] = dot;
        if (tc >= 0.0f)
            n[c] = t4[c] * dot;
    }
    float nsum = n[0];
#pragma unroll
    for (int c = 1; c <= D; ++c)
        nsum = nsum + n[c];
    if (DERIV) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            float acc = 0.0f;
#pragma unroll
            for (int c = 0; c <= D; ++c) {
                float temp = t2[c] * tt[c] * dots[c];
                acc        = (c == 0) ? temp * xc[c][d] : acc + temp * xc[c][d];
            }
            acc *= -8.0f;
            float gs = t4[0] * g[0][d];
#pragma unroll
            for (int c = 1; c <= D; ++c)
                gs = gs + t4[c] * g[c][d];
            acc += gs;
            acc *= scale;
            dn[d] = acc;
        }
    }
    return scale * nsum;
}

template<int DIM, bool DERIV> OSLD float simplex_eval(const float* x, int seed, float* dn)
{
    if (DIM == 1)
        return simplex1(x[0], seed, DERIV ? dn : nullptr);
    return simplexN<(DIM < 2 ? 2 : DIM), DERIV>(x, seed, dn);
}

template<int DIM, int NC, bool UNSIGNED> OSLD void simplex(float* out, const float* in)
{
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float r = simplex_eval<DIM, false>(in, c, nullptr);
        out[c]  = UNSIGNED ? 0.5f * (r + 1.0f) : r;
    }
}
template<int DIM, int NC, bool UNSIGNED> OSLD void simplex(Df* out, const Df* in)
{
    float x[4];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
        x[d] = in[d].val;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        float dn[4];
        float r = simplex_eval<DIM, true>(x, c, dn);
        if (UNSIGNED) {
            r = 0.5f * (r + 1.0f);
#pragma unroll
            for (int d = 0; d < DIM; ++d)
                dn[d] *= 0.5f;
        }
        float dx = dn[0] * in[0].dx, dy = dn[0] * in[0].dy;
#pragma unroll
        for (int d = 1; d < DIM; ++d) {
            dx = dx + dn[d] * in[d].dx;
            dy = dy + dn[d] * in[d].dy;
        }
        out[c] = mkd(r, dx, dy);
    }
}

}  // namespace osld
