// osl_b200_device.cuh — sm_100a device shadeop library (product code).
//
// This is the device runtime that generated group kernels are compiled
// against (NVRTC, --gpu-architecture=sm_100a) and that shadeops.cu exposes as
// batch kernels.  It replaces, for the GPU, the reference's shadeop runtime:
//   integer hash / cell / hash noise   src/include/OSL/oslnoise.h:229-289, 337-551
//   Perlin 1-4D, float/Vec3, +Dual2    src/include/OSL/oslnoise.h:846-1041, 1327-2280
//   noise front ends                   src/liboslexec/opnoise.cpp:71-272, 276-470
//   per-component math + duals         src/liboslexec/llvm_ops.cpp:135-700, include/OSL/dual.h
//   vector functions                   src/include/OSL/dual_vec.h:405-560
// The file is self-contained (no host headers) so the same text compiles under
// NVRTC and nvcc.  Floating-point contract: with --fmad=false every function
// here evaluates the same IEEE-754 operation sequence as the reference's
// non-SIMD (CGScalar) formulation; with --fmad=true only contraction differs.
#pragma once

#define OSLD __device__ __forceinline__
#ifndef OSLD_ENTRY_INLINE
#define OSLD_ENTRY_INLINE __forceinline__   // a material group called from the integrator
#endif

namespace osld {

typedef unsigned int u32;

struct V3 {
    float x, y, z;
};
struct Df {
    float val, dx, dy;
};
struct Dv {
    V3 val, dx, dy;
};

OSLD V3 mkv(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
OSLD V3 mkv(float a) { return mkv(a, a, a); }
OSLD Df mkd(float v) { Df r; r.val = v; r.dx = 0.0f; r.dy = 0.0f; return r; }
OSLD Df mkd(float v, float dx, float dy) { Df r; r.val = v; r.dx = dx; r.dy = dy; return r; }
OSLD Df as_dual(float a) { return mkd(a); }
OSLD Df as_dual(Df a) { return a; }
OSLD Dv mkdv(V3 v) { Dv r; r.val = v; r.dx = mkv(0.0f); r.dy = mkv(0.0f); return r; }
OSLD Dv mkdv(V3 v, V3 dx, V3 dy) { Dv r; r.val = v; r.dx = dx; r.dy = dy; return r; }

OSLD float vget(const V3& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : v.z); }
OSLD void vset(V3& v, int c, float f) { if (c == 0) v.x = f; else if (c == 1) v.y = f; else v.z = f; }

// ---- V3 arithmetic ---------------------------------------------------------
OSLD V3 operator+(V3 a, V3 b) { return mkv(a.x + b.x, a.y + b.y, a.z + b.z); }
OSLD V3 operator-(V3 a, V3 b) { return mkv(a.x - b.x, a.y - b.y, a.z - b.z); }
OSLD V3 operator*(V3 a, V3 b) { return mkv(a.x * b.x, a.y * b.y, a.z * b.z); }
OSLD V3 operator*(V3 a, float b) { return mkv(a.x * b, a.y * b, a.z * b); }
OSLD V3 operator*(float b, V3 a) { return mkv(a.x * b, a.y * b, a.z * b); }
OSLD V3 operator-(V3 a) { return mkv(-a.x, -a.y, -a.z); }
OSLD V3 cross3(V3 a, V3 b)
{
    return mkv(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// ---- Dual2<float> arithmetic (dual.h:400-600) --------------------------------
OSLD Df operator+(Df a, Df b) { return mkd(a.val + b.val, a.dx + b.dx, a.dy + b.dy); }
OSLD Df operator+(Df a, float b) { return mkd(a.val + b, a.dx, a.dy); }
OSLD Df operator+(float a, Df b) { return mkd(a + b.val, b.dx, b.dy); }
OSLD Df operator-(Df a, Df b) { return mkd(a.val - b.val, a.dx - b.dx, a.dy - b.dy); }
OSLD Df operator-(Df a, float b) { return mkd(a.val - b, a.dx, a.dy); }
OSLD Df operator-(float a, Df b) { return mkd(a - b.val, -b.dx, -b.dy); }
OSLD Df operator-(Df a) { return mkd(-a.val, -a.dx, -a.dy); }
OSLD Df operator*(Df a, Df b)
{
    return mkd(a.val * b.val, a.val * b.dx + a.dx * b.val, a.val * b.dy + a.dy * b.val);
}
OSLD Df operator*(Df a, float b) { return mkd(a.val * b, a.dx * b, a.dy * b); }
OSLD Df operator*(float b, Df a) { return mkd(a.val * b, a.dx * b, a.dy * b); }
OSLD Df operator/(Df a, Df b)
{
    float binv = 1.0f / b.val;
    float q    = a.val / b.val;
    return mkd(q, binv * (a.dx - q * b.dx), binv * (a.dy - q * b.dy));
}
OSLD Df operator/(float a, Df b)
{
    float binv = 1.0f / b.val;
    float q    = a / b.val;
    return mkd(q, binv * (-q * b.dx), binv * (-q * b.dy));
}
OSLD Dv operator-(Dv a, Dv b) { return mkdv(a.val - b.val, a.dx - b.dx, a.dy - b.dy); }

// chain rule (dual.h:767-800)
OSLD Df chain(Df u, float f, float df) { return mkd(f, df * u.dx, df * u.dy); }
OSLD Df chain(Df u, Df v, float f, float fu, float fv)
{
    return mkd(f, fu * u.dx + fv * v.dx, fu * u.dy + fv * v.dy);
}

// ---- scalar access helpers used by generated code ----------------------------
OSLD float nd(float a) { return a; }
OSLD float nd(Df a) { return a.val; }
OSLD V3 nd(V3 a) { return a; }
OSLD V3 nd(Dv a) { return a.val; }
OSLD int nd(int a) { return a; }
OSLD float getc(float a, int) { return a; }
OSLD float getc(int a, int) { return (float)a; }
OSLD Df getc(Df a, int) { return a; }
OSLD float getc(const V3& a, int c) { return vget(a, c); }
OSLD Df getc(const Dv& a, int c) { return mkd(vget(a.val, c), vget(a.dx, c), vget(a.dy, c)); }
OSLD void setc(float& d, int, float v) { d = v; }
OSLD void setc(float& d, int, Df v) { d = v.val; }
OSLD void setc(Df& d, int, float v) { d = mkd(v); }
OSLD void setc(Df& d, int, Df v) { d = v; }
OSLD void setc(V3& d, int c, float v) { vset(d, c, v); }
OSLD void setc(V3& d, int c, Df v) { vset(d, c, v.val); }
OSLD void setc(Dv& d, int c, float v) { vset(d.val, c, v); vset(d.dx, c, 0.0f); vset(d.dy, c, 0.0f); }
OSLD void setc(Dv& d, int c, Df v) { vset(d.val, c, v.val); vset(d.dx, c, v.dx); vset(d.dy, c, v.dy); }
OSLD void setc(int& d, int, int v) { d = v; }
OSLD void setc(int& d, int, float v) { d = (int)v; }

OSLD void assign(float& d, float s) { d = s; }
OSLD void assign(float& d, int s) { d = (float)s; }
OSLD void assign(float& d, Df s) { d = s.val; }
OSLD void assign(Df& d, float s) { d = mkd(s); }
OSLD void assign(Df& d, int s) { d = mkd((float)s); }
OSLD void assign(Df& d, Df s) { d = s; }
OSLD void assign(int& d, int s) { d = s; }
OSLD void assign(int& d, float s) { d = (int)s; }
OSLD void assign(int& d, Df s) { d = (int)s.val; }
OSLD void assign(V3& d, float s) { d = mkv(s); }
OSLD void assign(V3& d, int s) { d = mkv((float)s); }
OSLD void assign(V3& d, Df s) { d = mkv(s.val); }
OSLD void assign(V3& d, V3 s) { d = s; }
OSLD void assign(V3& d, const Dv& s) { d = s.val; }
OSLD void assign(Dv& d, float s) { d = mkdv(mkv(s)); }
OSLD void assign(Dv& d, int s) { d = mkdv(mkv((float)s)); }
OSLD void assign(Dv& d, Df s) { d = mkdv(mkv(s.val), mkv(s.dx), mkv(s.dy)); }
OSLD void assign(Dv& d, V3 s) { d = mkdv(s); }
OSLD void assign(Dv& d, const Dv& s) { d = s; }

// ---------------------------------------------------------------------------
// lookup3 integer hash (bit-exact contract).  Rotates map to one SHF each.
// ---------------------------------------------------------------------------
OSLD u32 rotl(u32 x, int k) { return __funnelshift_l(x, x, k); }

OSLD void bjmix(u32& a, u32& b, u32& c)
{
    a -= c; a ^= rotl(c, 4);  c += b;
    b -= a; b ^= rotl(a, 6);  a += c;
    c -= b; c ^= rotl(b, 8);  b += a;
    a -= c; a ^= rotl(c, 16); c += b;
    b -= a; b ^= rotl(a, 19); a += c;
    c -= b; c ^= rotl(b, 4);  b += a;
}
OSLD u32 bjfinal(u32 a, u32 b, u32 c)
{
    c ^= b; c -= rotl(b, 14);
    a ^= c; a -= rotl(c, 11);
    b ^= a; b -= rotl(a, 25);
    c ^= b; c -= rotl(b, 16);
    a ^= c; a -= rotl(c, 4);
    b ^= a; b -= rotl(a, 14);
    c ^= b; c -= rotl(b, 24);
    return c;
}
#define OSLD_SEED(N) (0xdeadbeefu + ((N) << 2) + 13u)
OSLD u32 inthash(u32 k0) { const u32 s = OSLD_SEED(1u); return bjfinal(s + k0, s, s); }
OSLD u32 inthash(u32 k0, u32 k1) { const u32 s = OSLD_SEED(2u); return bjfinal(s + k0, s + k1, s); }
OSLD u32 inthash(u32 k0, u32 k1, u32 k2)
{
    const u32 s = OSLD_SEED(3u);
    return bjfinal(s + k0, s + k1, s + k2);
}
OSLD u32 inthash(u32 k0, u32 k1, u32 k2, u32 k3)
{
    const u32 s = OSLD_SEED(4u);
    u32 a = s + k0, b = s + k1, c = s + k2;
    bjmix(a, b, c);
    return bjfinal(a + k3, b, c);
}
OSLD u32 inthash(u32 k0, u32 k1, u32 k2, u32 k3, u32 k4)
{
    const u32 s = OSLD_SEED(5u);
    u32 a = s + k0, b = s + k1, c = s + k2;
    bjmix(a, b, c);
    return bjfinal(a + k3, b + k4, c);
}
// 1/(2^32-1) rounded to float == 2^-32
OSLD float bits01(u32 h) { return (float)h * 2.3283064365386963e-10f; }

OSLD int ifloor(float x) { return (int)floorf(x); }
OSLD u32 fbits(float x) { return (u32)__float_as_int(x); }

OSLD int hash_i(int x) { return (int)inthash((u32)x); }
OSLD int hash_f(float x) { return (int)inthash(fbits(x)); }
OSLD int hash_ff(float x, float y) { return (int)inthash(fbits(x), fbits(y)); }
OSLD int hash_v(V3 p) { return (int)inthash(fbits(p.x), fbits(p.y), fbits(p.z)); }
OSLD int hash_vf(V3 p, float t) { return (int)inthash(fbits(p.x), fbits(p.y), fbits(p.z), fbits(t)); }

// ---- cell / hash noise.  CELL: key = floor; else key = float bits ----------
template<bool CELL> OSLD u32 nkey(float v) { return CELL ? (u32)ifloor(v) : fbits(v); }

// DIM inputs, NC outputs.  The 3-output forms append key 0,1,2; for DIM>=3 the
// bjmix of the first three keys is shared by the three results.
template<bool CELL, int DIM, int NC> OSLD void ihnoise(float* out, const float* in)
{
    u32 k[4];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
        k[i] = nkey<CELL>(in[i]);
    if (NC == 1) {
        u32 h = DIM == 1 ? inthash(k[0])
                : DIM == 2 ? inthash(k[0], k[1])
                : DIM == 3 ? inthash(k[0], k[1], k[2])
                           : inthash(k[0], k[1], k[2], k[3]);
        out[0] = bits01(h);
    } else if (DIM == 1) {
#pragma unroll
        for (u32 c = 0; c < 3; ++c)
            out[c] = bits01(inthash(k[0], c));
    } else if (DIM == 2) {
#pragma unroll
        for (u32 c = 0; c < 3; ++c)
            out[c] = bits01(inthash(k[0], k[1], c));
    } else if (DIM == 3) {
        const u32 s = OSLD_SEED(4u);
        u32 a = s + k[0], b = s + k[1], c = s + k[2];
        bjmix(a, b, c);
#pragma unroll
        for (u32 e = 0; e < 3; ++e)
            out[e] = bits01(bjfinal(a + e, b, c));
    } else {
        const u32 s = OSLD_SEED(5u);
        u32 a = s + k[0], b = s + k[1], c = s + k[2];
        bjmix(a, b, c);
        a += k[3];
#pragma unroll
        for (u32 e = 0; e < 3; ++e)
            out[e] = bits01(bjfinal(a, b + e, c));
    }
}
OSLD float pwrap(float s, float period)
{
    period = floorf(period);
    if (period < 1.0f)
        period = 1.0f;
    return s - period * floorf(s / period);
}

// ---------------------------------------------------------------------------
// Perlin gradient noise
// ---------------------------------------------------------------------------
OSLD int imod(int a, int b)
{
    int r = a % b;
    return r < 0 ? r + b : r;
}
OSLD int iperiod(float p)
{
    int i = ifloor(p);
    return i < 1 ? 1 : i;
}
OSLD float ffrac(float x, int* i)
{
    float f = floorf(x);
    *i      = (int)f;
    return x - f;   // x - float(int(floor(x))) : identical whenever floor(x) fits an int
}
OSLD Df ffrac(Df x, int* i) { return mkd(ffrac(x.val, i), x.dx, x.dy); }

OSLD float fade(float t) { return t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f); }
// same product chain on duals: ((t*t)*t) * ((t*(t*6-15))+10)
OSLD Df fade(Df t)
{
    Df t2 = t * t;
    Df t3 = t2 * t;
    Df a  = t * mkd(6.0f) - mkd(15.0f);
    Df b  = t * a + mkd(10.0f);
    return t3 * b;
}
OSLD float fsel(bool b, float t, float f) { return b ? t : f; }
OSLD Df fsel(bool b, Df t, Df f) { return b ? t : f; }
OSLD float negif(float v, u32 bit) { return __int_as_float(__float_as_int(v) ^ (int)(bit ? 0x80000000u : 0u)); }
OSLD Df negif(Df v, u32 bit)
{
    return mkd(negif(v.val, bit), negif(v.dx, bit), negif(v.dy, bit));
}
OSLD float ftwice(float a) { return 2.0f * a; }
OSLD Df ftwice(Df a) { return 2.0f * a; }

template<class S> OSLD S grad1(u32 hash, S x)
{
    u32 h   = hash & 15u;
    float g = (float)(1 + (h & 7u));
    if (h & 8u)
        g = -g;
    return g * x;
}
template<class S> OSLD S grad2(u32 hash, S x, S y)
{
    u32 h = hash & 7u;
    S u   = fsel(h < 4u, x, y);
    S v   = ftwice(fsel(h < 4u, y, x));
    return negif(u, h & 1u) + negif(v, h & 2u);
}
template<class S> OSLD S grad3(u32 hash, S x, S y, S z)
{
    u32 h = hash & 15u;
    S u   = fsel(h < 8u, x, y);
    S v   = fsel(h < 4u, y, fsel((h == 12u) | (h == 14u), x, z));
    return negif(u, h & 1u) + negif(v, h & 2u);
}
template<class S> OSLD S grad4(u32 hash, S x, S y, S z, S w)
{
    u32 h = hash & 31u;
    S u   = fsel(h < 24u, x, y);
    S v   = fsel(h < 16u, y, z);
    S s   = fsel(h < 8u, z, w);
    return negif(u, h & 1u) + negif(v, h & 2u) + negif(s, h & 4u);
}

OSLD float one_minus(float u) { return 1.0f - u; }
OSLD Df one_minus(Df u) { return mkd(1.0f - u.val, 0.0f - u.dx, 0.0f - u.dy); }
template<class S> OSLD S lerp1(S a, S b, S u) { return a * one_minus(u) + b * u; }
template<class S> OSLD S lerp2(S v0, S v1, S v2, S v3, S s, S t)
{
    S s1 = one_minus(s);
    return one_minus(t) * (v0 * s1 + v1 * s) + t * (v2 * s1 + v3 * s);
}
template<class S>
OSLD S lerp3(S v0, S v1, S v2, S v3, S v4, S v5, S v6, S v7, S s, S t, S r)
{
    S s1 = one_minus(s);
    S t1 = one_minus(t);
    S r1 = one_minus(r);
    return r1 * (t1 * (v0 * s1 + v1 * s) + t * (v2 * s1 + v3 * s))
           + r * (t1 * (v4 * s1 + v5 * s) + t * (v6 * s1 + v7 * s));
}

template<int NC> OSLD u32 hslice(u32 h, int c) { return NC == 1 ? h : ((h >> (8 * c)) & 0xFFu); }

// S = float | Df ; NC = 1 | 3 ; PER: wrap lattice coordinates by per[]
template<class S, int NC, bool PER> OSLD void perlin1(S* out, S x, const int* per)
{
    int X;
    S fx = ffrac(x, &X);
    S u  = fade(fx);
    int X1 = X + 1;
    if (PER) { X = imod(X, per[0]); X1 = imod(X1, per[0]); }
    u32 h0 = inthash((u32)X), h1 = inthash((u32)X1);
    S fx1 = fx - 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        out[c] = 0.2500f * lerp1(grad1(hslice<NC>(h0, c), fx), grad1(hslice<NC>(h1, c), fx1), u);
}
template<class S, int NC, bool PER> OSLD void perlin2(S* out, S x, S y, const int* per)
{
    int X, Y;
    S fx = ffrac(x, &X), fy = ffrac(y, &Y);
    S u = fade(fx), v = fade(fy);
    int X1 = X + 1, Y1 = Y + 1;
    if (PER) {
        X = imod(X, per[0]); X1 = imod(X1, per[0]);
        Y = imod(Y, per[1]); Y1 = imod(Y1, per[1]);
    }
    u32 h00 = inthash((u32)X, (u32)Y), h10 = inthash((u32)X1, (u32)Y);
    u32 h01 = inthash((u32)X, (u32)Y1), h11 = inthash((u32)X1, (u32)Y1);
    S fx1 = fx - 1.0f, fy1 = fy - 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        out[c] = 0.6616f
                 * lerp2(grad2(hslice<NC>(h00, c), fx, fy), grad2(hslice<NC>(h10, c), fx1, fy),
                         grad2(hslice<NC>(h01, c), fx, fy1), grad2(hslice<NC>(h11, c), fx1, fy1),
                         u, v);
}
template<class S, int NC, bool PER> OSLD void perlin3(S* out, S x, S y, S z, const int* per)
{
    int X, Y, Z;
    S fx = ffrac(x, &X), fy = ffrac(y, &Y), fz = ffrac(z, &Z);
    S u = fade(fx), v = fade(fy), w = fade(fz);
    int X1 = X + 1, Y1 = Y + 1, Z1 = Z + 1;
    if (PER) {
        X = imod(X, per[0]); X1 = imod(X1, per[0]);
        Y = imod(Y, per[1]); Y1 = imod(Y1, per[1]);
        Z = imod(Z, per[2]); Z1 = imod(Z1, per[2]);
    }
    u32 h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
        h[k] = inthash((u32)((k & 1) ? X1 : X), (u32)((k & 2) ? Y1 : Y), (u32)((k & 4) ? Z1 : Z));
    S fx1 = fx - 1.0f, fy1 = fy - 1.0f, fz1 = fz - 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        out[c] = 0.9820f
                 * lerp3(grad3(hslice<NC>(h[0], c), fx, fy, fz), grad3(hslice<NC>(h[1], c), fx1, fy, fz),
                         grad3(hslice<NC>(h[2], c), fx, fy1, fz), grad3(hslice<NC>(h[3], c), fx1, fy1, fz),
                         grad3(hslice<NC>(h[4], c), fx, fy, fz1), grad3(hslice<NC>(h[5], c), fx1, fy, fz1),
                         grad3(hslice<NC>(h[6], c), fx, fy1, fz1), grad3(hslice<NC>(h[7], c), fx1, fy1, fz1),
                         u, v, w);
}
template<class S, int NC, bool PER> OSLD void perlin4(S* out, S x, S y, S z, S w, const int* per)
{
    int X, Y, Z, W;
    S fx = ffrac(x, &X), fy = ffrac(y, &Y), fz = ffrac(z, &Z), fw = ffrac(w, &W);
    S u = fade(fx), v = fade(fy), t = fade(fz), s = fade(fw);
    int X1 = X + 1, Y1 = Y + 1, Z1 = Z + 1, W1 = W + 1;
    if (PER) {
        X = imod(X, per[0]); X1 = imod(X1, per[0]);
        Y = imod(Y, per[1]); Y1 = imod(Y1, per[1]);
        Z = imod(Z, per[2]); Z1 = imod(Z1, per[2]);
        W = imod(W, per[3]); W1 = imod(W1, per[3]);
    }
    // the bjmix of (x,y,z) is shared by the W and W+1 corners
    const u32 sd = OSLD_SEED(4u);
    u32 h[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        u32 a = sd + (u32)((k & 1) ? X1 : X), b = sd + (u32)((k & 2) ? Y1 : Y),
            c = sd + (u32)((k & 4) ? Z1 : Z);
        bjmix(a, b, c);
        h[k]     = bjfinal(a + (u32)W, b, c);
        h[k + 8] = bjfinal(a + (u32)W1, b, c);
    }
    S fx1 = fx - 1.0f, fy1 = fy - 1.0f, fz1 = fz - 1.0f, fw1 = fw - 1.0f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        S lo = lerp3(grad4(hslice<NC>(h[0], c), fx, fy, fz, fw), grad4(hslice<NC>(h[1], c), fx1, fy, fz, fw),
                     grad4(hslice<NC>(h[2], c), fx, fy1, fz, fw), grad4(hslice<NC>(h[3], c), fx1, fy1, fz, fw),
                     grad4(hslice<NC>(h[4], c), fx, fy, fz1, fw), grad4(hslice<NC>(h[5], c), fx1, fy, fz1, fw),
                     grad4(hslice<NC>(h[6], c), fx, fy1, fz1, fw), grad4(hslice<NC>(h[7], c), fx1, fy1, fz1, fw),
                     u, v, t);
        S hi = lerp3(grad4(hslice<NC>(h[8], c), fx, fy, fz, fw1), grad4(hslice<NC>(h[9], c), fx1, fy, fz, fw1),
                     grad4(hslice<NC>(h[10], c), fx, fy1, fz, fw1), grad4(hslice<NC>(h[11], c), fx1, fy1, fz, fw1),
                     grad4(hslice<NC>(h[12], c), fx, fy, fz1, fw1), grad4(hslice<NC>(h[13], c), fx1, fy, fz1, fw1),
                     grad4(hslice<NC>(h[14], c), fx, fy1, fz1, fw1), grad4(hslice<NC>(h[15], c), fx1, fy1, fz1, fw1),
                     u, v, t);
        out[c] = 0.8344f * lerp1(lo, hi, s);
    }
}
// unified entry: SIGNED selects snoise vs noise = 0.5*(s+1)
template<class S, int DIM, int NC, bool SIGNED, bool PER>
OSLD void perlin(S* out, const S* in, const int* per)
{
    if (DIM == 1) perlin1<S, NC, PER>(out, in[0], per);
    else if (DIM == 2) perlin2<S, NC, PER>(out, in[0], in[1], per);
    else if (DIM == 3) perlin3<S, NC, PER>(out, in[0], in[1], in[2], per);
    else perlin4<S, NC, PER>(out, in[0], in[1], in[2], in[3], per);
    if (!SIGNED) {
#pragma unroll
        for (int c = 0; c < NC; ++c)
            out[c] = 0.5f * (out[c] + 1.0f);
    }
}

// ---------------------------------------------------------------------------
// OIIO fast_* transcendentals (the reference default, USE_FAST_MATH=ON):
// same polynomials and evaluation order as OpenImageIO fmath.h.
// ---------------------------------------------------------------------------
#define OSLD_FLT_MIN 1.17549435e-38f
#define OSLD_FLT_MAX 3.40282347e+38f
#define OSLD_PI 3.14159265358979323846
OSLD float madd(float a, float b, float c) { return a * b + c; }

OSLD void fast_sincos(float x, float* sine, float* cosine)
{
    int q    = (int)rintf(x * (float)(1.0 / OSLD_PI));
    float qf = (float)q;
    x        = madd(qf, -0.78515625f * 4, x);
    x        = madd(qf, -0.00024187564849853515625f * 4, x);
    x        = madd(qf, -3.7747668102383613586e-08f * 4, x);
    x        = madd(qf, -1.2816720341285448015e-12f * 4, x);
    x        = (float)(OSLD_PI / 2) - ((float)(OSLD_PI / 2) - x);
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
    if (fabsf(su) > 1.0f)
        su = 0.0f;
    if (fabsf(cu) > 1.0f)
        cu = 0.0f;
    *sine   = su;
    *cosine = cu;
}
OSLD float fast_sin(float x) { float s, c; fast_sincos(x, &s, &c); return s; }
OSLD float fast_cos(float x) { float s, c; fast_sincos(x, &s, &c); return c; }
OSLD float fast_tan(float x)
{
    int q    = (int)rintf(x * (float)(2.0 / OSLD_PI));
    float qf = (float)q;
    x        = madd(qf, -0.78515625f * 2, x);
    x        = madd(qf, -0.00024187564849853515625f * 2, x);
    x        = madd(qf, -3.7747668102383613586e-08f * 2, x);
    x        = madd(qf, -1.2816720341285448015e-12f * 2, x);
    if ((q & 1) == 0)
        x = (float)(OSLD_PI / 4) - ((float)(OSLD_PI / 4) - x);
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
OSLD float fast_acos(float x)
{
    const float f = fabsf(x);
    const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
    const float a = sqrtf(1.0f - m)
                    * (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
    return x < 0 ? (float)OSLD_PI - a : a;
}
OSLD float fast_asin(float x)
{
    const float f = fabsf(x);
    const float m = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
    const float a = (float)(OSLD_PI / 2)
                    - sqrtf(1.0f - m)
                          * (1.5707963267f + m * (-0.213300989f + m * (0.077980478f + m * -0.02164095f)));
    return copysignf(a, x);
}
OSLD float fast_atan(float x)
{
    const float a = fabsf(x);
    const float k = a > 1.0f ? 1 / a : a;
    const float s = 1.0f - (1.0f - k);
    const float t = s * s;
    float r = s * madd(0.43157974f, t, 1.0f) / madd(madd(0.05831938f, t, 0.76443945f), t, 1.0f);
    if (a > 1.0f)
        r = 1.570796326794896557998982f - r;
    return copysignf(r, x);
}
OSLD float fast_atan2(float y, float x)
{
    const float a = fabsf(x);
    const float b = fabsf(y);
    const float k = (b == 0) ? 0.0f : ((a == b) ? 1.0f : (b > a ? a / b : b / a));
    const float s = 1.0f - (1.0f - k);
    const float t = s * s;
    float r = s * madd(0.43157974f, t, 1.0f) / madd(madd(0.05831938f, t, 0.76443945f), t, 1.0f);
    if (b > a)
        r = 1.570796326794896557998982f - r;
    if (fbits(x) & 0x80000000u)
        r = (float)OSLD_PI - r;
    return copysignf(r, y);
}
OSLD float fast_log2(float x)
{
    x = x < OSLD_FLT_MIN ? OSLD_FLT_MIN : (x > OSLD_FLT_MAX ? OSLD_FLT_MAX : x);
    u32 bits     = fbits(x);
    int exponent = (int)(bits >> 23) - 127;
    float f      = __int_as_float((int)((bits & 0x007FFFFFu) | 0x3f800000u)) - 1.0f;
    float f2     = f * f;
    float f4     = f2 * f2;
    float hi     = madd(f, -0.00931049621349f, 0.05206469089414f);
    float lo     = madd(f, 0.47868480909345f, -0.72116591947498f);
    hi           = madd(f, hi, -0.13753123777116f);
    hi           = madd(f, hi, 0.24187369696082f);
    hi           = madd(f, hi, -0.34730547155299f);
    lo           = madd(f, lo, 1.442689881667200f);
    return ((f4 * hi) + (f * lo)) + (float)exponent;
}
#define OSLD_LN2 0.69314718055994530942
#define OSLD_LN10 2.30258509299404568402
OSLD float fast_log(float x) { return fast_log2(x) * (float)OSLD_LN2; }
OSLD float fast_log10(float x) { return fast_log2(x) * (float)(OSLD_LN2 / OSLD_LN10); }
OSLD float fast_logb(float x)
{
    x = fabsf(x);
    if (x < OSLD_FLT_MIN) x = OSLD_FLT_MIN;
    if (x > OSLD_FLT_MAX) x = OSLD_FLT_MAX;
    return (float)((int)(fbits(x) >> 23) - 127);
}
OSLD float fast_exp2(float x)
{
    if (x < -126.0f) x = -126.0f;
    if (x > 126.0f) x = 126.0f;
    int m = (int)x;
    x -= (float)m;
    x       = 1.0f - (1.0f - x);
    float r = 1.33336498402e-3f;
    r       = madd(x, r, 9.810352697968e-3f);
    r       = madd(x, r, 5.551834031939e-2f);
    r       = madd(x, r, 0.2401793301105f);
    r       = madd(x, r, 0.693144857883f);
    r       = madd(x, r, 1.0f);
    return __int_as_float((int)(fbits(r) + ((u32)m << 23)));
}
OSLD float fast_exp(float x) { return fast_exp2(x * (float)(1.0 / OSLD_LN2)); }
OSLD float fast_expm1(float x)
{
    if (fabsf(x) < 0.03f) {
        float y = 1.0f - (1.0f - x);
        return copysignf(madd(0.5f, y * y, y), x);
    }
    return fast_exp(x) - 1.0f;
}
OSLD float fast_sinh(float x)
{
    float a = fabsf(x);
    if (a > 1.0f) {
        float e = fast_exp(a);
        return copysignf(0.5f * e - 0.5f / e, x);
    }
    a        = 1.0f - (1.0f - a);
    float a2 = a * a;
    float r  = 2.03945513931e-4f;
    r        = madd(r, a2, 8.32990277558e-3f);
    r        = madd(r, a2, 0.1666673421859f);
    r        = madd(r * a, a2, a);
    return copysignf(r, x);
}
OSLD float fast_cosh(float x) { float e = fast_exp(fabsf(x)); return 0.5f * e + 0.5f / e; }
OSLD float fast_tanh(float x) { float e = fast_exp(2.0f * fabsf(x)); return copysignf(1 - 2 / (1 + e), x); }
OSLD float fast_safe_pow(float x, float y)
{
    if (y == 0) return 1.0f;
    if (x == 0) return 0.0f;
    if (y == 1.0f) return x;
    if (y == 2.0f) return fminf(x * x, OSLD_FLT_MAX);
    float sign = 1.0f;
    if (x < 0) {
        int ybits = __float_as_int(y) & 0x7fffffff;
        if (ybits >= 0x4b800000) {
        } else if (ybits >= 0x3f800000) {
            int k = (ybits >> 23) - 127;
            int j = ybits >> (23 - k);
            if ((j << (23 - k)) == ybits)
                sign = __int_as_float((int)(0x3f800000u | ((u32)j << 31)));
            else
                return 0.0f;
        } else {
            return 0.0f;
        }
    }
    return sign * fast_exp2(y * fast_log2(fabsf(x)));
}
OSLD float fast_erf(float x)
{
    const float a1 = 0.0705230784f, a2 = 0.0422820123f, a3 = 0.0092705272f,
                a4 = 0.0001520143f, a5 = 0.0002765672f, a6 = 0.0000430638f;
    const float a = fabsf(x);
    const float b = 1.0f - (1.0f - a);
    const float r = madd(madd(madd(madd(madd(madd(a6, b, a5), b, a4), b, a3), b, a2), b, a1), b, 1.0f);
    const float s = r * r;
    const float t = s * s;
    const float u = t * t;
    const float v = u * u;
    return copysignf(1.0f - 1.0f / v, x);
}
OSLD float fast_erfc(float x) { return 1.0f - fast_erf(x); }
// Mike Giles' single-precision erfinv polynomial, as OIIO fast_ierf (fmath.h)
OSLD float fast_ierf(float x)
{
    float a = fminf(fabsf(x), 0.99999994f);
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
        w = sqrtf(w) - 3.0f;
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
OSLD float fast_cbrt(float x)
{
    float x0 = fabsf(x);
    float a  = __int_as_float((int)(0x2a5137a0u + fbits(x0) / 3u));
    a        = 0.333333333f * (2.0f * a + x0 / (a * a));
    a        = 0.333333333f * (2.0f * a + x0 / (a * a));
    a        = (x0 == 0) ? 0 : a;
    return copysignf(a, x);
}
OSLD float safe_sqrt(float x) { return x >= 0.0f ? sqrtf(x) : 0.0f; }
OSLD float safe_inversesqrt(float x) { return x > 0.0f ? 1.0f / sqrtf(x) : 0.0f; }
OSLD float safe_fmod(float a, float b)
{
    if (b != 0.0f) {
        int N = (int)(a / b);
        return a - (float)N * b;
    }
    return 0.0f;
}
OSLD bool finitef(float q) { return fabsf(q) <= OSLD_FLT_MAX; }
OSLD float safe_div(float a, float b)
{
    float q = a / b;
    return finitef(q) ? q : 0.0f;
}

// ---------------------------------------------------------------------------
// per-component ops over S in {float, Df}; int forms where OSL has them
// ---------------------------------------------------------------------------
OSLD float o_add(float a, float b) { return a + b; }
OSLD Df o_add(Df a, Df b) { return a + b; }
OSLD Df o_add(Df a, float b) { return a + mkd(b); }
OSLD Df o_add(float a, Df b) { return mkd(a) + b; }
OSLD int o_add(int a, int b) { return (int)((u32)a + (u32)b); }
OSLD float o_sub(float a, float b) { return a - b; }
OSLD Df o_sub(Df a, Df b) { return a - b; }
OSLD Df o_sub(Df a, float b) { return a - mkd(b); }
OSLD Df o_sub(float a, Df b) { return mkd(a) - b; }
OSLD int o_sub(int a, int b) { return (int)((u32)a - (u32)b); }
OSLD float o_mul(float a, float b) { return a * b; }
OSLD Df o_mul(Df a, Df b) { return a * b; }
OSLD Df o_mul(Df a, float b) { return a * mkd(b); }
OSLD Df o_mul(float a, Df b) { return mkd(a) * b; }
OSLD int o_mul(int a, int b) { return (int)((u32)a * (u32)b); }
OSLD float o_div(float a, float b) { return safe_div(a, b); }
OSLD Df o_div(Df a, Df b)
{
    float q    = safe_div(a.val, b.val);
    float binv = safe_div(1.0f, b.val);
    return mkd(q, binv * (a.dx - q * b.dx), binv * (a.dy - q * b.dy));
}
OSLD Df o_div(Df a, float b) { return o_div(a, mkd(b)); }
OSLD Df o_div(float a, Df b) { return o_div(mkd(a), b); }
OSLD int o_div(int a, int b) { return b != 0 ? a / b : 0; }
OSLD float o_divc(float a, float b) { return a / b; }
OSLD Df o_divc(Df a, Df b) { return a / b; }
OSLD Df o_divc(Df a, float b) { return a / mkd(b); }
OSLD Df o_divc(float a, Df b) { return mkd(a) / b; }
OSLD int o_divc(int a, int b) { return a / b; }
OSLD int o_mod(int a, int b) { return b != 0 ? a % b : 0; }
OSLD float o_neg(float a) { return -a; }
OSLD Df o_neg(Df a) { return -a; }
OSLD int o_neg(int a) { return -a; }

#define OSLD_UNARY(name, fexpr, dexpr)                 \
    OSLD float o_##name(float a) { return fexpr; }     \
    OSLD Df o_##name(Df a) { return dexpr; }

OSLD Df d_sin(Df a) { float s, c; fast_sincos(a.val, &s, &c); return chain(a, s, c); }
OSLD Df d_cos(Df a) { float s, c; fast_sincos(a.val, &s, &c); return chain(a, c, -s); }
OSLD Df d_tan(Df a) { float t = fast_tan(a.val), c = fast_cos(a.val); return chain(a, t, 1 / (c * c)); }
OSLD Df d_asin(Df a)
{
    float f  = fast_asin(a.val);
    float df = fabsf(a.val) < 1.0f ? 1.0f / sqrtf(1.0f - a.val * a.val) : 0.0f;
    return chain(a, f, df);
}
OSLD Df d_acos(Df a)
{
    float f  = fast_acos(a.val);
    float df = fabsf(a.val) < 1.0f ? -1.0f / sqrtf(1.0f - a.val * a.val) : 0.0f;
    return chain(a, f, df);
}
OSLD Df d_atan(Df a) { return chain(a, fast_atan(a.val), 1.0f / (1.0f + a.val * a.val)); }
OSLD Df d_sinh(Df a) { return chain(a, fast_sinh(a.val), fast_cosh(a.val)); }
OSLD Df d_cosh(Df a) { return chain(a, fast_cosh(a.val), fast_sinh(a.val)); }
OSLD Df d_tanh(Df a) { float t = fast_tanh(a.val), c = fast_cosh(a.val); return chain(a, t, 1.0f / (c * c)); }
OSLD Df d_log(Df a) { return chain(a, fast_log(a.val), a.val < OSLD_FLT_MIN ? 0.0f : 1.0f / a.val); }
OSLD Df d_log2(Df a)
{
    float al = a.val * (float)OSLD_LN2;
    return chain(a, fast_log2(a.val), al < OSLD_FLT_MIN ? 0.0f : 1.0f / al);
}
OSLD Df d_log10(Df a)
{
    float al = a.val * (float)OSLD_LN10;
    return chain(a, fast_log10(a.val), al < OSLD_FLT_MIN ? 0.0f : 1.0f / al);
}
OSLD Df d_exp(Df a) { float f = fast_exp(a.val); return chain(a, f, f); }
OSLD Df d_exp2(Df a) { float f = fast_exp2(a.val); return chain(a, f, f * (float)OSLD_LN2); }
OSLD Df d_expm1(Df a) { return chain(a, fast_expm1(a.val), fast_exp(a.val)); }
OSLD Df d_erf(Df a) { return chain(a, fast_erf(a.val), fast_exp(-a.val * a.val) * 1.128379167095512573896158903f); }
OSLD Df d_erfc(Df a) { return chain(a, fast_erfc(a.val), fast_exp(-a.val * a.val) * -1.128379167095512573896158903f); }
OSLD Df d_cbrt(Df a)
{
    if (a.val != 0.0f) {
        float f = fast_cbrt(a.val);
        return chain(a, f, 1.0f / (3.0f * f * f));
    }
    return mkd(0.0f);
}
OSLD Df d_sqrt(Df a)
{
    if (a.val > 0.0f) {
        float f = sqrtf(a.val);
        return chain(a, f, 0.5f / f);
    }
    return mkd(0.0f);
}
OSLD Df d_inversesqrt(Df a)
{
    if (a.val > 0.0f) {
        float f = 1.0f / sqrtf(a.val);
        return chain(a, f, -0.5f * f / a.val);
    }
    return mkd(0.0f);
}
OSLD Df d_fabs(Df a) { return a.val >= 0 ? a : -a; }
OSLD float sign_f(float x) { return x < 0.0f ? -1.0f : (x == 0.0f ? 0.0f : 1.0f); }

OSLD_UNARY(sin, fast_sin(a), d_sin(a))
OSLD_UNARY(cos, fast_cos(a), d_cos(a))
OSLD_UNARY(tan, fast_tan(a), d_tan(a))
OSLD_UNARY(asin, fast_asin(a), d_asin(a))
OSLD_UNARY(acos, fast_acos(a), d_acos(a))
OSLD_UNARY(atan, fast_atan(a), d_atan(a))
OSLD_UNARY(sinh, fast_sinh(a), d_sinh(a))
OSLD_UNARY(cosh, fast_cosh(a), d_cosh(a))
OSLD_UNARY(tanh, fast_tanh(a), d_tanh(a))
OSLD_UNARY(log, fast_log(a), d_log(a))
OSLD_UNARY(log2, fast_log2(a), d_log2(a))
OSLD_UNARY(log10, fast_log10(a), d_log10(a))
OSLD_UNARY(exp, fast_exp(a), d_exp(a))
OSLD_UNARY(exp2, fast_exp2(a), d_exp2(a))
OSLD_UNARY(expm1, fast_expm1(a), d_expm1(a))
OSLD_UNARY(erf, fast_erf(a), d_erf(a))
OSLD_UNARY(erfc, fast_erfc(a), d_erfc(a))
OSLD_UNARY(cbrt, fast_cbrt(a), d_cbrt(a))
OSLD_UNARY(sqrt, safe_sqrt(a), d_sqrt(a))
OSLD_UNARY(inversesqrt, safe_inversesqrt(a), d_inversesqrt(a))
OSLD_UNARY(abs, fabsf(a), d_fabs(a))
OSLD_UNARY(fabs, fabsf(a), d_fabs(a))
OSLD_UNARY(floor, floorf(a), mkd(floorf(a.val)))
OSLD_UNARY(ceil, ceilf(a), mkd(ceilf(a.val)))
OSLD_UNARY(round, roundf(a), mkd(roundf(a.val)))
OSLD_UNARY(trunc, truncf(a), mkd(truncf(a.val)))
OSLD_UNARY(sign, sign_f(a), mkd(sign_f(a.val)))
OSLD_UNARY(logb, fast_logb(a), mkd(fast_logb(a.val)))
OSLD int o_abs(int a) { return a < 0 ? -a : a; }
OSLD int o_fabs(int a) { return a < 0 ? -a : a; }

// binary / ternary with mixed float|Df operands: promote through a macro
#define OSLD_PROMOTE2(name)                                   \
    OSLD Df o_##name(Df a, float b) { return o_##name(a, mkd(b)); } \
    OSLD Df o_##name(float a, Df b) { return o_##name(mkd(a), b); }

OSLD float o_atan2(float y, float x) { return fast_atan2(y, x); }
OSLD Df o_atan2(Df y, Df x)
{
    float f     = fast_atan2(y.val, x.val);
    float denom = (x.val == 0 && y.val == 0) ? 0.0f : 1.0f / (x.val * x.val + y.val * y.val);
    return chain(y, x, f, -x.val * denom, y.val * denom);
}
OSLD_PROMOTE2(atan2)
OSLD float o_pow(float x, float y) { return fast_safe_pow(x, y); }
OSLD Df o_pow(Df u, Df v)
{
    float powuvm1 = fast_safe_pow(u.val, v.val - 1.0f);
    float powuv   = powuvm1 * u.val;
    float logu    = u.val > 0 ? fast_log(u.val) : 0.0f;
    return chain(u, v, powuv, v.val * powuvm1, logu * powuv);
}
OSLD_PROMOTE2(pow)
OSLD float o_fmod(float a, float b) { return safe_fmod(a, b); }
OSLD Df o_fmod(Df a, Df b) { return mkd(safe_fmod(a.val, b.val), a.dx, a.dy); }
OSLD_PROMOTE2(fmod)
OSLD int o_fmod(int a, int b) { return o_mod(a, b); }
OSLD float o_step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
// step has no derivative form (osl_step_fff only): the result's derivatives are zero
OSLD float o_step(Df edge, Df x) { return o_step(edge.val, x.val); }
OSLD float o_step(float edge, Df x) { return o_step(edge, x.val); }
OSLD float o_step(Df edge, float x) { return o_step(edge.val, x); }
OSLD float o_min(float a, float b) { return a <= b ? a : b; }
OSLD float o_max(float a, float b) { return a > b ? a : b; }
OSLD Df o_min(Df a, Df b) { return a.val <= b.val ? a : b; }
OSLD Df o_max(Df a, Df b) { return a.val > b.val ? a : b; }
OSLD_PROMOTE2(min)
OSLD_PROMOTE2(max)
OSLD int o_min(int a, int b) { return a <= b ? a : b; }
OSLD int o_max(int a, int b) { return a > b ? a : b; }

OSLD Df asd(float a) { return mkd(a); }
OSLD Df asd(Df a) { return a; }
OSLD float o_mix(float a, float b, float x) { return a * (1.0f - x) + b * x; }
OSLD Df o_mix_d(Df a, Df b, Df x)
{
    float omx = 1.0f - x.val;
    float r   = a.val * omx + b.val * x.val;
    float rx  = ((a.dx * omx - a.val * x.dx) + b.val * x.dx) + b.dx * x.val;
    float ry  = ((a.dy * omx - a.val * x.dy) + b.val * x.dy) + b.dy * x.val;
    return mkd(r, rx, ry);
}
OSLD float o_smoothstep(float e0, float e1, float x)
{
    if (x < e0) return 0.0f;
    else if (x >= e1) return 1.0f;
    float t = (x - e0) / (e1 - e0);
    return (3.0f - 2.0f * t) * (t * t);
}
OSLD Df o_smoothstep_d(Df e0, Df e1, Df x)
{
    if (x.val < e0.val) return mkd(0.0f);
    else if (x.val >= e1.val) return mkd(1.0f);
    Df t = (x - e0) / (e1 - e0);
    return (3.0f - 2.0f * t) * t * t;
}
OSLD float o_clamp(float x, float lo, float hi) { float t = x < lo ? lo : x; return t > hi ? hi : t; }
OSLD Df o_clamp_d(Df x, Df lo, Df hi) { Df t = x.val < lo.val ? lo : x; return t.val > hi.val ? hi : t; }
OSLD int o_clamp(int x, int lo, int hi) { int t = x < lo ? lo : x; return t > hi ? hi : t; }
OSLD float o_select(float a, float b, float c) { return c != 0.0f ? b : a; }
OSLD Df o_select_d(Df a, Df b, Df c) { return c.val != 0.0f ? b : a; }
// any Df operand -> dual form
#define OSLD_TERNARY_D(name)                                                              \
    template<class A, class B, class C> OSLD Df o_##name(A a, B b, C c) { return o_##name##_d(asd(a), asd(b), asd(c)); }
OSLD_TERNARY_D(mix)
OSLD_TERNARY_D(smoothstep)
OSLD_TERNARY_D(clamp)
OSLD_TERNARY_D(select)

// ---- vector functions --------------------------------------------------------
OSLD float imath_length(V3 v)
{
    float l2 = v.x * v.x + v.y * v.y + v.z * v.z;
    if (l2 < 2.0f * OSLD_FLT_MIN) {
        float ax = fabsf(v.x), ay = fabsf(v.y), az = fabsf(v.z);
        float m = ax;
        if (m < ay) m = ay;
        if (m < az) m = az;
        if (m == 0.0f) return 0.0f;
        ax /= m; ay /= m; az /= m;
        return m * sqrtf(ax * ax + ay * ay + az * az);
    }
    return sqrtf(l2);
}
OSLD Dv asdv(V3 a) { return mkdv(a); }
OSLD Dv asdv(const Dv& a) { return a; }
OSLD float o_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
OSLD Df o_dot_d(const Dv& a, const Dv& b)
{
    return getc(a, 0) * getc(b, 0) + getc(a, 1) * getc(b, 1) + getc(a, 2) * getc(b, 2);
}
OSLD Df o_dot(const Dv& a, const Dv& b) { return o_dot_d(a, b); }
OSLD Df o_dot(const Dv& a, V3 b) { return o_dot_d(a, mkdv(b)); }
OSLD Df o_dot(V3 a, const Dv& b) { return o_dot_d(mkdv(a), b); }
OSLD V3 o_cross(V3 a, V3 b) { return cross3(a, b); }
OSLD Dv dv3(Df x, Df y, Df z)
{
    return mkdv(mkv(x.val, y.val, z.val), mkv(x.dx, y.dx, z.dx), mkv(x.dy, y.dy, z.dy));
}
OSLD Dv o_cross_d(const Dv& a, const Dv& b)
{
    Df ax = getc(a, 0), ay = getc(a, 1), az = getc(a, 2);
    Df bx = getc(b, 0), by = getc(b, 1), bz = getc(b, 2);
    return dv3(ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx);
}
OSLD Dv o_cross(const Dv& a, const Dv& b) { return o_cross_d(a, b); }
OSLD Dv o_cross(const Dv& a, V3 b) { return o_cross_d(a, mkdv(b)); }
OSLD Dv o_cross(V3 a, const Dv& b) { return o_cross_d(mkdv(a), b); }
OSLD float o_length(V3 a) { return imath_length(a); }
OSLD Df o_length(const Dv& a)
{
    Df ax = getc(a, 0), ay = getc(a, 1), az = getc(a, 2);
    return d_sqrt(ax * ax + ay * ay + az * az);
}
OSLD float o_distance(V3 a, V3 b)
{
    float x = a.x - b.x, y = a.y - b.y, z = a.z - b.z;
    return sqrtf(x * x + y * y + z * z);
}
OSLD Df o_distance(const Dv& a, const Dv& b) { return o_length(a - b); }
OSLD Df o_distance(const Dv& a, V3 b) { return o_length(a - mkdv(b)); }
OSLD Df o_distance(V3 a, const Dv& b) { return o_length(mkdv(a) - b); }
OSLD V3 o_normalize(V3 v)
{
    float len = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
    if (len > 0.0f) {
        float inv = 1.0f / len;
        return mkv(v.x * inv, v.y * inv, v.z * inv);
    }
    return mkv(0.0f);
}
OSLD Dv o_normalize(const Dv& a)
{
    Df ax = getc(a, 0), ay = getc(a, 1), az = getc(a, 2);
    Df len = d_sqrt(ax * ax + ay * ay + az * az);
    if (len.val > 0.0f) {
        Df inv = 1.0f / len;
        return dv3(ax * inv, ay * inv, az * inv);
    }
    return mkdv(mkv(0.0f));
}
OSLD float filter_width(float dx, float dy) { return sqrtf(dx * dx + dy * dy); }
OSLD float o_filterwidth(Df x) { return filter_width(x.dx, x.dy); }
OSLD float o_filterwidth(float) { return 0.0f; }
OSLD V3 o_filterwidth(const Dv& x)
{
    return mkv(filter_width(x.dx.x, x.dy.x), filter_width(x.dx.y, x.dy.y), filter_width(x.dx.z, x.dy.z));
}
OSLD V3 o_filterwidth(V3) { return mkv(0.0f); }
OSLD V3 o_calculatenormal(const Dv& P, bool flip) { return flip ? cross3(P.dy, P.dx) : cross3(P.dx, P.dy); }
OSLD float o_area(const Dv& P) { return imath_length(cross3(P.dx, P.dy)); }
OSLD float o_Dx(Df a) { return a.dx; }
OSLD float o_Dy(Df a) { return a.dy; }
OSLD V3 o_Dx(const Dv& a) { return a.dx; }
OSLD V3 o_Dy(const Dv& a) { return a.dy; }
OSLD float o_Dx(float) { return 0.0f; }
OSLD float o_Dy(float) { return 0.0f; }
OSLD V3 o_Dx(V3) { return mkv(0.0f); }
OSLD V3 o_Dy(V3) { return mkv(0.0f); }

}  // namespace osld

#include "osl_b200_simplex.cuh"
#include "osl_b200_spline.cuh"
#include "osl_b200_gabor.cuh"
#include "osl_b200_matrix.cuh"
