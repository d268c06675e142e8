// osl_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain scalar C++ restatement of the arithmetic on OSL's data-parallel hot
// path, written from the behaviour of the reference sources cited at each
// function (paths relative to /root/reference).  Only tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs may use it; the CUDA product
// under openshadinglanguage_b200/ never includes or links this file.
//
// Canonical formulation = the reference's NON-SIMD branches (CGScalar,
// src/include/OSL/oslnoise.h:1363-1376,1433-1451,1504-1537), which is what the
// reference's batched (AVX-512) and __CUDA_ARCH__ builds execute.
//
// Build with  g++ -std=c++17 -O2 -ffp-contract=off  (the reference JIT default
// is no FMA contraction, CHANGES.md:36).
//
// Third-party arithmetic not present under /root/reference (OpenImageIO >= 3.0
// fmath.h / hash.h): bjhash::bjmix/bjfinal (Bob Jenkins lookup3, pinned
// bit-exactly by testsuite/hash/ref/out.txt), ifloor, lerp/bilerp/trilerp
// (pinned to 1e-3 by src/liboslnoise/oslnoise_test.cpp:85-171 and to 8 bits by
// the testsuite noise images), fast_* transcendentals (restated from OIIO's
// published polynomial forms; value-level PARITY UNPINNED — gated only by the
// reference's image thresholds and reduced-precision text goldens).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace oslo {

// ---------------------------------------------------------------------------
// basic types
// ---------------------------------------------------------------------------
struct V3 {
    float x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(float a) : x(a), y(a), z(a) {}
    V3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline V3 operator*(V3 a, float b) { return V3(a.x * b, a.y * b, a.z * b); }
inline V3 operator*(float b, V3 a) { return V3(a.x * b, a.y * b, a.z * b); }
inline V3 operator/(V3 a, V3 b) { return V3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline V3 operator/(V3 a, float b) { return V3(a.x / b, a.y / b, a.z / b); }
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z,
              a.x * b.y - a.y * b.x);
}

// Forward-mode dual number with two partials (dx, dy).
// Arithmetic rules follow src/include/OSL/dual.h:400-600.
template<class T> struct Dual2 {
    T val, dx, dy;
    Dual2() : val(), dx(), dy() {}
    Dual2(const T& v) : val(v), dx(), dy() {}
    Dual2(const T& v, const T& a, const T& b) : val(v), dx(a), dy(b) {}
};
typedef Dual2<float> Df;
typedef Dual2<V3> Dv;

template<class T> inline Dual2<T> operator+(const Dual2<T>& a, const Dual2<T>& b)
{
    return Dual2<T>(a.val + b.val, a.dx + b.dx, a.dy + b.dy);
}
template<class T> inline Dual2<T> operator+(const Dual2<T>& a, const T& b)
{
    return Dual2<T>(a.val + b, a.dx, a.dy);
}
template<class T> inline Dual2<T> operator+(const T& a, const Dual2<T>& b)
{
    return Dual2<T>(a + b.val, b.dx, b.dy);
}
template<class T> inline Dual2<T> operator-(const Dual2<T>& a, const Dual2<T>& b)
{
    return Dual2<T>(a.val - b.val, a.dx - b.dx, a.dy - b.dy);
}
template<class T> inline Dual2<T> operator-(const Dual2<T>& a, const T& b)
{
    return Dual2<T>(a.val - b, a.dx, a.dy);
}
template<class T> inline Dual2<T> operator-(const T& a, const Dual2<T>& b)
{
    return Dual2<T>(a - b.val, -b.dx, -b.dy);
}
template<class T> inline Dual2<T> operator-(const Dual2<T>& a)
{
    return Dual2<T>(-a.val, -a.dx, -a.dy);
}
// dual.h:502-512  (a.val*b.partial + a.partial*b.val)
inline Df operator*(const Df& a, const Df& b)
{
    return Df(a.val * b.val, a.val * b.dx + a.dx * b.val,
              a.val * b.dy + a.dy * b.val);
}
inline Dv operator*(const Dv& a, const Df& b)
{
    return Dv(a.val * b.val, a.val * b.dx + a.dx * b.val,
              a.val * b.dy + a.dy * b.val);
}
inline Dv operator*(const Df& b, const Dv& a) { return a * b; }
inline Dv operator*(const Dv& a, const Dv& b)
{
    return Dv(a.val * b.val, a.val * b.dx + a.dx * b.val,
              a.val * b.dy + a.dy * b.val);
}
inline Df operator*(const Df& a, float b) { return Df(a.val * b, a.dx * b, a.dy * b); }
inline Df operator*(float b, const Df& a) { return Df(a.val * b, a.dx * b, a.dy * b); }
inline Dv operator*(const Dv& a, float b) { return Dv(a.val * b, a.dx * b, a.dy * b); }
inline Dv operator*(float b, const Dv& a) { return Dv(a.val * b, a.dx * b, a.dy * b); }
inline Dv operator*(const Dv& a, const V3& b) { return Dv(a.val * b, a.dx * b, a.dy * b); }
inline Dv operator*(const V3& b, const Dv& a) { return Dv(a.val * b, a.dx * b, a.dy * b); }
// dual.h:569-581
inline Df operator/(const Df& a, const Df& b)
{
    float binv = 1.0f / b.val;
    float q    = a.val / b.val;
    return Df(q, binv * (a.dx - q * b.dx), binv * (a.dy - q * b.dy));
}
inline Df operator/(const Df& a, float b)
{
    float binv = 1.0f / b;
    return Df(a.val / b, binv * a.dx, binv * a.dy);
}
inline Df operator/(float a, const Df& b)
{
    float binv = 1.0f / b.val;
    float q    = a / b.val;
    return Df(q, binv * (-q * b.dx), binv * (-q * b.dy));
}
inline bool operator<(const Df& a, const Df& b) { return a.val < b.val; }
inline bool operator<(const Df& a, float b) { return a.val < b; }
inline bool operator>(const Df& a, float b) { return a.val > b; }

// dual.h:767-800  chain rule helpers
inline Df dualfunc(const Df& u, float f, float df)
{
    return Df(f, df * u.dx, df * u.dy);
}
inline Df dualfunc(const Df& u, const Df& v, float f, float dfdu, float dfdv)
{
    return Df(f, dfdu * u.dx + dfdv * v.dx, dfdu * u.dy + dfdv * v.dy);
}
inline Df dualfunc(const Df& u, const Df& v, const Df& w, float f, float dfdu,
                   float dfdv, float dfdw)
{
    return Df(f, dfdu * u.dx + dfdv * v.dx + dfdw * w.dx,
              dfdu * u.dy + dfdv * v.dy + dfdw * w.dy);
}

inline float val_of(float a) { return a; }
inline float val_of(const Df& a) { return a.val; }
inline Df comp(const Dv& v, int c) { return Df(v.val[c], v.dx[c], v.dy[c]); }
inline float comp(const V3& v, int c) { return v[c]; }
inline Dv make_dv(const Df& x, const Df& y, const Df& z)
{
    return Dv(V3(x.val, y.val, z.val), V3(x.dx, y.dx, z.dx),
              V3(x.dy, y.dy, z.dy));
}

inline uint32_t f2u(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float u2f(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// OIIO::ifloor — a true floor (pinned by oslnoise_test.cpp:256-263)
inline int ifloor(float x) { return (int)std::floor(x); }

// ---------------------------------------------------------------------------
// integer hash: Bob Jenkins lookup3 (OIIO hash.h bjhash::bjmix / bjfinal);
// call sites src/include/OSL/oslnoise.h:202,210,235-287
// ---------------------------------------------------------------------------
inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

inline void bjmix(uint32_t& a, uint32_t& b, uint32_t& c)
{
    a -= c; a ^= rotl32(c, 4);  c += b;
    b -= a; b ^= rotl32(a, 6);  a += c;
    c -= b; c ^= rotl32(b, 8);  b += a;
    a -= c; a ^= rotl32(c, 16); c += b;
    b -= a; b ^= rotl32(a, 19); a += c;
    c -= b; c ^= rotl32(b, 4);  b += a;
}
inline uint32_t bjfinal(uint32_t a, uint32_t b, uint32_t c)
{
    c ^= b; c -= rotl32(b, 14);
    a ^= c; a -= rotl32(c, 11);
    b ^= a; b -= rotl32(a, 25);
    c ^= b; c -= rotl32(b, 16);
    a ^= c; a -= rotl32(c, 4);
    b ^= a; b -= rotl32(a, 14);
    c ^= b; c -= rotl32(b, 24);
    return c;
}
// oslnoise.h:229-289 — seed 0xdeadbeef + (N<<2) + 13
inline uint32_t inthash(uint32_t k0)
{
    uint32_t s = 0xdeadbeefu + (1u << 2) + 13u;
    return bjfinal(s + k0, s, s);
}
inline uint32_t inthash(uint32_t k0, uint32_t k1)
{
    uint32_t s = 0xdeadbeefu + (2u << 2) + 13u;
    return bjfinal(s + k0, s + k1, s);
}
inline uint32_t inthash(uint32_t k0, uint32_t k1, uint32_t k2)
{
    uint32_t s = 0xdeadbeefu + (3u << 2) + 13u;
    return bjfinal(s + k0, s + k1, s + k2);
}
inline uint32_t inthash(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3)
{
    uint32_t s = 0xdeadbeefu + (4u << 2) + 13u;
    uint32_t a = s + k0, b = s + k1, c = s + k2;
    bjmix(a, b, c);
    return bjfinal(a + k3, b, c);
}
inline uint32_t inthash(uint32_t k0, uint32_t k1, uint32_t k2, uint32_t k3,
                        uint32_t k4)
{
    uint32_t s = 0xdeadbeefu + (5u << 2) + 13u;
    uint32_t a = s + k0, b = s + k1, c = s + k2;
    bjmix(a, b, c);
    return bjfinal(a + k3, b + k4, c);
}
// oslnoise.h:134-141
inline float bits_to_01(uint32_t bits)
{
    const float k = (float)(1.0 / (double)std::numeric_limits<uint32_t>::max());
    return bits * k;
}

// osl_hash_* (oslnoise.h:555-607)
inline int hash_i(int x) { return (int)inthash((uint32_t)x); }
inline int hash_f(float x) { return (int)inthash(f2u(x)); }
inline int hash_ff(float x, float y) { return (int)inthash(f2u(x), f2u(y)); }
inline int hash_v(V3 p) { return (int)inthash(f2u(p.x), f2u(p.y), f2u(p.z)); }
inline int hash_vf(V3 p, float t)
{
    return (int)inthash(f2u(p.x), f2u(p.y), f2u(p.z), f2u(t));
}

// ---------------------------------------------------------------------------
// cell / hash noise (oslnoise.h:337-480).  KEY: 0 = ifloor (cell), 1 = bits.
// ---------------------------------------------------------------------------
template<int KEY> inline uint32_t tokey(float v)
{
    return KEY == 0 ? (uint32_t)ifloor(v) : f2u(v);
}
template<int KEY> inline float ihnoise_f(float x) { return bits_to_01(inthash(tokey<KEY>(x))); }
template<int KEY> inline float ihnoise_f(float x, float y)
{
    return bits_to_01(inthash(tokey<KEY>(x), tokey<KEY>(y)));
}
template<int KEY> inline float ihnoise_f(V3 p)
{
    return bits_to_01(inthash(tokey<KEY>(p.x), tokey<KEY>(p.y), tokey<KEY>(p.z)));
}
template<int KEY> inline float ihnoise_f(V3 p, float t)
{
    return bits_to_01(inthash(tokey<KEY>(p.x), tokey<KEY>(p.y), tokey<KEY>(p.z),
                              tokey<KEY>(t)));
}
// Vec3 result: extra trailing key 0,1,2 (oslnoise.h:412-458)
template<int KEY> inline V3 ihnoise_v(float x)
{
    uint32_t k = tokey<KEY>(x);
    return V3(bits_to_01(inthash(k, 0u)), bits_to_01(inthash(k, 1u)),
              bits_to_01(inthash(k, 2u)));
}
template<int KEY> inline V3 ihnoise_v(float x, float y)
{
    uint32_t k0 = tokey<KEY>(x), k1 = tokey<KEY>(y);
    return V3(bits_to_01(inthash(k0, k1, 0u)), bits_to_01(inthash(k0, k1, 1u)),
              bits_to_01(inthash(k0, k1, 2u)));
}
template<int KEY> inline V3 ihnoise_v(V3 p)
{
    uint32_t k0 = tokey<KEY>(p.x), k1 = tokey<KEY>(p.y), k2 = tokey<KEY>(p.z);
    return V3(bits_to_01(inthash(k0, k1, k2, 0u)),
              bits_to_01(inthash(k0, k1, k2, 1u)),
              bits_to_01(inthash(k0, k1, k2, 2u)));
}
template<int KEY> inline V3 ihnoise_v(V3 p, float t)
{
    uint32_t k0 = tokey<KEY>(p.x), k1 = tokey<KEY>(p.y), k2 = tokey<KEY>(p.z),
             k3 = tokey<KEY>(t);
    return V3(bits_to_01(inthash(k0, k1, k2, k3, 0u)),
              bits_to_01(inthash(k0, k1, k2, k3, 1u)),
              bits_to_01(inthash(k0, k1, k2, k3, 2u)));
}
// periodic wrap (oslnoise.h:532-546)
inline float pwrap(float s, float period)
{
    period = std::floor(period);
    if (period < 1.0f)
        period = 1.0f;
    return s - period * std::floor(s / period);
}
inline V3 pwrap(V3 s, V3 p) { return V3(pwrap(s.x, p.x), pwrap(s.y, p.y), pwrap(s.z, p.z)); }

// ---------------------------------------------------------------------------
// Perlin gradient noise (oslnoise.h:846-1041, 1327-2280)
// ---------------------------------------------------------------------------
inline int imod(int a, int b)
{
    int r = a % b;
    if (r < 0)
        r += b;
    return r;
}
inline float floorfrac(float x, int* i)
{
    *i = ifloor(x);
    return x - *i;
}
inline Df floorfrac(const Df& x, int* i)
{
    float f = floorfrac(x.val, i);
    return Df(f, x.dx, x.dy);
}
template<class T> inline T fade(const T& t)
{
    return t * t * t * (t * (t * T(6.0f) - T(15.0f)) + T(10.0f));
}
inline float select(bool b, float t, float f) { return b ? t : f; }
inline Df select(bool b, const Df& t, const Df& f) { return b ? t : f; }
inline float negate_if(float v, bool c) { return c ? -v : v; }
inline Df negate_if(const Df& v, bool c) { return c ? -v : v; }

template<class T> inline T grad(int hash, const T& x)
{
    int h   = hash & 15;
    float g = 1 + (h & 7);
    if (h & 8)
        g = -g;
    return g * x;
}
template<class T> inline T grad(int hash, const T& x, const T& y)
{
    int h = hash & 7;
    T u   = select(h < 4, x, y);
    T v   = 2.0f * select(h < 4, y, x);
    return negate_if(u, h & 1) + negate_if(v, h & 2);
}
template<class T> inline T grad(int hash, const T& x, const T& y, const T& z)
{
    int h = hash & 15;
    T u   = select(h < 8, x, y);
    T v   = select(h < 4, y, select((h == 12) || (h == 14), x, z));
    return negate_if(u, h & 1) + negate_if(v, h & 2);
}
template<class T>
inline T grad(int hash, const T& x, const T& y, const T& z, const T& w)
{
    int h = hash & 31;
    T u   = select(h < 24, x, y);
    T v   = select(h < 16, y, z);
    T s   = select(h < 8, z, w);
    return negate_if(u, h & 1) + negate_if(v, h & 2) + negate_if(s, h & 4);
}

// OIIO fmath.h lerp family (form recalled from the published header).
template<class T> inline T lerp(const T& a, const T& b, const T& u)
{
    return a * (T(1.0f) - u) + b * u;
}
template<class T>
inline T bilerp(const T& v0, const T& v1, const T& v2, const T& v3, const T& s,
                const T& t)
{
    T s1 = T(1.0f) - s;
    return (T(1.0f) - t) * (v0 * s1 + v1 * s) + t * (v2 * s1 + v3 * s);
}
template<class T>
inline T trilerp(const T& v0, const T& v1, const T& v2, const T& v3, const T& v4,
                 const T& v5, const T& v6, const T& v7, const T& s, const T& t,
                 const T& r)
{
    T s1 = T(1.0f) - s;
    T t1 = T(1.0f) - t;
    T r1 = T(1.0f) - r;
    return r1 * (t1 * (v0 * s1 + v1 * s) + t * (v2 * s1 + v3 * s))
           + r * (t1 * (v4 * s1 + v5 * s) + t * (v6 * s1 + v7 * s));
}

// Corner hashing.  `per` == nullptr: plain; else periods (already >=1 ints).
struct PerlinHash {
    const int* per;
    PerlinHash(const int* p = nullptr) : per(p) {}
    uint32_t operator()(int x) const
    {
        return per ? inthash((uint32_t)imod(x, per[0])) : inthash((uint32_t)x);
    }
    uint32_t operator()(int x, int y) const
    {
        return per ? inthash((uint32_t)imod(x, per[0]), (uint32_t)imod(y, per[1]))
                   : inthash((uint32_t)x, (uint32_t)y);
    }
    uint32_t operator()(int x, int y, int z) const
    {
        return per ? inthash((uint32_t)imod(x, per[0]), (uint32_t)imod(y, per[1]),
                             (uint32_t)imod(z, per[2]))
                   : inthash((uint32_t)x, (uint32_t)y, (uint32_t)z);
    }
    uint32_t operator()(int x, int y, int z, int w) const
    {
        return per ? inthash((uint32_t)imod(x, per[0]), (uint32_t)imod(y, per[1]),
                             (uint32_t)imod(z, per[2]), (uint32_t)imod(w, per[3]))
                   : inthash((uint32_t)x, (uint32_t)y, (uint32_t)z, (uint32_t)w);
    }
};
// channel extraction: scalar noise uses the whole hash; vector noise slices
// bytes 0,1,2 (oslnoise.h:1096-1103)
inline int hchan(uint32_t h, int c) { return c < 0 ? (int)h : (int)((h >> (8 * c)) & 0xFF); }

// NC = 1 (scalar result, c=-1) or 3 (vector result).  T = float or Df.
template<class T, int NC>
inline void perlin1(T* out, const PerlinHash& hash, const T& x)
{
    int X;
    T fx = floorfrac(x, &X);
    T u  = fade(fx);
    uint32_t h0 = hash(X), h1 = hash(X + 1);
    for (int c = 0; c < NC; ++c) {
        int cc = NC == 1 ? -1 : c;
        out[c] = 0.2500f * lerp(grad(hchan(h0, cc), fx), grad(hchan(h1, cc), fx - 1.0f), u);
    }
}
template<class T, int NC>
inline void perlin2(T* out, const PerlinHash& hash, const T& x, const T& y)
{
    int X, Y;
    T fx = floorfrac(x, &X), fy = floorfrac(y, &Y);
    T u = fade(fx), v = fade(fy);
    uint32_t h00 = hash(X, Y), h10 = hash(X + 1, Y), h01 = hash(X, Y + 1),
             h11 = hash(X + 1, Y + 1);
    T fx1 = fx - 1.0f, fy1 = fy - 1.0f;
    for (int c = 0; c < NC; ++c) {
        int cc = NC == 1 ? -1 : c;
        out[c] = 0.6616f
                 * bilerp(grad(hchan(h00, cc), fx, fy), grad(hchan(h10, cc), fx1, fy),
                          grad(hchan(h01, cc), fx, fy1), grad(hchan(h11, cc), fx1, fy1),
                          u, v);
    }
}
template<class T, int NC>
inline void perlin3(T* out, const PerlinHash& hash, const T& x, const T& y, const T& z)
{
    int X, Y, Z;
    T fx = floorfrac(x, &X), fy = floorfrac(y, &Y), fz = floorfrac(z, &Z);
    T u = fade(fx), v = fade(fy), w = fade(fz);
    uint32_t h[8];
    for (int k = 0; k < 8; ++k)
        h[k] = hash(X + (k & 1), Y + ((k >> 1) & 1), Z + (k >> 2));
    T fx1 = fx - 1.0f, fy1 = fy - 1.0f, fz1 = fz - 1.0f;
    for (int c = 0; c < NC; ++c) {
        int cc = NC == 1 ? -1 : c;
        out[c] = 0.9820f
                 * trilerp(grad(hchan(h[0], cc), fx, fy, fz), grad(hchan(h[1], cc), fx1, fy, fz),
                           grad(hchan(h[2], cc), fx, fy1, fz), grad(hchan(h[3], cc), fx1, fy1, fz),
                           grad(hchan(h[4], cc), fx, fy, fz1), grad(hchan(h[5], cc), fx1, fy, fz1),
                           grad(hchan(h[6], cc), fx, fy1, fz1), grad(hchan(h[7], cc), fx1, fy1, fz1),
                           u, v, w);
    }
}
template<class T, int NC>
inline void perlin4(T* out, const PerlinHash& hash, const T& x, const T& y, const T& z,
                    const T& w)
{
    int X, Y, Z, W;
    T fx = floorfrac(x, &X), fy = floorfrac(y, &Y), fz = floorfrac(z, &Z),
      fw = floorfrac(w, &W);
    T u = fade(fx), v = fade(fy), t = fade(fz), s = fade(fw);
    uint32_t h[16];
    for (int k = 0; k < 16; ++k)
        h[k] = hash(X + (k & 1), Y + ((k >> 1) & 1), Z + ((k >> 2) & 1), W + (k >> 3));
    T fx1 = fx - 1.0f, fy1 = fy - 1.0f, fz1 = fz - 1.0f, fw1 = fw - 1.0f;
    for (int c = 0; c < NC; ++c) {
        int cc = NC == 1 ? -1 : c;
        T lo = trilerp(grad(hchan(h[0], cc), fx, fy, fz, fw), grad(hchan(h[1], cc), fx1, fy, fz, fw),
                       grad(hchan(h[2], cc), fx, fy1, fz, fw), grad(hchan(h[3], cc), fx1, fy1, fz, fw),
                       grad(hchan(h[4], cc), fx, fy, fz1, fw), grad(hchan(h[5], cc), fx1, fy, fz1, fw),
                       grad(hchan(h[6], cc), fx, fy1, fz1, fw), grad(hchan(h[7], cc), fx1, fy1, fz1, fw),
                       u, v, t);
        T hi = trilerp(grad(hchan(h[8], cc), fx, fy, fz, fw1), grad(hchan(h[9], cc), fx1, fy, fz, fw1),
                       grad(hchan(h[10], cc), fx, fy1, fz, fw1), grad(hchan(h[11], cc), fx1, fy1, fz, fw1),
                       grad(hchan(h[12], cc), fx, fy, fz1, fw1), grad(hchan(h[13], cc), fx1, fy, fz1, fw1),
                       grad(hchan(h[14], cc), fx, fy1, fz1, fw1), grad(hchan(h[15], cc), fx1, fy1, fz1, fw1),
                       u, v, t);
        out[c] = 0.8344f * lerp(lo, hi, s);
    }
}

// period helper: HashScalarPeriodic ctor (oslnoise.h:1163-1183)
inline int iperiod(float p)
{
    int i = ifloor(p);
    return i < 1 ? 1 : i;
}

// Generic front end.  SIGNED: snoise/psnoise; else noise = 0.5*(s+1)
// (oslnoise.h:2288-2342).  in[] are the 1..4 coordinates, per[] the optional
// integer periods.
template<class T, int NC, bool SIGNED>
inline void perlin_nd(T* out, int dim, const T* in, const int* per)
{
    PerlinHash h(per);
    switch (dim) {
    case 1: perlin1<T, NC>(out, h, in[0]); break;
    case 2: perlin2<T, NC>(out, h, in[0], in[1]); break;
    case 3: perlin3<T, NC>(out, h, in[0], in[1], in[2]); break;
    default: perlin4<T, NC>(out, h, in[0], in[1], in[2], in[3]); break;
    }
    if (!SIGNED)
        for (int c = 0; c < NC; ++c)
            out[c] = 0.5f * (out[c] + 1.0f);
}

}  // namespace oslo
