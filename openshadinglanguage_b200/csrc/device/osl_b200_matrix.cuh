// osl_b200_matrix.cuh — matrix shadeops (product code, host + device).
//
//   osl_mul_mmm/_mmf, osl_div_mmm/_mmf/_mfm/_m_ff, osl_transpose_mm, osl_determinant_fm,
//   osl_transform{,v,n}_vmv/_dvmdv, osl_prepend_matrix_from, osl_get_from_to_matrix,
//   osl_transform_triple                              src/liboslexec/opmatrix.cpp:28-344
//   robust_multVecMatrix, multDirMatrix, det4x4        src/include/OSL/Imathx/Imathx.h:32-58, 336-352, 456-495
// Imath 3.1 Matrix44 product / inverse (affine fast path, else Gauss-Jordan) restated
// from the published form.  The M44 part compiles for the host too: the launcher
// inverts the renderer's named transforms once per launch with the same arithmetic
// the reference applies per call (rs_get_inverse_matrix_*: copy, then invert()).
#pragma once

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#    define OSLM __host__ __device__ __forceinline__
#else
#    define OSLM inline
#endif

namespace osld {

struct M44 {
    float x[4][4];
};
OSLM M44 m44_diag(float f)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = i == j ? f : 0.0f;
    return r;
}
OSLM M44 m44_make(float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float k,
                  float l, float m, float n, float o, float p)
{
    M44 r;
    r.x[0][0] = a; r.x[0][1] = b; r.x[0][2] = c; r.x[0][3] = d;
    r.x[1][0] = e; r.x[1][1] = f; r.x[1][2] = g; r.x[1][3] = h;
    r.x[2][0] = i; r.x[2][1] = j; r.x[2][2] = k; r.x[2][3] = l;
    r.x[3][0] = m; r.x[3][1] = n; r.x[3][2] = o; r.x[3][3] = p;
    return r;
}
OSLM M44 m44_load(const float* p)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = p[4 * i + j];
    return r;
}
OSLM M44 operator*(const M44& a, const M44& b)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = a.x[i][0] * b.x[0][j] + a.x[i][1] * b.x[1][j] + a.x[i][2] * b.x[2][j] + a.x[i][3] * b.x[3][j];
    return r;
}
OSLM M44 operator*(const M44& a, float f)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = a.x[i][j] * f;
    return r;
}
OSLM M44 operator*(float f, const M44& a) { return a * f; }
OSLM M44 operator-(const M44& a)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = -a.x[i][j];
    return r;
}
OSLM bool operator==(const M44& a, const M44& b)
{
    bool eq = true;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            eq = eq && (a.x[i][j] == b.x[i][j]);
    return eq;
}
OSLM M44 m44_transposed(const M44& a)
{
    M44 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = a.x[j][i];
    return r;
}
// Imath Matrix44::gjInverse, non-throwing form (singular -> identity)
OSLM M44 m44_gj_inverse(const M44& m)
{
    int i, j, k;
    M44 s = m44_diag(1.0f);
    M44 t = m;
    for (i = 0; i < 3; i++) {
        int pivot       = i;
        float pivotsize = t.x[i][i];
        if (pivotsize < 0)
            pivotsize = -pivotsize;
        for (j = i + 1; j < 4; j++) {
            float tmp = t.x[j][i];
            if (tmp < 0)
                tmp = -tmp;
            if (tmp > pivotsize) {
                pivot     = j;
                pivotsize = tmp;
            }
        }
        if (pivotsize == 0)
            return m44_diag(1.0f);
        if (pivot != i) {
            for (j = 0; j < 4; j++) {
                float tmp     = t.x[i][j];
                t.x[i][j]     = t.x[pivot][j];
                t.x[pivot][j] = tmp;
                tmp           = s.x[i][j];
                s.x[i][j]     = s.x[pivot][j];
                s.x[pivot][j] = tmp;
            }
        }
        for (j = i + 1; j < 4; j++) {
            float f = t.x[j][i] / t.x[i][i];
            for (k = 0; k < 4; k++) {
                t.x[j][k] -= f * t.x[i][k];
                s.x[j][k] -= f * s.x[i][k];
            }
        }
    }
    for (i = 3; i >= 0; --i) {
        float f = t.x[i][i];
        if (f == 0)
            return m44_diag(1.0f);
        for (j = 0; j < 4; j++) {
            t.x[i][j] /= f;
            s.x[i][j] /= f;
        }
        for (j = 0; j < i; j++) {
            f = t.x[j][i];
            for (k = 0; k < 4; k++) {
                t.x[j][k] -= f * t.x[i][k];
                s.x[j][k] -= f * s.x[i][k];
            }
        }
    }
    return s;
}
// Imath Matrix44::inverse: affine fast path (3x3 cofactors + translated row), else Gauss-Jordan
OSLM M44 m44_inverse(const M44& m)
{
    const float(*x)[4] = m.x;
    if (x[0][3] != 0 || x[1][3] != 0 || x[2][3] != 0 || x[3][3] != 1)
        return m44_gj_inverse(m);
    M44 s = m44_make(x[1][1] * x[2][2] - x[2][1] * x[1][2], x[2][1] * x[0][2] - x[0][1] * x[2][2],
                     x[0][1] * x[1][2] - x[1][1] * x[0][2], 0,
                     x[2][0] * x[1][2] - x[1][0] * x[2][2], x[0][0] * x[2][2] - x[2][0] * x[0][2],
                     x[1][0] * x[0][2] - x[0][0] * x[1][2], 0,
                     x[1][0] * x[2][1] - x[2][0] * x[1][1], x[2][0] * x[0][1] - x[0][0] * x[2][1],
                     x[0][0] * x[1][1] - x[1][0] * x[0][1], 0, 0, 0, 0, 1);
    float r = x[0][0] * s.x[0][0] + x[0][1] * s.x[1][0] + x[0][2] * s.x[2][0];
    if (fabsf(r) >= 1) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s.x[i][j] /= r;
    } else {
        float mr = fabsf(r) / 1.17549435e-38f;
        bool ok  = true;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                ok = ok && (mr > fabsf(s.x[i][j]));
        if (!ok)
            return m44_diag(1.0f);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s.x[i][j] /= r;
    }
    s.x[3][0] = -x[3][0] * s.x[0][0] - x[3][1] * s.x[1][0] - x[3][2] * s.x[2][0];
    s.x[3][1] = -x[3][0] * s.x[0][1] - x[3][1] * s.x[1][1] - x[3][2] * s.x[2][1];
    s.x[3][2] = -x[3][0] * s.x[0][2] - x[3][1] * s.x[1][2] - x[3][2] * s.x[2][2];
    return s;
}
OSLM float m44_det2(float a, float b, float c, float d) { return a * d - b * c; }
OSLM float m44_det3(float a1, float a2, float a3, float b1, float b2, float b3, float c1, float c2, float c3)
{
    return a1 * m44_det2(b2, b3, c2, c3) - b1 * m44_det2(a2, a3, c2, c3) + c1 * m44_det2(a2, a3, b2, b3);
}
OSLM float m44_determinant(const M44& m)
{
    float a1 = m.x[0][0], b1 = m.x[0][1], c1 = m.x[0][2], d1 = m.x[0][3];
    float a2 = m.x[1][0], b2 = m.x[1][1], c2 = m.x[1][2], d2 = m.x[1][3];
    float a3 = m.x[2][0], b3 = m.x[2][1], c3 = m.x[2][2], d3 = m.x[2][3];
    float a4 = m.x[3][0], b4 = m.x[3][1], c4 = m.x[3][2], d4 = m.x[3][3];
    return a1 * m44_det3(b2, b3, b4, c2, c3, c4, d2, d3, d4) - b1 * m44_det3(a2, a3, a4, c2, c3, c4, d2, d3, d4)
           + c1 * m44_det3(a2, a3, a4, b2, b3, b4, d2, d3, d4) - d1 * m44_det3(a2, a3, a4, b2, b3, b4, c2, c3, c4);
}

#if defined(__CUDACC_RTC__) || defined(OSLD)
// ---- device-only part: transforms of V3 / Dv (needs the device library's types) ----------
OSLD float mvof(float a) { return a; }
OSLD float mvof(Df a) { return a.val; }
OSLD void mzero(float& a) { a = 0.0f; }
OSLD void mzero(Df& a) { a = mkd(0.0f); }
template<class S> OSLD void m44_transform_point(const M44& M, S& x, S& y, S& z)
{
    S a = x * M.x[0][0] + y * M.x[1][0] + z * M.x[2][0] + M.x[3][0];
    S b = x * M.x[0][1] + y * M.x[1][1] + z * M.x[2][1] + M.x[3][1];
    S c = x * M.x[0][2] + y * M.x[1][2] + z * M.x[2][2] + M.x[3][2];
    S w = x * M.x[0][3] + y * M.x[1][3] + z * M.x[2][3] + M.x[3][3];
    if (mvof(w) != 0.0f) {
        x = a / w;
        y = b / w;
        z = c / w;
    } else {
        mzero(x);
        mzero(y);
        mzero(z);
    }
}
OSLD V3 m44_transform_dir(const M44& M, V3 s)
{
    return mkv(s.x * M.x[0][0] + s.y * M.x[1][0] + s.z * M.x[2][0], s.x * M.x[0][1] + s.y * M.x[1][1] + s.z * M.x[2][1],
               s.x * M.x[0][2] + s.y * M.x[1][2] + s.z * M.x[2][2]);
}
// vectype: 0 point, 1 vector, 2 normal
OSLD V3 m44_transform(const M44& M, V3 v, int vectype)
{
    if (vectype == 0) {
        m44_transform_point(M, v.x, v.y, v.z);
        return v;
    }
    if (vectype == 1)
        return m44_transform_dir(M, v);
    return m44_transform_dir(m44_transposed(m44_inverse(M)), v);
}
OSLD Dv m44_transform(const M44& M, const Dv& v, int vectype)
{
    if (vectype == 0) {
        Df x = getc(v, 0), y = getc(v, 1), z = getc(v, 2);
        m44_transform_point(M, x, y, z);
        Dv r;
        setc(r, 0, x);
        setc(r, 1, y);
        setc(r, 2, z);
        return r;
    }
    const M44 T = vectype == 1 ? M : m44_transposed(m44_inverse(M));
    return mkdv(m44_transform_dir(T, v.val), m44_transform_dir(T, v.dx), m44_transform_dir(T, v.dy));
}
#endif

}  // namespace osld
