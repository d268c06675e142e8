// osl_oracle_matrix.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of OSL's matrix shadeops:
//   osl_mul_mmm/_mmf, osl_div_mmm/_mmf/_mfm/_m_ff, osl_transpose_mm,
//   osl_transform*_vmv/_dvmdv, osl_determinant_fm,
//   osl_get_matrix / osl_get_inverse_matrix / osl_prepend_matrix_from /
//   osl_get_from_to_matrix / osl_transform_triple      src/liboslexec/opmatrix.cpp:28-344
//   robust_multVecMatrix, multDirMatrix, det4x4          src/include/OSL/Imathx/Imathx.h:32-58, 336-352, 456-495
//   dual forms                                           src/include/OSL/dual_vec.h:357-395
// Third-party arithmetic restated from its published form: Imath 3.1 Matrix44
// product (row-major i,k,j sums), Matrix44::inverse (affine fast path: 3x3
// cofactors / determinant and the translated row; otherwise Gauss-Jordan with
// partial pivoting, gjInverse), transposed.
#pragma once

namespace oslo {

struct M44 {
    float x[4][4];
    M44() { *this = M44(1.0f); }
    explicit M44(float f)
    {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                x[i][j] = i == j ? f : 0.0f;
    }
    M44(float a, float b, float c, float d, float e, float f, float g, float h, float i, float j, float k, float l,
        float m, float n, float o, float p)
    {
        const float v[16] = { a, b, c, d, e, f, g, h, i, j, k, l, m, n, o, p };
        for (int r = 0; r < 4; ++r)
            for (int q = 0; q < 4; ++q)
                x[r][q] = v[4 * r + q];
    }
    float& operator[](int i) { return (&x[0][0])[i]; }
    const float& operator[](int i) const { return (&x[0][0])[i]; }
};
inline M44 operator*(const M44& a, const M44& b)
{
    M44 r(0.0f);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = a.x[i][0] * b.x[0][j] + a.x[i][1] * b.x[1][j] + a.x[i][2] * b.x[2][j] + a.x[i][3] * b.x[3][j];
    return r;
}
inline M44 operator*(const M44& a, float f)
{
    M44 r(0.0f);
    for (int i = 0; i < 16; ++i)
        r[i] = a[i] * f;
    return r;
}
inline M44 operator*(float f, const M44& a) { return a * f; }
inline M44 operator-(const M44& a)
{
    M44 r(0.0f);
    for (int i = 0; i < 16; ++i)
        r[i] = -a[i];
    return r;
}
inline bool operator==(const M44& a, const M44& b)
{
    for (int i = 0; i < 16; ++i)
        if (a[i] != b[i])
            return false;
    return true;
}
inline M44 m44_transposed(const M44& a)
{
    M44 r(0.0f);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            r.x[i][j] = a.x[j][i];
    return r;
}
// Imath Matrix44::gjInverse (non-throwing form: singular -> identity)
inline M44 m44_gj_inverse(const M44& m)
{
    int i, j, k;
    M44 s;
    M44 t(m);
    // forward elimination
    for (i = 0; i < 3; i++) {
        int pivot      = i;
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
            return M44();
        if (pivot != i) {
            for (j = 0; j < 4; j++) {
                float tmp;
                tmp           = t.x[i][j];
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
    // backward substitution
    for (i = 3; i >= 0; --i) {
        float f;
        if ((f = t.x[i][i]) == 0)
            return M44();
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
// Imath Matrix44::inverse
inline M44 m44_inverse(const M44& m)
{
    const float(*x)[4] = m.x;
    if (x[0][3] != 0 || x[1][3] != 0 || x[2][3] != 0 || x[3][3] != 1)
        return m44_gj_inverse(m);
    M44 s(x[1][1] * x[2][2] - x[2][1] * x[1][2], x[2][1] * x[0][2] - x[0][1] * x[2][2],
          x[0][1] * x[1][2] - x[1][1] * x[0][2], 0,
          x[2][0] * x[1][2] - x[1][0] * x[2][2], x[0][0] * x[2][2] - x[2][0] * x[0][2],
          x[1][0] * x[0][2] - x[0][0] * x[1][2], 0,
          x[1][0] * x[2][1] - x[2][0] * x[1][1], x[2][0] * x[0][1] - x[0][0] * x[2][1],
          x[0][0] * x[1][1] - x[1][0] * x[0][1], 0, 0, 0, 0, 1);
    float r = x[0][0] * s.x[0][0] + x[0][1] * s.x[1][0] + x[0][2] * s.x[2][0];
    if (std::fabs(r) >= 1) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                s.x[i][j] /= r;
    } else {
        float mr = std::fabs(r) / std::numeric_limits<float>::min();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                if (mr > std::fabs(s.x[i][j]))
                    s.x[i][j] /= r;
                else
                    return M44();
            }
    }
    s.x[3][0] = -x[3][0] * s.x[0][0] - x[3][1] * s.x[1][0] - x[3][2] * s.x[2][0];
    s.x[3][1] = -x[3][0] * s.x[0][1] - x[3][1] * s.x[1][1] - x[3][2] * s.x[2][1];
    s.x[3][2] = -x[3][0] * s.x[0][2] - x[3][1] * s.x[1][2] - x[3][2] * s.x[2][2];
    return s;
}
inline float det2x2(float a, float b, float c, float d) { return a * d - b * c; }
inline float det3x3(float a1, float a2, float a3, float b1, float b2, float b3, float c1, float c2, float c3)
{
    return a1 * det2x2(b2, b3, c2, c3) - b1 * det2x2(a2, a3, c2, c3) + c1 * det2x2(a2, a3, b2, b3);
}
inline float m44_determinant(const M44& m)
{
    float a1 = m.x[0][0], b1 = m.x[0][1], c1 = m.x[0][2], d1 = m.x[0][3];
    float a2 = m.x[1][0], b2 = m.x[1][1], c2 = m.x[1][2], d2 = m.x[1][3];
    float a3 = m.x[2][0], b3 = m.x[2][1], c3 = m.x[2][2], d3 = m.x[2][3];
    float a4 = m.x[3][0], b4 = m.x[3][1], c4 = m.x[3][2], d4 = m.x[3][3];
    return a1 * det3x3(b2, b3, b4, c2, c3, c4, d2, d3, d4) - b1 * det3x3(a2, a3, a4, c2, c3, c4, d2, d3, d4)
           + c1 * det3x3(a2, a3, a4, b2, b3, b4, d2, d3, d4) - d1 * det3x3(a2, a3, a4, b2, b3, b4, c2, c3, c4);
}
// point: robust_multVecMatrix; S = float or Df per component
template<class S> inline void m44_transform_point(const M44& M, S& x, S& y, S& z)
{
    S a = x * M.x[0][0] + y * M.x[1][0] + z * M.x[2][0] + M.x[3][0];
    S b = x * M.x[0][1] + y * M.x[1][1] + z * M.x[2][1] + M.x[3][1];
    S c = x * M.x[0][2] + y * M.x[1][2] + z * M.x[2][2] + M.x[3][2];
    S w = x * M.x[0][3] + y * M.x[1][3] + z * M.x[2][3] + M.x[3][3];
    if (val_of(w) != 0.0f) {
        x = a / w;
        y = b / w;
        z = c / w;
    } else {
        x = S(0.0f);
        y = S(0.0f);
        z = S(0.0f);
    }
}
inline V3 m44_transform_dir(const M44& M, const V3& s)
{
    return V3(s.x * M.x[0][0] + s.y * M.x[1][0] + s.z * M.x[2][0], s.x * M.x[0][1] + s.y * M.x[1][1] + s.z * M.x[2][1],
              s.x * M.x[0][2] + s.y * M.x[1][2] + s.z * M.x[2][2]);
}
// vectype: 0 point, 1 vector, 2 normal (osl_transform_vmv / _transformv_ / _transformn_)
inline V3 m44_transform(const M44& M, const V3& v, int vectype)
{
    if (vectype == 0) {
        float x = v.x, y = v.y, z = v.z;
        m44_transform_point(M, x, y, z);
        return V3(x, y, z);
    }
    if (vectype == 1)
        return m44_transform_dir(M, v);
    return m44_transform_dir(m44_transposed(m44_inverse(M)), v);
}
inline Dv m44_transform(const M44& M, const Dv& v, int vectype)
{
    if (vectype == 0) {
        Df x = comp(v, 0), y = comp(v, 1), z = comp(v, 2);
        m44_transform_point(M, x, y, z);
        return make_dv(x, y, z);
    }
    M44 T = vectype == 1 ? M : m44_transposed(m44_inverse(M));
    return Dv(m44_transform_dir(T, v.val), m44_transform_dir(T, v.dx), m44_transform_dir(T, v.dy));
}

// ---- named coordinate systems (osl_get_matrix & friends) -------------------------------
struct NamedTransform {
    const char* name;
    float m[16];
};
struct TransformSet {
    int n = 0;
    const NamedTransform* t = nullptr;  // "shader" and "object" are entries like any renderer-named space
    const char* commonspace_synonym = "world";
};
inline bool xf_get_matrix(const TransformSet& ts, const char* from, M44& r)
{
    if (!std::strcmp(from, "common") || !std::strcmp(from, ts.commonspace_synonym)) {
        r = M44();
        return true;
    }
    for (int i = 0; i < ts.n; ++i)
        if (!std::strcmp(ts.t[i].name, from)) {
            for (int k = 0; k < 16; ++k)
                r[k] = ts.t[i].m[k];
            return true;
        }
    r = M44();
    return false;
}
inline bool xf_get_inverse_matrix(const TransformSet& ts, const char* to, M44& r)
{
    M44 m;
    bool ok = xf_get_matrix(ts, to, m);
    if (!ok || !std::strcmp(to, "common") || !std::strcmp(to, ts.commonspace_synonym)) {
        r = M44();
        return ok;
    }
    r = m44_inverse(m);  // rs_get_inverse_matrix_*: result = M; result.invert()
    return true;
}
inline bool xf_get_from_to_matrix(const TransformSet& ts, const char* from, const char* to, M44& r)
{
    M44 Mfrom, Mto;
    bool ok = xf_get_matrix(ts, from, Mfrom);
    ok &= xf_get_inverse_matrix(ts, to, Mto);
    r = Mfrom * Mto;
    return ok;
}
// osl_transform_triple: identity copy when a space is unknown
template<class T> inline bool xf_transform_triple(const TransformSet& ts, const char* from, const char* to, const T& in, T& out, int vectype)
{
    M44 M;
    bool ok;
    if (!std::strcmp(from, "common"))
        ok = xf_get_inverse_matrix(ts, to, M);
    else if (!std::strcmp(to, "common"))
        ok = xf_get_matrix(ts, from, M);
    else
        ok = xf_get_from_to_matrix(ts, from, to, M);
    out = ok ? m44_transform(M, in, vectype) : in;
    return ok;
}

}  // namespace oslo
