// osl_oracle_gabor.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Restatement of OSL's Gabor noise (sparse convolution, always with
// derivatives):
//   NoiseParams                       src/liboslexec/oslexec_pvt.h:2615-2630
//   fast_rng, gabor_kernel, slice /
//   filter helpers, wrap,
//   make_orthonormals                 src/liboslnoise/gabornoise.h:69-230
//   GaborParams, gabor_sample,
//   gabor_cell, gabor_grid,
//   gabor_setup_filter, gabor/gabor3,
//   pgabor/pgabor3                    src/liboslnoise/gabornoise.cpp:19-414
//   wrappers (1-D/2-D slice 3-D, 4-D
//   ignores time)                     src/liboslexec/opnoise.cpp:484-637
// Third-party arithmetic restated from its published form: Imath 3.1
// Matrix22::inverse (adjugate / determinant with the singular-matrix guard),
// Matrix22/33 products (row-major i,j,k order), Vec3::normalize.
// libm expf / logf / sincosf are the host's (the reference calls std::exp,
// OIIO::sincos = sincosf); exp2 is OIIO::fast_exp2 (OSL_FAST_MATH=1 default,
// CMakeLists.txt:147-149).
#pragma once
#include <climits>

namespace oslo {

struct NoiseParams {
    int anisotropic = 0;
    int do_filter   = 1;
    V3 direction    = V3(1.0f, 0.0f, 0.0f);
    float bandwidth = 1.0f;
    float impulses  = 16.0f;
};

namespace gabor_impl {

const float Gabor_Frequency      = 2.0f;
const float Gabor_Impulse_Weight = 1.0f;
const float Gabor_Truncate       = 0.02f;
const double TWO_PI_D            = M_PI * 2.0;

struct M22 {
    float x[2][2];
    M22() { x[0][0] = 1; x[0][1] = 0; x[1][0] = 0; x[1][1] = 1; }
    M22(float a, float b, float c, float d) { x[0][0] = a; x[0][1] = b; x[1][0] = c; x[1][1] = d; }
    M22 transposed() const { return M22(x[0][0], x[1][0], x[0][1], x[1][1]); }
    M22 inverse() const
    {
        M22 s(x[1][1], -x[0][1], -x[1][0], x[0][0]);
        float r = x[0][0] * x[1][1] - x[1][0] * x[0][1];
        if (std::fabs(r) >= 1) {
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    s.x[i][j] /= r;
        } else {
            float mr = std::fabs(r) / std::numeric_limits<float>::min();
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j) {
                    if (mr > std::fabs(s.x[i][j]))
                        s.x[i][j] /= r;
                    else
                        return M22();
                }
        }
        return s;
    }
};
inline M22 operator*(const M22& a, const M22& b)
{
    M22 t(0, 0, 0, 0);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k)
                t.x[i][j] += a.x[i][k] * b.x[k][j];
    return t;
}
inline M22 operator*(float s, const M22& a) { return M22(a.x[0][0] * s, a.x[0][1] * s, a.x[1][0] * s, a.x[1][1] * s); }
inline M22 operator+(const M22& a, const M22& b)
{
    return M22(a.x[0][0] + b.x[0][0], a.x[0][1] + b.x[0][1], a.x[1][0] + b.x[1][0], a.x[1][1] + b.x[1][1]);
}
inline float determinant(const M22& M) { return M.x[0][0] * M.x[1][1] - M.x[0][1] * M.x[1][0]; }
struct M33 {
    float x[3][3];
};
inline M33 matrix33_cols(const V3& a, const V3& b, const V3& c)
{
    M33 m;
    m.x[0][0] = a.x; m.x[0][1] = b.x; m.x[0][2] = c.x;
    m.x[1][0] = a.y; m.x[1][1] = b.y; m.x[1][2] = c.y;
    m.x[2][0] = a.z; m.x[2][1] = b.z; m.x[2][2] = c.z;
    return m;
}
inline M33 operator*(const M33& a, const M33& b)
{
    M33 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            t.x[i][j] = a.x[i][0] * b.x[0][j] + a.x[i][1] * b.x[1][j] + a.x[i][2] * b.x[2][j];
    return t;
}
inline V3 mult_matrix(const M33& M, const V3& s)
{
    return V3(s.x * M.x[0][0] + s.y * M.x[1][0] + s.z * M.x[2][0], s.x * M.x[0][1] + s.y * M.x[1][1] + s.z * M.x[2][1],
              s.x * M.x[0][2] + s.y * M.x[1][2] + s.z * M.x[2][2]);
}
inline Dv mult_matrix(const M33& M, const Dv& src)
{
    Df s0 = comp(src, 0), s1 = comp(src, 1), s2 = comp(src, 2);
    Df a = s0 * M.x[0][0] + s1 * M.x[1][0] + s2 * M.x[2][0];
    Df b = s0 * M.x[0][1] + s1 * M.x[1][1] + s2 * M.x[2][1];
    Df c = s0 * M.x[0][2] + s1 * M.x[1][2] + s2 * M.x[2][2];
    return make_dv(a, b, c);
}
struct V2 {
    float x, y;
};
inline V2 mul_m22_v2(const M22& m, const V2& v)
{
    return V2 { v.x * m.x[0][0] + v.y * m.x[1][0], v.x * m.x[0][1] + v.y * m.x[1][1] };
}
inline float gclamp(float x, float lo, float hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
inline V3 gnormalized(const V3& v)
{
    float l = imath_length(v);
    if (l == 0.0f)
        return V3(0.0f);
    return V3(v.x / l, v.y / l, v.z / l);
}
inline Df dexp(const Df& a)
{
    float f = std::exp(a.val);
    return dualfunc(a, f, f);
}
inline Df dcos(const Df& a)
{
    float s, c;
    sincosf(a.val, &s, &c);
    return dualfunc(a, c, -s);
}
inline Df ddot(const Dv& a, const Dv& b)
{
    return comp(a, 0) * comp(b, 0) + comp(a, 1) * comp(b, 1) + comp(a, 2) * comp(b, 2);
}
inline Df ddot(const Dv& a, const V3& b) { return comp(a, 0) * b.x + comp(a, 1) * b.y + comp(a, 2) * b.z; }

struct fast_rng {
    uint32_t m_seed;
    fast_rng(const V3& p, int seed)
    {
        m_seed = inthash((uint32_t)ifloor(p.x), (uint32_t)ifloor(p.y), (uint32_t)ifloor(p.z), (uint32_t)seed);
        if (!m_seed)
            m_seed = 1;
    }
    float operator()() { return (m_seed *= 3039177861u) / float(UINT_MAX); }
    int poisson(float mean)
    {
        float g     = expf(-mean);
        unsigned em = 0;
        float t     = (*this)();
        while (t > g) {
            ++em;
            t *= (*this)();
        }
        return (int)em;
    }
};

// 3-D kernel (Equation 1)
inline Df gabor_kernel(const Df& weight, const V3& omega, const Df& phi, float bandwidth, const Dv& x)
{
    Df g = dexp(float(-M_PI) * (bandwidth * bandwidth) * ddot(x, x));
    Df h = dcos(float(TWO_PI_D) * ddot(x, omega) + phi);
    return weight * g * h;
}
// 2-D kernel: x given as two duals
inline Df gabor_kernel2(const Df& weight, const V2& omega, const Df& phi, float bandwidth, const Df& xx, const Df& xy)
{
    Df g = dexp(float(-M_PI) * (bandwidth * bandwidth) * (xx * xx + xy * xy));
    Df h = dcos(float(TWO_PI_D) * (xx * omega.x + xy * omega.y) + phi);
    return weight * g * h;
}
inline void slice_gabor_kernel_3d(const Df& d, float w, float a, const V3& omega, float phi, Df& w_s, V2& omega_s,
                                  Df& phi_s)
{
    w_s       = w * dexp(float(-M_PI) * (a * a) * (d * d));
    omega_s.x = omega.x;
    omega_s.y = omega.y;
    phi_s     = phi - float(TWO_PI_D) * d * omega.z;
}
inline void filter_gabor_kernel_2d(const M22& filter, const Df& w, float a, const V2& omega, const Df& phi, Df& w_f,
                                   float& a_f, V2& omega_f, Df& phi_f)
{
    M22 Sigma_f = filter;
    Df c_G      = w;
    V2 mu_G     = omega;
    M22 Sigma_G = (a * a / float(TWO_PI_D)) * M22();
    float c_F   = 1.0f / (float(TWO_PI_D) * std::sqrt(determinant(Sigma_f)));
    M22 Sigma_F = float(1.0 / (4.0 * M_PI * M_PI)) * Sigma_f.inverse();
    M22 Sigma_G_Sigma_F = Sigma_G + Sigma_F;
    V2 t                = mul_m22_v2(Sigma_G_Sigma_F.inverse(), mu_G);
    Df c_GF = c_F * c_G * (1.0f / (float(TWO_PI_D) * std::sqrt(determinant(Sigma_G_Sigma_F))))
              * expf(-0.5f * (t.x * mu_G.x + t.y * mu_G.y));
    M22 Sigma_G_i   = Sigma_G.inverse();
    M22 Sigma_GF    = (Sigma_F.inverse() + Sigma_G_i).inverse();
    M22 Sigma_GF_Gi = Sigma_GF * Sigma_G_i;
    V2 mu_GF        = mul_m22_v2(Sigma_GF_Gi, mu_G);
    w_f             = c_GF;
    a_f             = std::sqrt(float(TWO_PI_D * (double)std::sqrt(determinant(Sigma_GF))));
    omega_f         = mu_GF;
    phi_f           = phi;
}
inline float gwrap(float s, float period)
{
    period = std::floor(period);
    if (period < 1.0f)
        period = 1.0f;
    return s - period * std::floor(s / period);
}
inline V3 gwrap(const V3& s, const V3& p) { return V3(gwrap(s.x, p.x), gwrap(s.y, p.y), gwrap(s.z, p.z)); }
inline void make_orthonormals(V3& v, V3& a, V3& b)
{
    v = gnormalized(v);
    if (std::fabs(v.x) < 0.9f)
        a = V3(0.0f, v.z, -v.y);
    else
        a = V3(-v.z, 0.0f, v.x);
    a = gnormalized(a);
    b = cross(v, a);
}

struct GaborParams {
    V3 omega;
    int anisotropic;
    bool do_filter;
    float a;
    float weight;
    V3 N;
    M22 filter;
    M33 local;
    float det_filter = 0.0f;
    float bandwidth;
    bool periodic;
    V3 period;
    float lambda;
    float sqrt_lambda_inv;
    float radius, radius2, radius3, radius_inv;

    GaborParams(const NoiseParams& opt)
        : omega(opt.direction), anisotropic(opt.anisotropic), do_filter(opt.do_filter != 0),
          weight(Gabor_Impulse_Weight), bandwidth(gclamp(opt.bandwidth, 0.01f, 100.0f)), periodic(false)
    {
        float TWO_to_bandwidth            = fast_exp2(bandwidth);
        const float SQRT_PI_OVER_LN2      = 2.128934e+00f;
        a = Gabor_Frequency * ((TWO_to_bandwidth - 1.0) / (TWO_to_bandwidth + 1.0)) * SQRT_PI_OVER_LN2;
        radius     = std::sqrt(-logf(Gabor_Truncate) / float(M_PI)) / a;
        radius2    = radius * radius;
        radius3    = radius2 * radius;
        radius_inv = 1.0f / radius;
        float impulses  = gclamp(opt.impulses, 1.0f, 32.0f);
        lambda          = impulses / (float(1.33333 * M_PI) * radius3);
        sqrt_lambda_inv = 1.0f / std::sqrt(lambda);
    }
};

inline void gabor_sample(GaborParams& gp, fast_rng& rng, V3& omega, float& phi)
{
    if (gp.anisotropic == 1) {
        omega = gp.omega;
    } else if (gp.anisotropic == 0) {
        float omega_t     = float(TWO_PI_D) * rng();
        float cos_omega_p = lerp(-1.0f, 1.0f, rng());
        float sin_omega_p = std::sqrt(std::max(0.0f, 1.0f - cos_omega_p * cos_omega_p));
        float sin_omega_t, cos_omega_t;
        fast_sincos(omega_t, &sin_omega_t, &cos_omega_t);
        omega = gnormalized(V3(cos_omega_t * sin_omega_p, sin_omega_t * sin_omega_p, cos_omega_p));
    } else {
        float omega_r = imath_length(gp.omega);
        float omega_t = float(TWO_PI_D) * rng();
        float sin_omega_t, cos_omega_t;
        fast_sincos(omega_t, &sin_omega_t, &cos_omega_t);
        omega = omega_r * V3(cos_omega_t, sin_omega_t, 0.0f);
    }
    phi = float(TWO_PI_D) * rng();
}

inline Df gabor_cell(GaborParams& gp, const V3& c_i, const Dv& x_c_i, int seed)
{
    fast_rng rng(gp.periodic ? gwrap(c_i, gp.period) : c_i, seed);
    int n_impulses = rng.poisson(gp.lambda * gp.radius3);
    Df sum(0.0f);
    for (int i = 0; i < n_impulses; i++) {
        float z_rng = rng(), y_rng = rng(), x_rng = rng();
        V3 x_i_c(x_rng, y_rng, z_rng);
        Dv x_k_i = gp.radius * (x_c_i - x_i_c);
        float phi_i;
        V3 omega_i;
        gabor_sample(gp, rng, omega_i, phi_i);
        const V3& xv = x_k_i.val;
        if (xv.x * xv.x + xv.y * xv.y + xv.z * xv.z < gp.radius2) {
            if (!gp.do_filter) {
                sum = sum + gabor_kernel(Df(gp.weight), omega_i, Df(phi_i), gp.a, x_k_i);
            } else {
                V3 omega_i_t = mult_matrix(gp.local, omega_i);
                Df d_i       = -ddot(x_k_i, gp.N);
                Df w_i_t_s, phi_i_t_s;
                V2 omega_i_t_s;
                slice_gabor_kernel_3d(d_i, gp.weight, gp.a, omega_i_t, phi_i, w_i_t_s, omega_i_t_s, phi_i_t_s);
                Df w_i_t_s_f, phi_i_t_s_f;
                float a_i_t_s_f;
                V2 omega_i_t_s_f;
                filter_gabor_kernel_2d(gp.filter, w_i_t_s, gp.a, omega_i_t_s, phi_i_t_s, w_i_t_s_f, a_i_t_s_f,
                                       omega_i_t_s_f, phi_i_t_s_f);
                Dv xkit = mult_matrix(gp.local, x_k_i);
                Df gk   = gabor_kernel2(w_i_t_s_f, omega_i_t_s_f, phi_i_t_s_f, a_i_t_s_f, comp(xkit, 0), comp(xkit, 1));
                if (!std::isfinite(gk.val))
                    gk = gabor_kernel(Df(gp.weight), omega_i, Df(phi_i), gp.a, x_k_i);
                sum = sum + gk;
            }
        }
    }
    return sum;
}

inline Df gabor_grid(GaborParams& gp, const Dv& x_g, int seed)
{
    V3 floor_x_g(std::floor(x_g.val.x), std::floor(x_g.val.y), std::floor(x_g.val.z));
    Dv x_c = x_g - floor_x_g;
    Df sum(0.0f);
    for (int k = -1; k <= 1; k++)
        for (int j = -1; j <= 1; j++)
            for (int i = -1; i <= 1; i++) {
                V3 c((float)i, (float)j, (float)k);
                V3 c_i   = floor_x_g + c;
                Dv x_c_i = x_c - c;
                sum      = sum + gabor_cell(gp, c_i, x_c_i, seed);
            }
    return sum * gp.sqrt_lambda_inv;
}
inline Df gabor_evaluate(GaborParams& gp, const Dv& x, int seed)
{
    Dv x_g = x * gp.radius_inv;
    return gabor_grid(gp, x_g, seed);
}
inline void gabor_setup_filter(const Dv& P, GaborParams& gp)
{
    V3 n, t, b;
    n = cross(P.dx, P.dy);
    if (dot(n, n) < 1.0e-6f) {
        gp.do_filter = false;
        return;
    }
    make_orthonormals(n, t, b);
    M33 Mtex_to_tan    = matrix33_cols(t, b, n);
    M33 Mscreen_to_tex = matrix33_cols(P.dx, P.dy, V3(0.0f, 0.0f, 0.0f));
    M33 Mscreen_to_tan = Mscreen_to_tex * Mtex_to_tan;
    M22 M_scr_tan(Mscreen_to_tan.x[0][0], Mscreen_to_tan.x[0][1], Mscreen_to_tan.x[1][0], Mscreen_to_tan.x[1][1]);
    float sigma_f_scr = 0.5f;
    M22 Sigma_f_scr(sigma_f_scr * sigma_f_scr, 0.0f, 0.0f, sigma_f_scr * sigma_f_scr);
    M22 M_scr_tan_t = M_scr_tan.transposed();
    M22 Sigma_f_tan = M_scr_tan_t * Sigma_f_scr * M_scr_tan;
    gp.N          = n;
    gp.filter     = Sigma_f_tan;
    gp.det_filter = determinant(Sigma_f_tan);
    gp.local      = Mtex_to_tan;
    if (gp.det_filter < 1.0e-18f)
        gp.do_filter = false;
}
inline float gabor_scale(const GaborParams& gp)
{
    float gabor_variance = 1.0f / (4.0f * std::sqrt(2.0f) * (gp.a * gp.a * gp.a));
    float scale          = 1.0f / (3.0f * std::sqrt(gabor_variance));
    scale *= 0.5f;
    return scale;
}

}  // namespace gabor_impl

// NC = 1: gabor / pgabor; NC = 3: gabor3 / pgabor3.  period == nullptr: aperiodic.
template<int NC> inline void gabor_noise(Df* out, const Dv& P, const V3* period, const NoiseParams& opt)
{
    using namespace gabor_impl;
    GaborParams gp(opt);
    if (period) {
        gp.periodic = true;
        gp.period   = *period;
    }
    if (gp.do_filter)
        gabor_setup_filter(P, gp);
    Df r[3];
    for (int c = 0; c < NC; ++c)
        r[c] = gabor_evaluate(gp, P, c);
    float scale = gabor_scale(gp);
    for (int c = 0; c < NC; ++c)
        out[c] = r[c] * scale;
}

}  // namespace oslo
