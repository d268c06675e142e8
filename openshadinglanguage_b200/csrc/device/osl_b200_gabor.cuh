// osl_b200_gabor.cuh — Gabor noise (sparse convolution, always with derivatives).
//
//   NoiseParams                                src/liboslexec/oslexec_pvt.h:2615-2630
//   fast_rng / kernel / slice / filter / wrap  src/liboslnoise/gabornoise.h:69-230
//   GaborParams, gabor_sample, gabor_cell,
//   gabor_grid, gabor_setup_filter, gabor*,
//   pgabor*                                    src/liboslnoise/gabornoise.cpp:19-414
//   1-D / 2-D slice 3-D, 4-D ignores time      src/liboslexec/opnoise.cpp:484-632
//
// Same operation sequence as the reference's scalar code, reorganised for the
// GPU: the position dual lives in three scalar duals (registers, no indexed
// aggregates), everything that depends only on (options, filter footprint) —
// the Poisson threshold, the filter's 2x2 matrix algebra — is evaluated once
// per call instead of once per impulse, and the 27-cell walk is a flat loop.
// Transcendentals: the reference calls libm expf / sincosf here (not OIIO
// fast_*); on the device these are CUDA's expf / sincosf (<= 2 ulp), so values
// agree with the CPU to a few 1e-7 per impulse, not bit for bit.  The integer
// side (cell hash, LCG, impulse counts, acceptance tests) is exact: the
// Poisson threshold exp(-mean) is taken in double and rounded once.
#pragma once

namespace osld {

struct NoiseParams {
    int anisotropic;
    int do_filter;
    V3 direction;
    float bandwidth;
    float impulses;
};
OSLD NoiseParams noise_params_default()
{
    NoiseParams o;
    o.anisotropic = 0;
    o.do_filter   = 1;
    o.direction   = mkv(1.0f, 0.0f, 0.0f);
    o.bandwidth   = 1.0f;
    o.impulses    = 16.0f;
    return o;
}

namespace gabor {

#define OSLD_TWO_PI_F ((float)(OSLD_PI * 2.0))

struct M22 {
    float a, b, c, d;  // x[0][0], x[0][1], x[1][0], x[1][1]
};
OSLD M22 m22(float a, float b, float c, float d) { M22 m; m.a = a; m.b = b; m.c = c; m.d = d; return m; }
OSLD M22 m22_mul(M22 p, M22 q)
{
    // Imath Matrix22::operator*: tmp = 0; tmp[i][j] += x[i][k] * v[k][j]
    return m22(0.0f + p.a * q.a + p.b * q.c, 0.0f + p.a * q.b + p.b * q.d, 0.0f + p.c * q.a + p.d * q.c,
               0.0f + p.c * q.b + p.d * q.d);
}
OSLD M22 m22_scale(float s, M22 m) { return m22(m.a * s, m.b * s, m.c * s, m.d * s); }
OSLD M22 m22_add(M22 p, M22 q) { return m22(p.a + q.a, p.b + q.b, p.c + q.c, p.d + q.d); }
OSLD float m22_det(M22 m) { return m.a * m.d - m.b * m.c; }
OSLD M22 m22_inverse(M22 m)
{
    // Imath 3.1 Matrix22::inverse(): adjugate / determinant, identity when singular
    M22 s   = m22(m.d, -m.b, -m.c, m.a);
    float r = m.a * m.d - m.c * m.b;
    if (fabsf(r) >= 1) {
        return m22(s.a / r, s.b / r, s.c / r, s.d / r);
    }
    float mr = fabsf(r) / 1.17549435e-38f;
    if (mr > fabsf(s.a) && mr > fabsf(s.b) && mr > fabsf(s.c) && mr > fabsf(s.d))
        return m22(s.a / r, s.b / r, s.c / r, s.d / r);
    return m22(1.0f, 0.0f, 0.0f, 1.0f);
}
OSLD void m22_mulv(M22 m, float vx, float vy, float& ox, float& oy)
{
    ox = vx * m.a + vy * m.c;
    oy = vx * m.b + vy * m.d;
}
OSLD V3 gnormalized(V3 v)
{
    float l = imath_length(v);
    if (l == 0.0f)
        return mkv(0.0f);
    return mkv(v.x / l, v.y / l, v.z / l);
}
OSLD float gclamp(float x, float lo, float hi) { return (x < lo) ? lo : ((x > hi) ? hi : x); }
OSLD Df dexp(Df a)
{
    float f = expf(a.val);
    return chain(a, f, f);
}
OSLD Df dcos(Df a)
{
    float s, c;
    sincosf(a.val, &s, &c);
    return chain(a, c, -s);
}
OSLD float gwrap(float s, float period)
{
    period = floorf(period);
    if (period < 1.0f)
        period = 1.0f;
    return s - period * floorf(s / period);
}

struct Params {
    V3 omega;
    int anisotropic;
    bool do_filter;
    bool periodic;
    float a, weight;
    V3 N;
    // tangent frame (columns t, b, n of Mtex_to_tan)
    V3 lt, lb;
    V3 period;
    float lambda, sqrt_lambda_inv;
    float radius, radius2, radius3, radius_inv;
    float poisson_g;
    // per-call constants of filter_gabor_kernel_2d (they depend only on the filter and a)
    float c_F, k_GF, a_f;
    M22 SGSF_inv, Sigma_GF_Gi;
};

OSLD void params_init(Params& gp, const NoiseParams& opt)
{
    gp.omega       = opt.direction;
    gp.anisotropic = opt.anisotropic;
    gp.do_filter   = opt.do_filter != 0;
    gp.weight      = 1.0f;
    gp.periodic    = false;
    float bandwidth          = gclamp(opt.bandwidth, 0.01f, 100.0f);
    float TWO_to_bandwidth   = fast_exp2(bandwidth);
    // host value of the reference's constant (gabornoise.cpp:52-55)
    const float SQRT_PI_OVER_LN2 = 2.128934e+00f;
    // the reference evaluates this product in double (2.0f * (T - 1.0) / (T + 1.0) * c)
    gp.a = (float)((double)2.0f * (((double)TWO_to_bandwidth - 1.0) / ((double)TWO_to_bandwidth + 1.0))
                   * (double)SQRT_PI_OVER_LN2);
    // -logf(0.02f) as glibc rounds it
    const float NEG_LOG_TRUNCATE = __int_as_float(0x407a5e96);
    gp.radius     = sqrtf(NEG_LOG_TRUNCATE / (float)OSLD_PI) / gp.a;
    gp.radius2    = gp.radius * gp.radius;
    gp.radius3    = gp.radius2 * gp.radius;
    gp.radius_inv = 1.0f / gp.radius;
    float impulses     = gclamp(opt.impulses, 1.0f, 32.0f);
    gp.lambda          = impulses / ((float)(1.33333 * OSLD_PI) * gp.radius3);
    gp.sqrt_lambda_inv = 1.0f / sqrtf(gp.lambda);
    gp.poisson_g       = (float)exp(-(double)(gp.lambda * gp.radius3));
}

// gabor_setup_filter (gabornoise.cpp:251-290) + the impulse-independent part of
// filter_gabor_kernel_2d (gabornoise.h:154-182)
OSLD void setup_filter(Params& gp, V3 Pdx, V3 Pdy)
{
    V3 n = cross3(Pdx, Pdy);
    if (n.x * n.x + n.y * n.y + n.z * n.z < 1.0e-6f) {
        gp.do_filter = false;
        return;
    }
    // make_orthonormals
    n = gnormalized(n);
    V3 t;
    if (fabsf(n.x) < 0.9f)
        t = mkv(0.0f, n.z, -n.y);
    else
        t = mkv(-n.z, 0.0f, n.x);
    t    = gnormalized(t);
    V3 b = cross3(n, t);
    // Mscreen_to_tan = cols(Pdx, Pdy, 0) * cols(t, b, n); only its upper-left 2x2 is used
    //   M[i][j] = S[i][0]*T[0][j] + S[i][1]*T[1][j] + S[i][2]*T[2][j],  S[i] = (Pdx_i, Pdy_i, 0)
    //   T[0] = (t.x, b.x, n.x), T[1] = (t.y, b.y, n.y), T[2] = (t.z, b.z, n.z)
    float m00 = Pdx.x * t.x + Pdy.x * t.y + 0.0f * t.z;
    float m01 = Pdx.x * b.x + Pdy.x * b.y + 0.0f * b.z;
    float m10 = Pdx.y * t.x + Pdy.y * t.y + 0.0f * t.z;
    float m11 = Pdx.y * b.x + Pdy.y * b.y + 0.0f * b.z;
    M22 M_scr_tan   = m22(m00, m01, m10, m11);
    M22 Sigma_f_scr = m22(0.25f, 0.0f, 0.0f, 0.25f);
    M22 M_scr_tan_t = m22(m00, m10, m01, m11);
    M22 Sigma_f_tan = m22_mul(m22_mul(M_scr_tan_t, Sigma_f_scr), M_scr_tan);
    gp.N  = n;
    gp.lt = t;
    gp.lb = b;
    float det_filter = m22_det(Sigma_f_tan);
    if (det_filter < 1.0e-18f) {
        gp.do_filter = false;
        return;
    }
    const float a = gp.a;
    M22 Sigma_G   = m22_scale(a * a / OSLD_TWO_PI_F, m22(1.0f, 0.0f, 0.0f, 1.0f));
    gp.c_F        = 1.0f / (OSLD_TWO_PI_F * sqrtf(m22_det(Sigma_f_tan)));
    M22 Sigma_F   = m22_scale((float)(1.0 / (4.0 * OSLD_PI * OSLD_PI)), m22_inverse(Sigma_f_tan));
    M22 SGSF      = m22_add(Sigma_G, Sigma_F);
    gp.k_GF       = 1.0f / (OSLD_TWO_PI_F * sqrtf(m22_det(SGSF)));
    gp.SGSF_inv   = m22_inverse(SGSF);
    M22 Sigma_G_i = m22_inverse(Sigma_G);
    M22 Sigma_GF  = m22_inverse(m22_add(m22_inverse(Sigma_F), Sigma_G_i));
    gp.Sigma_GF_Gi = m22_mul(Sigma_GF, Sigma_G_i);
    gp.a_f         = sqrtf((float)((OSLD_PI * 2.0) * (double)sqrtf(m22_det(Sigma_GF))));
}

struct Rng {
    u32 seed;
    OSLD float next() { return (float)(seed *= 3039177861u) / 4294967296.0f; }
};

// unfiltered 3-D kernel, weight 1 and constant phase (gabor_kernel, gabornoise.h:113-121)
OSLD Df kernel3(const Params& gp, V3 omega, float phi, Df X, Df Y, Df Z)
{
    Df g = dexp(((float)(-OSLD_PI) * (gp.a * gp.a)) * (X * X + Y * Y + Z * Z));
    Df h = dcos(OSLD_TWO_PI_F * (X * omega.x + Y * omega.y + Z * omega.z) + mkd(phi));
    return mkd(gp.weight) * g * h;
}

OSLD Df cell(const Params& gp, V3 c_i, Df xc, Df yc, Df zc, int seed)
{
    V3 h = c_i;
    if (gp.periodic)
        h = mkv(gwrap(c_i.x, gp.period.x), gwrap(c_i.y, gp.period.y), gwrap(c_i.z, gp.period.z));
    Rng rng;
    rng.seed = inthash((u32)ifloor(h.x), (u32)ifloor(h.y), (u32)ifloor(h.z), (u32)seed);
    if (!rng.seed)
        rng.seed = 1;
    // fast_rng::poisson
    int n_impulses = 0;
    {
        float t = rng.next();
        while (t > gp.poisson_g) {
            ++n_impulses;
            t *= rng.next();
        }
    }
    Df sum = mkd(0.0f);
    for (int i = 0; i < n_impulses; i++) {
        float z_rng = rng.next(), y_rng = rng.next(), x_rng = rng.next();
        Df X = gp.radius * (xc - x_rng), Y = gp.radius * (yc - y_rng), Z = gp.radius * (zc - z_rng);
        // gabor_sample
        V3 omega;
        if (gp.anisotropic == 1) {
            omega = gp.omega;
        } else if (gp.anisotropic == 0) {
            float omega_t     = OSLD_TWO_PI_F * rng.next();
            float ru          = rng.next();
            float cos_omega_p = -1.0f * (1.0f - ru) + 1.0f * ru;
            float sin_omega_p = sqrtf(fmaxf(0.0f, 1.0f - cos_omega_p * cos_omega_p));
            float so, co;
            fast_sincos(omega_t, &so, &co);
            omega = gnormalized(mkv(co * sin_omega_p, so * sin_omega_p, cos_omega_p));
        } else {
            float omega_r = imath_length(gp.omega);
            float omega_t = OSLD_TWO_PI_F * rng.next();
            float so, co;
            fast_sincos(omega_t, &so, &co);
            omega = omega_r * mkv(co, so, 0.0f);
        }
        float phi = OSLD_TWO_PI_F * rng.next();
        if (X.val * X.val + Y.val * Y.val + Z.val * Z.val < gp.radius2) {
            if (!gp.do_filter) {
                sum = sum + kernel3(gp, omega, phi, X, Y, Z);
            } else {
                // impulse anisotropy into tangent space (multMatrix with cols t, b, n)
                float ox = omega.x * gp.lt.x + omega.y * gp.lt.y + omega.z * gp.lt.z;
                float oy = omega.x * gp.lb.x + omega.y * gp.lb.y + omega.z * gp.lb.z;
                float oz = omega.x * gp.N.x + omega.y * gp.N.y + omega.z * gp.N.z;
                // slice_gabor_kernel_3d
                Df d     = -(X * gp.N.x + Y * gp.N.y + Z * gp.N.z);
                Df w_s   = gp.weight * dexp(((float)(-OSLD_PI) * (gp.a * gp.a)) * (d * d));
                Df phi_s = phi - OSLD_TWO_PI_F * d * oz;
                // filter_gabor_kernel_2d, impulse-dependent part
                float tx, ty;
                m22_mulv(gp.SGSF_inv, ox, oy, tx, ty);
                Df w_f = gp.c_F * w_s * gp.k_GF * expf(-0.5f * (tx * ox + ty * oy));
                float fx, fy;
                m22_mulv(gp.Sigma_GF_Gi, ox, oy, fx, fy);
                // position into tangent space, 2-D kernel
                Df xt = X * gp.lt.x + Y * gp.lt.y + Z * gp.lt.z;
                Df yt = X * gp.lb.x + Y * gp.lb.y + Z * gp.lb.z;
                Df g  = dexp(((float)(-OSLD_PI) * (gp.a_f * gp.a_f)) * (xt * xt + yt * yt));
                Df hh = dcos(OSLD_TWO_PI_F * (xt * fx + yt * fy) + phi_s);
                Df gk = w_f * g * hh;
                if (!finitef(gk.val))
                    gk = kernel3(gp, omega, phi, X, Y, Z);
                sum = sum + gk;
            }
        }
    }
    return sum;
}

OSLD Df evaluate(const Params& gp, Df x, Df y, Df z, int seed)
{
    // gabor_evaluate + gabor_grid
    Df gx = x * gp.radius_inv, gy = y * gp.radius_inv, gz = z * gp.radius_inv;
    V3 fl = mkv(floorf(gx.val), floorf(gy.val), floorf(gz.val));
    Df cx = gx - fl.x, cy = gy - fl.y, cz = gz - fl.z;
    Df sum = mkd(0.0f);
#pragma unroll 1
    for (int n = 0; n < 27; ++n) {
        // k outermost, i innermost, each -1..1 (summation order is part of the result)
        int k = n / 9 - 1, j = (n / 3) % 3 - 1, i = n % 3 - 1;
        V3 c  = mkv((float)i, (float)j, (float)k);
        sum   = sum + cell(gp, fl + c, cx - c.x, cy - c.y, cz - c.z, seed);
    }
    return sum * gp.sqrt_lambda_inv;
}

}  // namespace gabor

// NC = 1: gabor / pgabor, NC = 3: gabor3 / pgabor3; period == nullptr: aperiodic
template<int NC> OSLD void gabor_noise(Df* out, Df x, Df y, Df z, const V3* period, const NoiseParams& opt)
{
    gabor::Params gp;
    gabor::params_init(gp, opt);
    if (period) {
        gp.periodic = true;
        gp.period   = *period;
    }
    if (gp.do_filter)
        gabor::setup_filter(gp, mkv(x.dx, y.dx, z.dx), mkv(x.dy, y.dy, z.dy));
    float gabor_variance = 1.0f / (4.0f * sqrtf(2.0f) * (gp.a * gp.a * gp.a));
    float scale          = 1.0f / (3.0f * sqrtf(gabor_variance));
    scale *= 0.5f;
#pragma unroll 1
    for (int c = 0; c < NC; ++c)
        out[c] = gabor::evaluate(gp, x, y, z, c) * scale;
}

}  // namespace osld
