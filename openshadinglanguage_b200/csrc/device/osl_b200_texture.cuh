// osl_b200_texture.cuh — 2-D texture() on the device.
//
//   osl_texture                      src/liboslexec/optexture.cpp:235-310
//   llvm_gen_texture(+_options)      src/liboslexec/llvm_gen.cpp:2480-2830
//   RendererServices::texture        src/liboslexec/rendservices.cpp:166-232
//   (OptiX analogue: tex2DGrad,      src/testrender/cuda/optix_raytracer.cu:203-210)
//
// The reference hands the lookup to OpenImageIO's TextureSystem.  This is the same
// filter written for a warp of independent points: footprint ellipse from the four
// derivatives, aspect clamped at 32, ceil(aspect - 0.3) probes along the major axis
// with Gaussian line weights, each probe a 4x4 B-spline bicubic (or bilinear /
// closest) on level 0.  Images are float4 texels in HBM (16-byte loads, rows
// contiguous: the 16 taps of one probe touch four 64-byte row segments that stay in
// L2 for the neighbouring points); there is no MIP pyramid because the images this
// path sees (light probes) are not MIP-mapped and OIIO does not build one unless asked
// ("automip").  Derivatives of the result are not produced (zero).
//
// Texture file names must be constants by the time the group is compiled (instance
// parameter values are): the code generator gives each a slot in the module's
// `osl_tex_` table, which the host fills after loading the module.
#pragma once

namespace osld {

struct TexDesc {
    const float4* px;  // [h][w], top scanline first; missing channels hold 0, alpha 1
    int w, h, nch, pad_;
};

enum { TEX_BLACK = 0, TEX_CLAMP = 1, TEX_PERIODIC = 2, TEX_MIRROR = 3 };
enum { TEX_CLOSEST = 0, TEX_BILINEAR = 1, TEX_BICUBIC = 2, TEX_SMARTCUBIC = 3 };

struct TexOpt {
    int swrap, twrap, interp;
    float swidth, twidth, sblur, tblur, fill;
};
OSLD TexOpt tex_default_options()
{
    TexOpt o;
    o.swrap = o.twrap = TEX_BLACK;
    o.interp = TEX_SMARTCUBIC;
    o.swidth = o.twidth = 1.0f;
    o.sblur = o.tblur = o.fill = 0.0f;
    return o;
}

OSLD bool tex_wrap(int& c, int n, int mode)
{
    if (mode == TEX_CLAMP) {
        c = c < 0 ? 0 : (c >= n ? n - 1 : c);
        return true;
    }
    if (mode == TEX_PERIODIC) {
        c %= n;
        if (c < 0)
            c += n;
        return true;
    }
    if (mode == TEX_MIRROR) {
        int iter = c / n;
        c -= iter * n;
        bool flip = (iter & 1) != 0;
        if (c < 0) {
            c += n;
            flip = !flip;
        }
        if (flip)
            c = n - 1 - c;
        return true;
    }
    return c >= 0 && c < n;
}

OSLD float tex_floorfrac(float x, int* i)
{
    float f = floorf(x);
    *i      = (int)f;
    return x - f;
}

OSLD void tex_bspline(float* w, float f)
{
    float g = 1.0f - f;
    w[0]    = (1.0f / 6.0f) * g * g * g;
    w[1]    = (2.0f / 3.0f) - 0.5f * f * f * (2.0f - f);
    w[2]    = (2.0f / 3.0f) - 0.5f * g * g * (2.0f - g);
    w[3]    = (1.0f / 6.0f) * f * f * f;
}

// one probe: acc += weight * filtered texel
OSLD void tex_probe(const TexDesc& im, const TexOpt& o, int interp, float s, float t, float weight, V3& acc)
{
    s *= (float)im.w;
    t *= (float)im.h;
    if (interp == TEX_CLOSEST) {
        int si, ti;
        tex_floorfrac(s, &si);
        tex_floorfrac(t, &ti);
        if (tex_wrap(si, im.w, o.swrap) && tex_wrap(ti, im.h, o.twrap)) {
            float4 p = __ldg(im.px + (size_t)ti * im.w + si);
            acc.x += weight * p.x, acc.y += weight * p.y, acc.z += weight * p.z;
        }
        return;
    }
    s -= 0.5f;
    t -= 0.5f;
    int si, ti;
    float sf = tex_floorfrac(s, &si), tf = tex_floorfrac(t, &ti);
    float ws[4], wt[4];
    int first, n;
    if (interp == TEX_BILINEAR) {
        first = 0, n = 2;
        ws[0] = 1.0f - sf, ws[1] = sf;
        wt[0] = 1.0f - tf, wt[1] = tf;
    } else {
        first = -1, n = 4;
        tex_bspline(ws, sf);
        tex_bspline(wt, tf);
    }
    for (int j = 0; j < n; ++j) {
        int tj = ti + first + j;
        if (!tex_wrap(tj, im.h, o.twrap))
            continue;
        V3 row = mkv(0.0f);
        for (int i = 0; i < n; ++i) {
            int sx = si + first + i;
            if (!tex_wrap(sx, im.w, o.swrap))
                continue;
            float4 p = __ldg(im.px + (size_t)tj * im.w + sx);
            row.x += ws[i] * p.x, row.y += ws[i] * p.y, row.z += ws[i] * p.z;
        }
        float wj = weight * wt[j];
        acc.x += wj * row.x, acc.y += wj * row.y, acc.z += wj * row.z;
    }
}

// Filtered lookup of the first min(nchannels, file channels) channels; the rest take `fill`.
OSLD V3 texture_lookup(const TexDesc& im, const TexOpt& o, float s, float t, float dsdx, float dtdx, float dsdy,
                       float dtdy, int nchannels)
{
    if (!im.px)
        return mkv(o.fill);
    dsdx *= o.swidth, dtdx *= o.twidth, dsdy *= o.swidth, dtdy *= o.twidth;
    const float eps = 1.0e-8f, eps2 = eps * eps;
    float dxlen2 = dsdx * dsdx + dtdx * dtdx, dylen2 = dsdy * dsdy + dtdy * dtdy;
    if (dxlen2 < eps2) {
        if (dylen2 < eps2) {
            dsdx = eps, dsdy = 0.0f, dtdx = 0.0f, dtdy = eps;
        } else {
            float scale = eps / sqrtf(dylen2);
            dsdx = dtdy * scale, dtdx = -dsdy * scale;
        }
    } else if (dylen2 < eps2) {
        float scale = eps / sqrtf(dxlen2);
        dsdy = -dtdx * scale, dtdy = dsdx * scale;
    }
    // footprint ellipse (a handful of fp64 operations per lookup, as in the reference's library)
    double A = (double)(dtdx * dtdx) + (double)(dtdy * dtdy);
    double B = -2.0 * (double)(dsdx * dtdx + dsdy * dtdy);
    double C = (double)(dsdx * dsdx) + (double)(dsdy * dsdy);
    double root   = sqrt((A - C) * (A - C) + B * B);
    double Aprime = (A + C - root) * 0.5, Cprime = (A + C + root) * 0.5;
    float fa = (float)Aprime, fc = (float)Cprime;
    float majorlength = fminf(fc > 0.0f ? sqrtf(fc) : 0.0f, 1000.0f);
    float minorlength = fminf(fa > 0.0f ? sqrtf(fa) : 0.0f, 1000.0f);
    float theta       = fast_atan2((float)B, (float)(A - C)) * 0.5f + 1.57079632679489661923f;
    if (o.sblur + o.tblur != 0.0f) {
        float st, ct;
        fast_sincos(theta, &st, &ct);
        st = fabsf(st), ct = fabsf(ct);
        majorlength += o.sblur * ct + o.tblur * st;
        minorlength += o.sblur * st + o.tblur * ct;
    }
    const float maxaniso = 32.0f;
    float aspect = majorlength / minorlength;
    aspect       = aspect < 1.0f ? 1.0f : (aspect > 1.0e6f ? 1.0e6f : aspect);
    if (aspect > maxaniso) {
        aspect      = maxaniso;
        minorlength = majorlength / maxaniso;
    }
    float smajor, tmajor;
    fast_sincos(theta, &tmajor, &smajor);
    float L = 2.0f * (majorlength - minorlength);
    smajor *= L, tmajor *= L;
    int nsamples = (int)ceilf(aspect - 0.3f);
    nsamples     = nsamples < 1 ? 1 : (nsamples > 64 ? 64 : nsamples);
    float invsamples = 1.0f / (float)nsamples;
    int interp = o.interp == TEX_SMARTCUBIC ? TEX_BICUBIC : o.interp;  // level 0
    V3 acc     = mkv(0.0f);
    if (nsamples <= 2) {
        for (int k = 0; k < nsamples; ++k) {
            float pos = ((float)k + 0.5f) * invsamples - 0.5f;
            tex_probe(im, o, interp, s + pos * smajor, t + pos * tmajor, nsamples == 1 ? 1.0f : 0.5f, acc);
        }
    } else {
        // Gaussian line weights, symmetric about the centre; normalised by their sum taken in
        // index order (the weights are recomputed rather than stored: no local array)
        float scale = majorlength / L, sumw = 0.0f;
        for (int k = 0; k < nsamples; ++k) {
            int i   = k < nsamples - 1 - k ? k : nsamples - 1 - k;
            float x = (2.0f * ((float)i + 0.5f) * invsamples - 1.0f) * scale;
            sumw += fast_exp(-2.0f * x * x);
        }
        for (int k = 0; k < nsamples; ++k) {
            int i     = k < nsamples - 1 - k ? k : nsamples - 1 - k;
            float x   = (2.0f * ((float)i + 0.5f) * invsamples - 1.0f) * scale;
            float pos = ((float)k + 0.5f) * invsamples - 0.5f;
            tex_probe(im, o, interp, s + pos * smajor, t + pos * tmajor, fast_exp(-2.0f * x * x) / sumw, acc);
        }
    }
    int nc = nchannels < im.nch ? nchannels : im.nch;
    if (nc < 1) acc.x = o.fill;
    if (nc < 2) acc.y = o.fill;
    if (nc < 3) acc.z = o.fill;
    return acc;
}

}  // namespace osld
