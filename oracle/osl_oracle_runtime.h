// osl_oracle_runtime.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Per-point execution scaffolding for the C++ that oracle/oso2cpp.py emits:
// the scalar ShaderGlobals record (reference: include/OSL/shaderglobals.h:55-146),
// the SoA launch description (same field order as include/osl_b200.h so tests
// feed both sides identical buffers), printf capture (reference journal /
// rs_printfmt path), and the range runner that mirrors testshade's
// one-execute-per-point loop (src/testshade/testshade.cpp:1585-1672).
#pragma once
#include "osl_oracle_ops.h"
#include "osl_oracle_gabor.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "osl_oracle_color.h"
#include "osl_oracle_matrix.h"
#include "osl_oracle_texture.h"

#ifndef OSLO_COLORSPACE
#    define OSLO_COLORSPACE "Rec709" /* ShadingSystem attribute "colorspace" default */
#endif
namespace oslo {
// uniform colour state of the shading system (ShadingStateUniform::m_colorsystem)
inline const ColorSystem& colorsystem()
{
    static const ColorSystem cs = [] {
        ColorSystem c;
        colorsystem_setup(c, OSLO_COLORSPACE);
        return c;
    }();
    return cs;
}
}  // namespace oslo

namespace oslo {

// Field order is the contract shared with include/osl_b200.h (b200_sg_field).
enum SGField {
    SG_P = 0, SG_dPdx, SG_dPdy, SG_dPdz, SG_I, SG_dIdx, SG_dIdy, SG_N, SG_Ng,
    SG_u, SG_dudx, SG_dudy, SG_v, SG_dvdx, SG_dvdy, SG_dPdu, SG_dPdv,
    SG_time, SG_dtime, SG_dPdtime, SG_Ps, SG_dPsdx, SG_dPsdy,
    SG_surfacearea, SG_raytype, SG_flipHandedness, SG_backfacing,
    SG_NFIELDS
};

struct Launch {
    const float* varying[SG_NFIELDS];  // SoA planes (x[n],y[n],z[n]) or NULL
    float uniform[SG_NFIELDS][4];      // used when varying[f] == NULL
    long long plane_stride;            // elements between x/y/z planes
    const int* shadeindex;             // NULL => iota
    void* output_base;                 // renderer output arena
    const void* userdata_base;
    // named coordinate systems of the renderer, "shader" / "object" included
    // (ShaderGlobals::shader2common / object2common + RendererServices::get_matrix)
    int ntransforms;
    const struct NamedTransform* transforms;
    // userdata the renderer supplies per point (RendererServices::get_userdata /
    // SymLocationDesc arena UserData, llvm_instance.cpp:805-970): value of point i at
    // userdata_base + offset + stride * shadeindex, val[,dx,dy] when derivs; an optional
    // int32 plane says whether the point has the value at all
    int nuserdata;
    const struct UserDataDesc* userdata;
};
struct UserDataDesc {
    const char* name;
    int ncomp;    // 1 or 3
    int is_int;
    long long offset, stride;
    int derivs;
    long long valid_offset, valid_stride;   // valid_offset < 0: every point has it
};
inline TransformSet xf_set(const Launch* L)
{
    TransformSet ts;
    if (L) {
        ts.n = L->ntransforms;
        ts.t = L->transforms;
    }
    return ts;
}
inline M44 m44_diag(float f) { return M44(f); }

struct Ctx {
    std::string* out;  // printf capture (NULL: discard)
    // error messages already reported: the shading system prints a repeated error once
    // (attribute "error_repeats" = 0, shadingsys.cpp)
    std::vector<std::string> errseen = {};
};
// ShadingSystem attribute "error_repeats" (testshade --options error_repeats=1): report identical
// messages again
inline int& oracle_error_repeats()
{
    static int v = 0;
    return v;
}

struct Clos;
struct ClosurePool;

struct SG {
    const Clos* Ci    = nullptr;  // closure output (surface shaders)
    ClosurePool* pool = nullptr;  // per-point closure arena (renderer-owned)
    Dv P;
    V3 dPdz;
    Dv I;
    V3 N, Ng;
    Df u, v;
    V3 dPdu, dPdv;
    float time, dtime;
    V3 dPdtime;
    Dv Ps;
    float surfacearea;
    int raytype, flipHandedness, backfacing;
    int shadeindex;
    Ctx* ctx;
};

inline float ldf(const Launch* L, int f, int c, long long i)
{
    return L->varying[f] ? L->varying[f][c * L->plane_stride + i] : L->uniform[f][c];
}
inline V3 ldv(const Launch* L, int f, long long i)
{
    return V3(ldf(L, f, 0, i), ldf(L, f, 1, i), ldf(L, f, 2, i));
}
inline int ldi(const Launch* L, int f, long long i)
{
    return L->varying[f] ? ((const int*)L->varying[f])[i] : (int)f2u(L->uniform[f][0]);
}

// Build the per-point globals record (what testshade's setup_shaderglobals +
// the renderer would have put in ShaderGlobals).
inline void load_sg(SG& sg, const Launch* L, long long i, Ctx* ctx)
{
    sg.P       = Dv(ldv(L, SG_P, i), ldv(L, SG_dPdx, i), ldv(L, SG_dPdy, i));
    sg.dPdz    = ldv(L, SG_dPdz, i);
    sg.I       = Dv(ldv(L, SG_I, i), ldv(L, SG_dIdx, i), ldv(L, SG_dIdy, i));
    sg.N       = ldv(L, SG_N, i);
    sg.Ng      = ldv(L, SG_Ng, i);
    sg.u       = Df(ldf(L, SG_u, 0, i), ldf(L, SG_dudx, 0, i), ldf(L, SG_dudy, 0, i));
    sg.v       = Df(ldf(L, SG_v, 0, i), ldf(L, SG_dvdx, 0, i), ldf(L, SG_dvdy, 0, i));
    sg.dPdu    = ldv(L, SG_dPdu, i);
    sg.dPdv    = ldv(L, SG_dPdv, i);
    sg.time    = ldf(L, SG_time, 0, i);
    sg.dtime   = ldf(L, SG_dtime, 0, i);
    sg.dPdtime = ldv(L, SG_dPdtime, i);
    sg.Ps      = Dv(ldv(L, SG_Ps, i), ldv(L, SG_dPsdx, i), ldv(L, SG_dPsdy, i));
    sg.surfacearea    = ldf(L, SG_surfacearea, 0, i);
    sg.raytype        = ldi(L, SG_raytype, i);
    sg.flipHandedness = ldi(L, SG_flipHandedness, i);
    sg.backfacing     = ldi(L, SG_backfacing, i);
    sg.shadeindex     = L->shadeindex ? L->shadeindex[i] : (int)i;
    sg.ctx            = ctx;
}

// osl_bind_interpolated_param (llvm_instance.cpp:805-970): fetch the value of an interpolated
// ([[ int lockgeom = 0 ]]) parameter from the renderer's userdata; false = not supplied for this
// point (wrong name, wrong type, or masked out): the caller then runs the default / init ops.
inline const float* userdata_ptr(const Launch* L, const SG& sg, const char* name, int ncomp, int is_int, bool* derivs)
{
    for (int k = 0; L && k < L->nuserdata; ++k) {   // (renderer materials run without a Launch: no userdata)
        const UserDataDesc& d = L->userdata[k];
        if (d.ncomp != ncomp || d.is_int != is_int || std::strcmp(d.name, name) != 0)
            continue;
        const char* base = (const char*)L->userdata_base;
        if (d.valid_offset >= 0) {
            int v;
            std::memcpy(&v, base + d.valid_offset + d.valid_stride * (long long)sg.shadeindex, 4);
            if (!v)
                return nullptr;
        }
        *derivs = d.derivs != 0;
        return (const float*)(base + d.offset + d.stride * (long long)sg.shadeindex);
    }
    return nullptr;
}
inline bool bind_userdata(const Launch* L, const SG& sg, const char* name, float& dst)
{
    bool dv;
    const float* p = userdata_ptr(L, sg, name, 1, 0, &dv);
    if (p) dst = p[0];
    return p != nullptr;
}
inline bool bind_userdata(const Launch* L, const SG& sg, const char* name, Df& dst)
{
    bool dv;
    const float* p = userdata_ptr(L, sg, name, 1, 0, &dv);
    if (p) dst = dv ? Df(p[0], p[1], p[2]) : Df(p[0]);
    return p != nullptr;
}
inline bool bind_userdata(const Launch* L, const SG& sg, const char* name, V3& dst)
{
    bool dv;
    const float* p = userdata_ptr(L, sg, name, 3, 0, &dv);
    if (p) dst = V3(p[0], p[1], p[2]);
    return p != nullptr;
}
inline bool bind_userdata(const Launch* L, const SG& sg, const char* name, Dv& dst)
{
    bool dv;
    const float* p = userdata_ptr(L, sg, name, 3, 0, &dv);
    if (p)
        dst = dv ? Dv(V3(p[0], p[1], p[2]), V3(p[3], p[4], p[5]), V3(p[6], p[7], p[8])) : Dv(V3(p[0], p[1], p[2]));
    return p != nullptr;
}
inline bool bind_userdata(const Launch* L, const SG& sg, const char* name, int& dst)
{
    bool dv;
    const float* p = userdata_ptr(L, sg, name, 1, 1, &dv);
    if (p) std::memcpy(&dst, p, 4);
    return p != nullptr;
}
template<class T> inline bool bind_userdata(const Launch*, const SG&, const char*, T&) { return false; }

// ---------------------------------------------------------------------------
// printf capture.  The generator splits the format at gen time and calls one
// helper per conversion, so no varargs cross the boundary.
// ---------------------------------------------------------------------------
inline void pf_lit(SG& sg, const char* s)
{
    if (sg.ctx && sg.ctx->out)
        sg.ctx->out->append(s);
}
inline void pf_fmt(SG& sg, const char* spec, double v)
{
    if (!(sg.ctx && sg.ctx->out))
        return;
    char buf[128];
    snprintf(buf, sizeof buf, spec, v);
    sg.ctx->out->append(buf);
}
inline void pf_fmt(SG& sg, const char* spec, int v)
{
    if (!(sg.ctx && sg.ctx->out))
        return;
    char buf[128];
    snprintf(buf, sizeof buf, spec, v);
    sg.ctx->out->append(buf);
}
inline void pf_fmt(SG& sg, const char* spec, const char* v)
{
    if (!(sg.ctx && sg.ctx->out))
        return;
    char buf[1024];
    snprintf(buf, sizeof buf, spec, v ? v : "");
    sg.ctx->out->append(buf);
}
// error() / warning(): prefix, message, newline; repeated messages are reported once
// (error_repeats = 0; shadingsys.cpp ShadingSystemImpl::error)
inline void report_message(SG& sg, const char* prefix, const std::string& msg)
{
    if (!(sg.ctx && sg.ctx->out))
        return;
    std::string full = std::string(prefix) + msg + "\n";
    if (!oracle_error_repeats()) {
        for (const std::string& s : sg.ctx->errseen)
            if (s == full)
                return;
        sg.ctx->errseen.push_back(full);
    }
    sg.ctx->out->append(full);
}
inline void pf_f(SG& sg, const char* spec, float v) { pf_fmt(sg, spec, (double)v); }
inline void pf_f(SG& sg, const char* spec, const Df& v) { pf_fmt(sg, spec, (double)v.val); }
inline void pf_f(SG& sg, const char* spec, int v) { pf_fmt(sg, spec, (double)v); }
inline void pf_i(SG& sg, const char* spec, int v) { pf_fmt(sg, spec, v); }
inline void pf_i(SG& sg, const char* spec, float v) { pf_fmt(sg, spec, (int)v); }
inline void pf_v(SG& sg, const char* spec, const V3& v)
{
    pf_fmt(sg, spec, (double)v.x);
    pf_lit(sg, " ");
    pf_fmt(sg, spec, (double)v.y);
    pf_lit(sg, " ");
    pf_fmt(sg, spec, (double)v.z);
}
inline void pf_v(SG& sg, const char* spec, const Dv& v) { pf_v(sg, spec, v.val); }
inline void pf_s(SG& sg, const char* spec, const char* v) { pf_fmt(sg, spec, v); }
inline void pf_v(SG& sg, const char* spec, const M44& m)
{
    for (int i = 0; i < 16; ++i) {
        if (i)
            pf_lit(sg, " ");
        pf_fmt(sg, spec, (double)m[i]);
    }
}

// named-space lookups that report unknown names the way the reference does with the
// default unknown_coordsys_error=1 (opmatrix.cpp:129-134, 160-165, 188-196): the message
// goes through the error handler, which testshade prints inline as "ERROR: ..."
inline void xf_unknown(SG& sg, const char* name)
{
    if (!(sg.ctx && sg.ctx->out))
        return;
    std::string msg = std::string("ERROR: Unknown transformation \"") + (name ? name : "") + "\"\n";
    for (const std::string& s : sg.ctx->errseen)
        if (s == msg)
            return;
    sg.ctx->errseen.push_back(msg);
    pf_lit(sg, msg.c_str());
}
inline bool xf_get_matrix_err(SG& sg, const TransformSet& ts, const char* from, M44& r)
{
    bool ok = xf_get_matrix(ts, from, r);
    if (!ok) {
        xf_unknown(sg, from);   // once in osl_get_matrix ...
        xf_unknown(sg, from);   // ... and once more in osl_prepend_matrix_from
    }
    return ok;
}
inline bool xf_get_from_to_matrix_err(SG& sg, const TransformSet& ts, const char* from, const char* to, M44& r)
{
    M44 a, b;
    if (!xf_get_matrix(ts, from, a))
        xf_unknown(sg, from);
    if (!xf_get_inverse_matrix(ts, to, b))
        xf_unknown(sg, to);
    return xf_get_from_to_matrix(ts, from, to, r);
}
template<class T>
inline bool xf_transform_triple_err(SG& sg, const TransformSet& ts, const char* from, const char* to, const T& in, T& out,
                                    int vectype)
{
    M44 t;
    if (std::strcmp(from, "common") && !xf_get_matrix(ts, from, t))
        xf_unknown(sg, from);
    if (std::strcmp(to, "common") && !xf_get_inverse_matrix(ts, to, t))
        xf_unknown(sg, to);
    return xf_transform_triple(ts, from, to, in, out, vectype);
}

inline bool str_eq(const char* a, const char* b)
{
    if (a == b)
        return true;
    if (!a || !b)
        return false;
    return std::strcmp(a, b) == 0;
}

// output placement: output_base + offset + stride*shadeindex
// (llvm_instance.cpp:1807-1848)
inline float* outp(const Launch* L, const SG& sg, long long offset, long long stride)
{
    return (float*)((char*)L->output_base + offset + stride * (long long)sg.shadeindex);
}
inline void wr(float* p, float v) { p[0] = v; }
inline void wr(float* p, int v) { ((int*)p)[0] = v; }
inline void wr(float* p, const V3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
inline void wr(float* p, const Df& v) { p[0] = v.val; }
inline void wr(float* p, const Dv& v) { wr(p, v.val); }
inline void wrd(float* p, float v) { p[0] = v; p[1] = 0; p[2] = 0; }
inline void wrd(float* p, const Df& v) { p[0] = v.val; p[1] = v.dx; p[2] = v.dy; }
inline void wrd(float* p, const V3& v) { wr(p, v); for (int i = 3; i < 9; ++i) p[i] = 0; }
inline void wrd(float* p, const Dv& v) { wr(p, v.val); wr(p + 3, v.dx); wr(p + 6, v.dy); }

}  // namespace oslo
