// osl_oracle_closure.h — CPU ORACLE (test infrastructure, NOT product code).
//
// Closure trees as the reference builds them: ClosureComponent / ClosureMul /
// ClosureAdd PODs (src/include/OSL/oslclosure.h:65-142) bump-allocated from a
// 1 KB per-point pool (src/testshade/render_state.h:27-54), with the
// construction rules of src/liboslexec/opclosure.cpp:18-105 (zero weight ->
// NULL, unit weight -> passthrough).  Closure ids follow
// src/testrender/shading.h:25-60; parameter blocks are stored as consecutive
// 32-bit words in registration order (src/testrender/shading.cpp:194-297).
#pragma once
#include "osl_oracle.h"

namespace oslo {

enum ClosureIDs {
    CL_ADD = -2, CL_MUL = -1, COMPONENT_BASE_ID = 0,
    EMISSION_ID = 1, BACKGROUND_ID, DIFFUSE_ID, OREN_NAYAR_ID, TRANSLUCENT_ID, PHONG_ID, WARD_ID,
    MICROFACET_ID, REFLECTION_ID, FRESNEL_REFLECTION_ID, REFRACTION_ID, TRANSPARENT_ID, DEBUG_ID,
    HOLDOUT_ID, MX_OREN_NAYAR_DIFFUSE_ID, MX_BURLEY_DIFFUSE_ID, MX_DIELECTRIC_ID, MX_CONDUCTOR_ID,
    MX_GENERALIZED_SCHLICK_ID, MX_TRANSLUCENT_ID, MX_TRANSPARENT_ID, MX_SUBSURFACE_ID, MX_SHEEN_ID,
    MX_UNIFORM_EDF_ID, MX_ANISOTROPIC_VDF_ID, MX_MEDIUM_VDF_ID, MX_LAYER_ID, SPI_THINLAYER, EMPTY_ID
};

struct Clos {
    int id;
};
struct ClosComp : Clos {
    V3 w;
    float params[24];  // only the first nparams words are allocated
};
struct ClosMul : Clos {
    V3 weight;
    const Clos* closure;
};
struct ClosAdd : Clos {
    const Clos* a;
    const Clos* b;
};

struct ClosurePool {
    alignas(8) char buf[1024];
    int used = 0;
    void reset() { used = 0; }
    void* alloc(size_t size)
    {
        size_t at = (size_t(used) + 7) & ~size_t(7);
        if (at + size > sizeof buf)
            return nullptr;
        used = int(at + size);
        return buf + at;
    }
};

inline bool is_zero(const V3& w) { return w.x == 0.0f && w.y == 0.0f && w.z == 0.0f; }
inline bool is_one(const V3& w) { return w.x == 1.0f && w.y == 1.0f && w.z == 1.0f; }

// osl_allocate_closure_component / osl_allocate_weighted_closure_component
inline ClosComp* clos_component(ClosurePool* pool, int id, int nparams, const V3* w)
{
    if (w && is_zero(*w))
        return nullptr;
    ClosComp* c = (ClosComp*)pool->alloc(sizeof(Clos) + sizeof(V3) + 4 * size_t(nparams));
    if (c) {
        c->id = id;
        c->w  = w ? *w : V3(1.0f);
    }
    return c;
}
inline const Clos* clos_mul(ClosurePool* pool, const Clos* a, const V3& w)
{
    if (!a || is_zero(w))
        return nullptr;
    if (is_one(w))
        return a;
    ClosMul* m = (ClosMul*)pool->alloc(sizeof(ClosMul));
    if (m) {
        m->id      = CL_MUL;
        m->weight  = w;
        m->closure = a;
    }
    return m;
}
inline const Clos* clos_mul(ClosurePool* pool, const Clos* a, float w)
{
    if (!a || w == 0.0f)
        return nullptr;
    if (w == 1.0f)
        return a;
    return clos_mul(pool, a, V3(w));
}
inline const Clos* clos_add(ClosurePool* pool, const Clos* a, const Clos* b)
{
    if (!a)
        return b;
    if (!b)
        return a;
    ClosAdd* s = (ClosAdd*)pool->alloc(sizeof(ClosAdd));
    if (s) {
        s->id = CL_ADD;
        s->a  = a;
        s->b  = b;
    }
    return s;
}
inline void putp(float* p, float v) { *p = v; }
inline void putp(float* p, int v) { std::memcpy(p, &v, 4); }
inline void putp(float* p, const V3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
// String closure params (only microfacet's `dist` on this path) are stored as a
// small code: the reference compares ustringhash against uh_ggx / uh_beckmann /
// uh_default (shading.cpp:1518-1534); anything else adds no lobe.
inline int closure_string_code(const char* s)
{
    if (!s)
        return 0;
    if (!std::strcmp(s, "ggx"))
        return 1;
    if (!std::strcmp(s, "beckmann"))
        return 2;
    if (!std::strcmp(s, "default"))
        return 3;
    return 0;
}
inline void putp(float* p, const char* s) { putp(p, closure_string_code(s)); }
// closure-typed parameter (layer's top / base): the pointer itself, two words
inline void putp(float* p, const Clos* c) { std::memcpy(p, &c, sizeof c); }

}  // namespace oslo
