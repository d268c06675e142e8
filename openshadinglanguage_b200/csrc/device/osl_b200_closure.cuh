// osl_b200_closure.cuh — closure construction on the device (product code).
//
// Replaces osl_allocate[_weighted]_closure_component, osl_add_closure_closure,
// osl_mul_closure_{float,color} (src/liboslexec/opclosure.cpp:18-105) and the
// renderer's per-point bump allocator (src/testshade/render_state.h:27-54).
// A closure value is a 32-bit word offset into a per-thread pool (0 = NULL)
// instead of a pointer, so the tree is position independent and the pool can
// live in local memory or, interleaved by thread, in the CTA's shared memory:
//   component : [id][w.x][w.y][w.z][params ...]
//   mul       : [CL_MUL][w.x][w.y][w.z][child]
//   add       : [CL_ADD][a][b]
// Closure ids follow the renderer's registry (src/testrender/shading.h:25-60).
#pragma once

namespace osld {

enum ClosureIDs {
    CL_ADD = -2, CL_MUL = -1, COMPONENT_BASE_ID = 0,
    EMISSION_ID = 1, BACKGROUND_ID, DIFFUSE_ID, OREN_NAYAR_ID, TRANSLUCENT_ID, PHONG_ID, WARD_ID,
    MICROFACET_ID, REFLECTION_ID, FRESNEL_REFLECTION_ID, REFRACTION_ID, TRANSPARENT_ID, DEBUG_ID,
    HOLDOUT_ID, MX_OREN_NAYAR_DIFFUSE_ID, MX_BURLEY_DIFFUSE_ID, MX_DIELECTRIC_ID, MX_CONDUCTOR_ID,
    MX_GENERALIZED_SCHLICK_ID, MX_TRANSLUCENT_ID, MX_TRANSPARENT_ID, MX_SUBSURFACE_ID, MX_SHEEN_ID,
    MX_UNIFORM_EDF_ID, MX_ANISOTROPIC_VDF_ID, MX_MEDIUM_VDF_ID, MX_LAYER_ID, SPI_THINLAYER, EMPTY_ID
};

#ifndef OSLD_POOL_WORDS
#define OSLD_POOL_WORDS 256  // 1 KB, the reference's StackClosurePool size
#endif
// Storage behind a pool: the allocatable words plus a few words of slack, because the closure
// tree walkers read "the normal" (3 words after the header) of a component before they look at
// its id, and the last component of a tightly sized arena may be shorter than that.
#define OSLD_POOL_STORE (OSLD_POOL_WORDS + 4)

// Word i of a thread's arena sits at p[i * s]: s = 1 for a private (local-memory) array,
// s = CTA size when the arenas of a CTA are staged in shared memory, interleaved by thread
// so that a warp touching "its word i" hits 32 different banks.  The code generator sizes
// the arena from the closure ops of the material (OSLD_POOL_WORDS) so that it fits there.
struct PoolPtr {
    float* p;
    int s;
    OSLD float& operator[](int i) const { return p[i * s]; }
    OSLD PoolPtr operator+(int k) const
    {
        PoolPtr r;
        r.p = p + k * s;
        r.s = s;
        return r;
    }
};

struct ClosurePool {
    int used;   // next free word; word 0 is reserved so that offset 0 means NULL
    PoolPtr w;  // storage: bind() before use
    OSLD void bind(float* storage, int stride)
    {
        w.p = storage;
        w.s = stride;
    }
    OSLD void reset() { used = 1; }
    OSLD int alloc(int nwords)
    {
        if (used + nwords > OSLD_POOL_WORDS)
            return 0;
        int at = used;
        used += nwords;
        return at;
    }
    OSLD int id(int c) const { return __float_as_int(w[c]); }
    OSLD V3 weight(int c) const { return mkv(w[c + 1], w[c + 2], w[c + 3]); }
    OSLD int child(int c, int k) const { return __float_as_int(w[c + k]); }
};

OSLD bool v3_is_zero(V3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
OSLD bool v3_is_one(V3 a) { return a.x == 1.0f && a.y == 1.0f && a.z == 1.0f; }

// weighted == false: weight 1 (osl_allocate_closure_component)
OSLD int clos_component(ClosurePool& p, int id, int nparams, bool weighted, V3 wt)
{
    if (weighted && v3_is_zero(wt))
        return 0;
    int c = p.alloc(4 + nparams);
    if (c) {
        p.w[c] = __int_as_float(id);
        V3 ww  = weighted ? wt : mkv(1.0f);
        p.w[c + 1] = ww.x;
        p.w[c + 2] = ww.y;
        p.w[c + 3] = ww.z;
    }
    return c;
}
OSLD int clos_mul(ClosurePool& p, int a, V3 wt)
{
    if (!a || v3_is_zero(wt))
        return 0;
    if (v3_is_one(wt))
        return a;
    int c = p.alloc(5);
    if (c) {
        p.w[c]     = __int_as_float(CL_MUL);
        p.w[c + 1] = wt.x;
        p.w[c + 2] = wt.y;
        p.w[c + 3] = wt.z;
        p.w[c + 4] = __int_as_float(a);
    }
    return c;
}
OSLD int clos_mul(ClosurePool& p, int a, float wt)
{
    if (!a || wt == 0.0f)
        return 0;
    if (wt == 1.0f)
        return a;
    return clos_mul(p, a, mkv(wt));
}
OSLD int clos_add(ClosurePool& p, int a, int b)
{
    if (!a)
        return b;
    if (!b)
        return a;
    int c = p.alloc(3);
    if (c) {
        p.w[c]     = __int_as_float(CL_ADD);
        p.w[c + 1] = __int_as_float(a);
        p.w[c + 2] = __int_as_float(b);
    }
    return c;
}
OSLD void putp(PoolPtr q, float v) { q[0] = v; }
OSLD void putp(PoolPtr q, int v) { q[0] = __int_as_float(v); }
OSLD void putp(PoolPtr q, V3 v) { q[0] = v.x; q[1] = v.y; q[2] = v.z; }

}  // namespace osld
