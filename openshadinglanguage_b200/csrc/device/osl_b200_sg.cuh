// osl_b200_sg.cuh — the full per-point ShaderGlobals record used when groups
// are called from the wavefront renderer (product code).  Grid (testshade)
// kernels get a specialised SG holding only the fields the group reads; the
// renderer dispatches materials dynamically, so it fills every field of
// ShaderGlobals (src/include/OSL/shaderglobals.h:55-146) the way
// globals_from_hit does (src/testrender/simpleraytracer.cpp:889-932).
#pragma once

namespace osld {

struct SG {
    V3 P, P_dx, P_dy, dPdz;
    V3 I, I_dx, I_dy;
    V3 N, Ng;
    float u, u_dx, u_dy, v, v_dx, v_dy;
    V3 dPdu, dPdv;
    float time, dtime;
    V3 dPdtime;
    V3 Ps, Ps_dx, Ps_dy;
    float surfacearea;
    int raytype, flipHandedness, backfacing, shadeindex;
    int Ci;             // closure output: word offset into *pool (0 = none)
    ClosurePool* pool;  // renderer-owned per-point closure arena
    OSLD Dv P_d() const { return mkdv(P, P_dx, P_dy); }
    OSLD Dv I_d() const { return mkdv(I, I_dx, I_dy); }
    OSLD Dv Ps_d() const { return mkdv(Ps, Ps_dx, Ps_dy); }
    OSLD Df u_d() const { return mkd(u, u_dx, u_dy); }
    OSLD Df v_d() const { return mkd(v, v_dx, v_dy); }
};

}  // namespace osld
