"""Host-side mirror of the testshade grid harness (the caller of the hot path).

`grid_globals` produces, for a whole W x H grid at once and in the SoA layout
of include/osl_b200.h, the ShaderGlobals that the reference's testshade fills
one point at a time in setup_shaderglobals()
(src/testshade/testshade.cpp:957-1046): u,v across the patch, P=(u,v,1),
constant or varying derivatives, N=Ng=(0,0,1), dPdu/dPdv, surfacearea=1 and the
camera raytype bit.  Point index = y*xres + x (testshade.cpp:1138-1158).
"""
import numpy as np


def harness_transforms():
    """testshade's setup_transformations() (src/testshade/testshade.cpp:925-950) as row-major
    float32 4x4 lists: "shader" = translate(1,0,0) . rotate z 45 deg, "object" =
    translate(0,1,0) . rotate z 90 deg, and the renderer-named "myspace" = scale(1,2,1)."""
    f = np.float32

    def make(t, angle):
        c, s = f(np.cos(f(angle))), f(np.sin(f(angle)))
        return [c, s, f(0), f(0), -s, c, f(0), f(0), f(0), f(0), f(1), f(0), f(t[0]), f(t[1]), f(t[2]), f(1)]
    my = [1, 0, 0, 0, 0, 2, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    return {"shader": [float(x) for x in make((1, 0, 0), np.pi / 4)],
            "object": [float(x) for x in make((0, 1, 0), np.pi / 2)],
            "myspace": [float(x) for x in my]}


def grid_globals(xres, yres, center=False, vary_udxdy=False, vary_vdxdy=False, vary_pdxdy=False,
                 uscale=1.0, vscale=1.0, uoffset=0.0, voffset=0.0, raytype_bit=1):
    """-> (varying: {field: float32 array [comps*n]}, uniform: {field: [values]})"""
    f = np.float32
    n = xres * yres
    ix = (np.arange(n, dtype=np.int64) % xres).astype(f)
    iy = (np.arange(n, dtype=np.int64) // xres).astype(f)
    if center:
        uu = (ix + f(0.5)) / f(xres)
        vv = (iy + f(0.5)) / f(yres)
        du, dv = f(uscale) / f(xres), f(vscale) / f(yres)
    else:
        uu = np.full(n, 0.5, f) if xres == 1 else ix / f(xres - 1)
        vv = np.full(n, 0.5, f) if yres == 1 else iy / f(yres - 1)
        du, dv = f(uscale) / f(max(1, xres - 1)), f(vscale) / f(max(1, yres - 1))
    u = (f(uscale) * uu + f(uoffset)).astype(f)
    v = (f(vscale) * vv + f(voffset)).astype(f)
    varying = {"u": u, "v": v, "P": np.concatenate([u, v, np.ones(n, f)])}
    uniform = {"N": [0, 0, 1], "Ng": [0, 0, 1], "dPdu": [1, 0, 0], "dPdv": [0, 1, 0],
               "surfacearea": [1.0], "raytype": [raytype_bit],
               # named coordinate systems, not a ShaderGlobals field: api.py passes them
               # as b200_globals.transforms
               "transforms": harness_transforms()}
    one = f(1.0)
    if vary_udxdy:
        varying["dudx"], varying["dudy"] = (one - u).astype(f), u.copy()
    else:
        uniform["dudx"] = [float(du)]
    if vary_vdxdy:
        varying["dvdx"], varying["dvdy"] = (one - v).astype(f), v.copy()
    else:
        uniform["dvdy"] = [float(dv)]
    if vary_pdxdy:
        half = f(0.5)
        varying["dPdx"] = np.concatenate([one - u, one - v, u * half]).astype(f)
        varying["dPdy"] = np.concatenate([one - v, one - u, v * half]).astype(f)
    else:
        uniform["dPdx"] = [float(f(uscale) / f(max(1, xres - 1))), 0.0, 0.0]
        uniform["dPdy"] = [0.0, float(f(vscale) / f(max(1, yres - 1))), 0.0]
    return varying, uniform
