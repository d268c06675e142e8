"""Shared test helpers: fixture loading, the standard testshade-style cases and
image quantisation.  The CPU oracle is imported here (tests may use it)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def oso(name):
    with open(os.path.join(GOLDEN, "oso", name + ".oso")) as f:
        return f.read()


def golden_image(name):
    d = np.load(os.path.join(GOLDEN, "images", name + ".npz"))
    return d["pixels"], int(d["step"]), tuple(int(x) for x in d["shape"])


def golden_text(name):
    with open(os.path.join(GOLDEN, "text", name + ".txt")) as f:
        return f.read()


def noise_vectors():
    with open(os.path.join(GOLDEN, "noise_vectors.json")) as f:
        return json.load(f)


def quantize_u8(img):
    """testshade -od uint8: OIIO float->uint8 is clamp(v,0,1)*255 + 0.5, truncated."""
    return (np.clip(img, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


# testshade-style single-output image cases that have a reference golden image.
# (testsuite/<name>/run.py: "-g R R -od uint8 -o Cout out.tif [-param ...] shader")
IMAGE_CASES = {
    "noise": dict(shader="noise_test", res=512, params={}),
    "cellnoise": dict(shader="cellnoise_test", res=512, params={}),
    "hashnoise": dict(shader="hashnoise_test", res=128, params={}),
    "noise-cell": dict(shader="testnoise", res=512,
                       params=dict(noisename="cell", offset=0.0, scale=1.0)),
    "noise-perlin": dict(shader="testnoise", res=512, params=dict(noisename="perlin")),
    "noise-simplex": dict(shader="testnoise", res=512, params=dict(noisename="simplex")),
    "noise-gabor": dict(shader="testnoise", res=512, params=dict(noisename="gabor")),
    "noise-gabor2d-filter": dict(shader="gabor2d_filter_test", res=512, params={}),
    "noise-gabor3d-filter": dict(shader="gabor3d_filter_test", res=512, params={}),
    "pnoise-gabor": dict(shader="testpnoise", res=512, params=dict(noisename="gabor")),
    "pnoise": dict(shader="pnoise_test", res=512, params={}),
    "pnoise-cell": dict(shader="testpnoise", res=512,
                        params=dict(noisename="cell", offset=0.0, scale=1.0)),
    "pnoise-perlin": dict(shader="testpnoise", res=512, params=dict(noisename="perlin")),
    # noise("name", ...) with a name picked per point from a string array: run-time dispatch
    # (GenericNoise / GenericPNoise, opnoise.cpp:704-900)
    "noise-generic": dict(shader="noise_generic_test", res=512, params={}),
    "pnoise-generic": dict(shader="pnoise_generic_test", res=512, params={}),
}


def image_case_group(case):
    c = IMAGE_CASES[case]
    layers = [dict(oso=oso(c["shader"]), name="layer0", params=dict(c["params"]))]
    outputs = [dict(name="Cout", offset=0, stride=12)]
    return layers, outputs, c["res"]


# testsuite/spline: "-g 256 256 -od uint8 -o Cspline color.tif -o DxCspline dcolor.tif
#   -o Fspline float.tif -o DxFspline dfloat.tif -o NumKnots numknots.tif test"
# one dense arena per output like testshade's setup_output_images
SPLINE_OUTPUTS = [("Cspline", 3, "spline-color"), ("DxCspline", 3, "spline-dcolor"),
                  ("Fspline", 1, "spline-float"), ("DxFspline", 1, "spline-dfloat"),
                  ("NumKnots", 1, "spline-numknots")]


def spline_case(res=256):
    layers = [dict(oso=oso("spline_test"), name="layer0", params={})]
    outputs, off = [], 0
    for name, ch, _ in SPLINE_OUTPUTS:
        outputs.append(dict(name=name, offset=off, stride=4 * ch))
        off += 4 * ch * res * res
    return layers, outputs, off // 4


# tests/shaders/color_ops.osl: one dense arena per output
COLOR_OPS_OUTPUTS = [("Cto", 3), ("Cback", 3), ("Cctor", 3), ("Csrgb", 3), ("Clin", 3), ("DxCto", 3),
                     ("DyCto", 3), ("Lum", 1), ("DxLum", 1), ("BB", 3), ("WL", 3)]


# Reference testsuite directories whose run.py is one `testshade [-g X Y] [-center] test` with a
# text golden (tools/make_fixtures.py TESTSUITE_TEXT): dir -> (grid x, grid y, center).
# Between them they print the value of nearly every math / logic / control-flow op, so they
# pin the restated OIIO fast_* transcendentals at print precision (trig, exponential, hyperb,
# miscmath, geomath, blendmath ...).
TESTSUITE_TEXT = {
    "arithmetic": (1, 1, 0), "array-derivs": (2, 2, 0), "blendmath": (1, 1, 0), "breakcont": (1, 1, 0),
    "bug-locallifetime": (1, 1, 0), "bug-peep": (2, 2, 0), "comparison": (1, 1, 0),
    "const-array-fill": (1, 1, 0), "const-array-params": (1, 1, 0), "derivs": (2, 2, 0),
    "derivs-muldiv-clobber": (1, 1, 0), "exit": (1, 1, 0), "exponential": (1, 1, 0),
    "function-earlyreturn": (2, 2, 0), "function-outputelem": (2, 2, 0), "function-simple": (1, 1, 0),
    "geomath": (2, 2, 0), "hex": (1, 1, 0), "hyperb": (1, 1, 0), "ieee_fp": (1, 1, 0), "incdec": (1, 1, 0),
    "intbits": (1, 1, 0), "logic": (1, 1, 0), "loop": (2, 2, 0), "miscmath": (2, 2, 0),
    "named-components": (1, 1, 0), "oslc-literalfold": (1, 1, 0), "pragma-nowarn": (1, 1, 0),
    "printf-whole-array": (1, 1, 0), "select": (2, 2, 0), "shortcircuit": (2, 2, 0),
    "spline-boundarybug": (1, 1, 0), "splineinverse": (3, 1, 1), "ternary": (2, 2, 0),
    "transitive-assign": (1, 1, 0), "trig": (1, 1, 0), "typecast": (2, 2, 0), "userdata-defaults": (1, 1, 0),
    "userdata": (2, 2, 0),
    "vecctr": (1, 1, 0), "vector": (1, 1, 0),
}


def testsuite_text_want(d):
    return "\n".join(l for l in golden_text("ts_" + d).split("\n") if not l.startswith("Compiled"))


# tests/shaders/texture_ops.osl
TEXTURE_OPS_OUTPUTS = [("Cdef", 3), ("Cperiodic", 3), ("Cmirror", 3), ("Cclamp", 3), ("Cbilinear", 3), ("Cclosest", 3),
                       ("Cblur", 3), ("Cwide", 3), ("Cexplicit", 3), ("Fone", 1), ("Cnoderiv", 3)]
TEXTURES = os.path.join(GOLDEN, "textures")


# tests/shaders/matrix_ops.osl
MATRIX_OPS_OUTPUTS = [("Pshader", 3), ("Vobj", 3), ("Nmy", 3), ("Pback", 3), ("Pm", 3), ("DxPm", 3), ("Det", 1),
                      ("Row", 3), ("Ok", 1), ("Punk", 3), ("Eq", 1), ("Nm", 3), ("PjP", 3), ("PjN", 3), ("DyV", 3)]


def multi_output_case(shader, outputs_spec, res, params=None):
    layers = [dict(oso=oso(shader), name="layer0", params=dict(params or {}))]
    outputs, off = [], 0
    for name, ch in outputs_spec:
        outputs.append(dict(name=name, offset=off, stride=4 * ch))
        off += 4 * ch * res * res
    return layers, outputs, off // 4


def multi_output_split(arena, outputs_spec, res):
    out, o = {}, 0
    for name, ch in outputs_spec:
        out[name] = arena[o:o + ch * res * res].reshape(res * res, ch)
        o += ch * res * res
    return out


def color_ops_case(space, res=96):
    layers = [dict(oso=oso("color_ops"), name="layer0", params=dict(space=space))]
    outputs, off = [], 0
    for name, ch in COLOR_OPS_OUTPUTS:
        outputs.append(dict(name=name, offset=off, stride=4 * ch))
        off += 4 * ch * res * res
    return layers, outputs, off // 4


def color_ops_split(arena, res=96):
    out, o = {}, 0
    for name, ch in COLOR_OPS_OUTPUTS:
        out[name] = arena[o:o + ch * res * res].reshape(res * res, ch)
        o += ch * res * res
    return out


def check_spline_images(arena, res=256):
    o = 0
    for name, ch, gold in SPLINE_OUTPUTS:
        img = arena[o:o + ch * res * res].reshape(res, res, ch)
        o += ch * res * res
        ref, step, shape = golden_image(gold)
        assert shape == (res, res)
        d = np.abs(quantize_u8(img)[::step, ::step].astype(int) - ref[..., :ch].astype(int))
        assert (d > 1).mean() <= 0.0005, (name, (d > 1).mean())


LAYERS_LAZY = dict(
    layers=[("layers_lazy_a", "alayer"), ("layers_lazy_b", "blayer"), ("layers_lazy_c", "clayer")],
    connections=[("alayer", "f_out", "clayer", "f_in"), ("alayer", "c_out", "clayer", "c_in"),
                 ("blayer", "out", "clayer", "unused")])


def layers_group(with_outputs=True, derivs=True):
    """BASELINE config 2: the layers-lazy 3-layer group with alayer.f_out and
    alayer.c_out as renderer outputs placed in one interleaved record
    (val,dx,dy => 12 + 36 = 48 B/pt)."""
    layers = [dict(oso=oso(s), name=n, params={}) for s, n in LAYERS_LAZY["layers"]]
    outputs = []
    if with_outputs:
        if derivs:
            outputs = [dict(name="alayer.f_out", offset=0, stride=48, derivs=True),
                       dict(name="alayer.c_out", offset=12, stride=48, derivs=True)]
        else:
            outputs = [dict(name="alayer.f_out", offset=0, stride=16, derivs=False),
                       dict(name="alayer.c_out", offset=4, stride=16, derivs=False)]
    return layers, list(LAYERS_LAZY["connections"]), outputs


def testshade_userdata(n, var, uni, extra=()):
    """What testshade's SimpleRenderer::get_userdata supplies (src/testshade/simplerend.cpp:517-590):
    s = u and t = v with their derivatives, face_idx = int(4u), and three partially available
    values used by testsuite/userdata-partial: red = u where P.x > 0.5, green = v where
    P.x < 0.5, blue = 1-u where int(P.y*12) is even.  `extra`: --userdata NAME VALUE entries
    (name, numpy value), uniform over the grid.  -> entries for pack_userdata()."""
    def field(name, comps=1):
        if name in var:
            return np.asarray(var[name], np.float32).reshape(comps, n).T.copy()
        vals = list(uni.get(name, [0.0] * comps)) + [0.0] * comps
        return np.tile(np.asarray(vals[:comps], np.float32), (n, 1))
    u, dudx, dudy = field("u"), field("dudx"), field("dudy")
    v, dvdx, dvdy = field("v"), field("dvdx"), field("dvdy")
    P = field("P", 3)
    ents = [dict(name="s", data=np.hstack([u, dudx, dudy]), derivs=True),
            dict(name="t", data=np.hstack([v, dvdx, dvdy]), derivs=True),
            dict(name="face_idx", data=(np.float32(4) * u[:, 0]).astype(np.int32)),
            dict(name="red", data=np.hstack([u, dudx, dudy]), derivs=True, valid=(P[:, 0] > np.float32(0.5)).astype(np.int32)),
            dict(name="green", data=np.hstack([v, dvdx, dvdy]), derivs=True, valid=(P[:, 0] < np.float32(0.5)).astype(np.int32)),
            dict(name="blue", data=np.hstack([np.float32(1) - u, -dudx, -dudy]), derivs=True,
                 valid=(((P[:, 1] * np.float32(12)).astype(np.int32) % 2) == 0).astype(np.int32))]
    for name, val in extra:
        a = np.asarray(val)
        ents.append(dict(name=name, data=np.tile(a.reshape(1, -1), (n, 1))))
    return ents


def _displace_planes(g):
    """displace_geometry()'s per-point globals ([n, 3] rows / [n]) -> SoA planes (x[n] y[n] z[n])."""
    return {k: np.ascontiguousarray(v.T).reshape(-1) if v.ndim == 2 else np.ascontiguousarray(v) for k, v in g.items()}


def oracle_displace(layers, conns, g, n):
    """Scene.prepare(displace=...): the displacement group over the n vertex points on the CPU oracle."""
    from oracle import oracle
    ls = [dict(oso=oso(l["shader"]), name=l["name"], params=l["params"]) for l in layers]
    og = oracle.OracleGroup(ls, conns, [dict(name="P", offset=0, stride=12, derivs=False)])
    out = np.zeros((n, 3), np.float32)
    og.run(n, _displace_planes(g), {}, out, nthreads=os.cpu_count() or 8)
    return out


def device_displace(dev, options="fma=0"):
    """The same through the product (api.device_displacer): one b200_group_execute over the vertex batch,
    P read back as a renderer output."""
    from openshadinglanguage_b200 import api
    return api.device_displacer(oso, dev, options)
