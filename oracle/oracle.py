"""oracle.py — CPU ORACLE driver (test infrastructure, NOT product code).

Python front end for the scalar C++ oracle: builds a group with oso2cpp,
loads the shared object with ctypes and runs it over SoA globals.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.
"""
import ctypes
import os

import numpy as np

from . import oso2cpp

SG_FIELDS = ["P", "dPdx", "dPdy", "dPdz", "I", "dIdx", "dIdy", "N", "Ng", "u", "dudx",
             "dudy", "v", "dvdx", "dvdy", "dPdu", "dPdv", "time", "dtime", "dPdtime",
             "Ps", "dPsdx", "dPsdy", "surfacearea", "raytype", "flipHandedness",
             "backfacing"]
SG_VEC = {"P", "dPdx", "dPdy", "dPdz", "I", "dIdx", "dIdy", "N", "Ng", "dPdu", "dPdv",
          "dPdtime", "Ps", "dPsdx", "dPsdy"}
SG_INT = {"raytype", "flipHandedness", "backfacing"}
NF = len(SG_FIELDS)


class Launch(ctypes.Structure):
    _fields_ = [("varying", ctypes.c_void_p * NF),
                ("uniform", (ctypes.c_float * 4) * NF),
                ("plane_stride", ctypes.c_longlong),
                ("shadeindex", ctypes.c_void_p),
                ("output_base", ctypes.c_void_p),
                ("userdata_base", ctypes.c_void_p),
                ("ntransforms", ctypes.c_int),
                ("transforms", ctypes.c_void_p),
                ("nuserdata", ctypes.c_int),
                ("userdata", ctypes.c_void_p)]


class UserDataDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("ncomp", ctypes.c_int), ("is_int", ctypes.c_int),
                ("offset", ctypes.c_longlong), ("stride", ctypes.c_longlong), ("derivs", ctypes.c_int),
                ("valid_offset", ctypes.c_longlong), ("valid_stride", ctypes.c_longlong)]


def pack_userdata(entries):
    """entries: list of dict(name, data [n, ncomp] or [n, 3*ncomp] with derivs (val,dx,dy), or int32 [n];
    derivs=bool, valid=int32[n] or None) -> (arena uint8 array, list of descriptor dicts).  One dense
    array per entry (offset = its start, stride = its row size): the layout testshade uses for
    its outputs, here for the UserData arena."""
    blobs, descs, off = [], [], 0
    for e in entries:
        a = np.ascontiguousarray(e["data"])
        is_int = a.dtype.kind in "iu"
        a = a.astype(np.int32 if is_int else np.float32).reshape(len(a), -1)
        derivs = bool(e.get("derivs"))
        ncomp = a.shape[1] // (3 if derivs else 1)
        d = dict(name=e["name"], ncomp=ncomp, is_int=int(is_int), offset=off, stride=a.shape[1] * 4,
                 derivs=int(derivs), valid_offset=-1, valid_stride=0)
        blobs.append(a.tobytes())
        off += (len(blobs[-1]) + 15) // 16 * 16
        blobs[-1] = blobs[-1].ljust((len(blobs[-1]) + 15) // 16 * 16, b"\0")
        if e.get("valid") is not None:
            v = np.ascontiguousarray(e["valid"], np.int32)
            d["valid_offset"], d["valid_stride"] = off, 4
            blobs.append(v.tobytes().ljust((v.nbytes + 15) // 16 * 16, b"\0"))
            off += len(blobs[-1])
        descs.append(d)
    return np.frombuffer(b"".join(blobs) or b"\0" * 16, np.uint8).copy(), descs


class NamedTransform(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("m", ctypes.c_float * 16)]


def testshade_transforms():
    """setup_transformations (src/testshade/testshade.cpp:925-950): "shader" = translate(1,0,0)
    then rotate z 45 deg, "object" = translate(0,1,0) then rotate z 90 deg, "myspace" =
    scale(1,2,1).  Imath's Matrix44::translate / rotate / scale in float32; these are harness
    inputs handed identically to the oracle and to the device."""
    f32 = np.float32

    def translate(M, t):
        M = M.copy()
        for j in range(4):
            M[3, j] = f32(M[3, j] + f32(f32(f32(t[0] * M[0, j]) + f32(t[1] * M[1, j])) + f32(t[2] * M[2, j])))
        return M

    def rotate_z(M, rz):
        c, s = f32(np.cos(f32(rz))), f32(np.sin(f32(rz)))
        m = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], f32)   # Matrix44::rotate with rx = ry = 0
        P = M.copy()
        for i in range(3):
            for j in range(4):
                M[i, j] = f32(f32(f32(P[0, j] * m[i, 0]) + f32(P[1, j] * m[i, 1])) + f32(P[2, j] * m[i, 2]))
        return M
    I = np.eye(4, dtype=f32)
    Mshad = rotate_z(translate(I, (f32(1), f32(0), f32(0))), np.pi / 4)
    Mobj = rotate_z(translate(I, (f32(0), f32(1), f32(0))), np.pi / 2)
    Mmy = I.copy()
    Mmy[1, :] *= f32(2)
    return {"shader": Mshad.reshape(-1).tolist(), "object": Mobj.reshape(-1).tolist(),
            "myspace": Mmy.reshape(-1).tolist()}


def testshade_globals(xres, yres, center=False, vary_udxdy=False, vary_vdxdy=False,
                      vary_pdxdy=False, uscale=1.0, vscale=1.0, uoffset=0.0, voffset=0.0,
                      raytype=1):
    """Restates testshade's setup_shaderglobals (src/testshade/testshade.cpp:957-1046)
    for the whole grid at once; returns (varying dict of float32 arrays,
    uniform dict).  shadeindex = y*xres + x."""
    f32 = np.float32
    x = np.tile(np.arange(xres, dtype=f32), yres)
    y = np.repeat(np.arange(yres, dtype=f32), xres)
    us, vs, uo, vo = f32(uscale), f32(vscale), f32(uoffset), f32(voffset)
    var, uni = {}, {}
    if center:
        u = us * ((x + f32(0.5)) / f32(xres)) + uo
        v = vs * ((y + f32(0.5)) / f32(yres)) + vo
        dudx, dvdy = us / f32(xres), vs / f32(yres)
    else:
        u = us * (np.full_like(x, 0.5) if xres == 1 else x / f32(xres - 1)) + uo
        v = vs * (np.full_like(y, 0.5) if yres == 1 else y / f32(yres - 1)) + vo
        dudx, dvdy = us / f32(max(1, xres - 1)), vs / f32(max(1, yres - 1))
    u = u.astype(f32)
    v = v.astype(f32)
    var["u"], var["v"] = u, v
    if vary_udxdy:
        var["dudx"], var["dudy"] = (f32(1) - u).astype(f32), u.copy()
    else:
        uni["dudx"] = [float(dudx)]
    if vary_vdxdy:
        var["dvdx"], var["dvdy"] = (f32(1) - v).astype(f32), v.copy()
    else:
        uni["dvdy"] = [float(dvdy)]
    var["P"] = np.stack([u, v, np.ones_like(u)]).astype(f32)
    if vary_pdxdy:
        var["dPdx"] = np.stack([f32(1) - u, f32(1) - v, (u.astype(np.float64) * 0.5).astype(f32)]).astype(f32)
        var["dPdy"] = np.stack([f32(1) - v, f32(1) - u, (v.astype(np.float64) * 0.5).astype(f32)]).astype(f32)
    else:
        uni["dPdx"] = [float(us / f32(max(1, xres - 1))), 0.0, 0.0]
        uni["dPdy"] = [0.0, float(vs / f32(max(1, yres - 1))), 0.0]
    uni["dPdu"] = [1.0, 0.0, 0.0]
    uni["dPdv"] = [0.0, 1.0, 0.0]
    uni["N"] = [0.0, 0.0, 1.0]
    uni["Ng"] = [0.0, 0.0, 1.0]
    uni["surfacearea"] = [1.0]
    uni["raytype"] = [raytype]
    uni["transforms"] = testshade_transforms()
    return var, uni


def make_launch(n, varying, uniform, output, shadeindex=None, keep=None):
    """Fill a Launch from numpy arrays; `keep` collects references."""
    L = Launch()
    keep = keep if keep is not None else []
    for i, f in enumerate(SG_FIELDS):
        L.varying[i] = None
        if f in varying:
            a = np.ascontiguousarray(varying[f], dtype=np.int32 if f in SG_INT else np.float32)
            assert a.size == n * (3 if f in SG_VEC else 1), (f, a.shape, n)
            keep.append(a)
            L.varying[i] = a.ctypes.data
        vals = list(uniform.get(f, []))
        for c in range(4):
            if f in SG_INT:
                iv = int(vals[0]) if vals else 0
                L.uniform[i][c] = np.array([iv], dtype=np.int32).view(np.float32)[0] if c == 0 else 0.0
            else:
                L.uniform[i][c] = float(vals[c]) if c < len(vals) else 0.0
    L.plane_stride = n
    if shadeindex is not None:
        si = np.ascontiguousarray(shadeindex, dtype=np.int32)
        keep.append(si)
        L.shadeindex = si.ctypes.data
    else:
        L.shadeindex = None
    L.output_base = output.ctypes.data if output is not None else None
    L.userdata_base = None
    L.nuserdata, L.userdata = 0, None
    ud = uniform.get("userdata")
    if ud:
        arena, descs = ud if isinstance(ud, tuple) else pack_userdata(ud)
        arr = (UserDataDesc * len(descs))()
        for k, d in enumerate(descs):
            nm = d["name"].encode()
            keep.append(nm)
            arr[k].name, arr[k].ncomp, arr[k].is_int = nm, d["ncomp"], d["is_int"]
            arr[k].offset, arr[k].stride, arr[k].derivs = d["offset"], d["stride"], d["derivs"]
            arr[k].valid_offset, arr[k].valid_stride = d["valid_offset"], d["valid_stride"]
        keep += [arena, arr]
        L.userdata_base, L.nuserdata, L.userdata = arena.ctypes.data, len(descs), ctypes.cast(arr, ctypes.c_void_p)
    keep.append(output)
    xf = uniform.get("transforms") or {}
    arr = (NamedTransform * max(1, len(xf)))()
    for k, (name, m) in enumerate(xf.items()):
        arr[k].name = name.encode()
        for c in range(16):
            arr[k].m[c] = float(m[c])
    keep.append(arr)
    L.ntransforms = len(xf)
    L.transforms = ctypes.cast(arr, ctypes.c_void_p)
    return L, keep


_shadeops = None


def build_shadeops():
    """Compile oracle_shadeops.cpp -> oracle/_build/liboracle_shadeops.so"""
    import subprocess
    here = oso2cpp.HERE
    bdir = os.path.join(here, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "liboracle_shadeops.so")
    srcs = [os.path.join(here, f) for f in ("oracle_shadeops.cpp", "osl_oracle.h", "osl_oracle_ops.h", "osl_oracle_simplex.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                            "-I", here, srcs[0], "-o", so + ".tmp"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle shadeops compile failed:\n" + r.stderr[:4000])
        os.replace(so + ".tmp", so)
    return so


def build_bsdl_check():
    """Compile oracle_bsdl_check.cpp -> oracle/_build/liboracle_bsdl_check.so (the restated libbsdl
    lobes behind the signature of oracle/ref_bsdl.cpp)."""
    import subprocess
    here = oso2cpp.HERE
    bdir = os.path.join(here, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, "liboracle_bsdl_check.so")
    srcs = [os.path.join(here, "oracle_bsdl_check.cpp")] + [os.path.join(here, f) for f in sorted(os.listdir(here))
                                                            if f.endswith(".h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                            "-I", here, srcs[0], "-o", so + ".tmp"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle bsdl check compile failed:\n" + r.stderr[:4000])
        os.replace(so + ".tmp", so)
    return so


BSDL_LUTS = os.path.join(os.path.dirname(oso2cpp.HERE), "openshadinglanguage_b200", "data", "bsdl_luts.bin")


def bsdl_luts():
    """The energy tables of the libbsdl microfacet lobes (float32, layout in osl_oracle_mxlobes.h):
    data baked by tools/bake_bsdl_luts.cpp and shipped with the product."""
    luts = np.fromfile(BSDL_LUTS, np.float32)
    ltc = os.path.join(os.path.dirname(BSDL_LUTS), "zeltner_ltc.bin")   # tools/bake_zeltner_ltc.py
    thin = os.path.join(os.path.dirname(BSDL_LUTS), "thinlayer_lut.bin")   # spi::Thinlayer, same baker
    return np.concatenate([luts, np.fromfile(ltc, np.float32), np.fromfile(thin, np.float32)])


def shadeops():
    global _shadeops
    if _shadeops is None:
        L = ctypes.CDLL(build_shadeops())
        L.oracle_noise.argtypes = [ctypes.c_int] * 4 + [ctypes.c_longlong, ctypes.c_void_p,
                                                       ctypes.c_void_p, ctypes.c_void_p]
        L.oracle_hash.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p]
        _shadeops = L
    return _shadeops


KINDS = {"noise": 0, "snoise": 1, "cellnoise": 2, "hashnoise": 3, "simplex": 4, "usimplex": 5}


def noise(kind, outdim, inp, period=None, derivs=False):
    """inp: float32 [indim*(3 if derivs else 1), n] planes -> [outdim*(3 if derivs else 1), n]"""
    inp = np.ascontiguousarray(inp, np.float32)
    rows, n = inp.shape
    indim = rows // 3 if derivs else rows
    out = np.zeros((outdim * (3 if derivs else 1), n), np.float32)
    per = None if period is None else np.ascontiguousarray(period, np.float32)
    rc = shadeops().oracle_noise(KINDS[kind], outdim, indim, 1 if derivs else 0, n, inp.ctypes.data,
                                 None if per is None else per.ctypes.data, out.ctypes.data)
    assert rc == 0
    return out


def hash_(inp):
    inp = np.ascontiguousarray(inp, np.float32)
    indim, n = inp.shape
    out = np.zeros(n, np.int32)
    shadeops().oracle_hash(indim, n, inp.ctypes.data, out.ctypes.data)
    return out


class OracleGroup:
    """layers: list of dict(oso=<text>, name=<layername>, params={...})"""

    def __init__(self, layers, connections=(), outputs=(), opt="-O2", flags=(), textures=None, name="group",
                 attributes=None):
        ls = [oso2cpp.Layer(l["oso"], l["name"], l.get("params")) for l in layers]
        self.group = oso2cpp.Group(ls, connections, outputs, name=name, attributes=attributes)
        self.so = oso2cpp.build_group(self.group, opt=opt, extra_flags=flags)
        self.lib = ctypes.CDLL(self.so)
        self.lib.oracle_run_mt.argtypes = [ctypes.POINTER(Launch), ctypes.c_longlong, ctypes.c_int]
        self.lib.oracle_run_capture.argtypes = [ctypes.POINTER(Launch), ctypes.c_longlong,
                                                ctypes.c_longlong]
        self.lib.oracle_run_capture.restype = ctypes.c_char_p
        register_textures(self.lib, textures)

    def run(self, n, varying, uniform, output, shadeindex=None, nthreads=1):
        L, keep = make_launch(n, varying, uniform, output, shadeindex)
        self.lib.oracle_run_mt(ctypes.byref(L), n, nthreads)
        return output

    def run_capture(self, n, varying, uniform, output=None, shadeindex=None, error_repeats=False):
        L, keep = make_launch(n, varying, uniform, output, shadeindex)
        self.lib.oracle_set_error_repeats(1 if error_repeats else 0)
        return self.lib.oracle_run_capture(ctypes.byref(L), 0, n).decode()


# ---------------------------------------------------------------------------
# testrender oracle
# ---------------------------------------------------------------------------
class RenderScene(ctypes.Structure):
    _fields_ = [("nverts", ctypes.c_int), ("ntris", ctypes.c_int), ("nnodes", ctypes.c_int),
                ("nlightprims", ctypes.c_int), ("nshaders", ctypes.c_int), ("nmeshes", ctypes.c_int),
                ("verts", ctypes.c_void_p), ("normals", ctypes.c_void_p), ("uvs", ctypes.c_void_p),
                ("triangles", ctypes.c_void_p), ("n_triangles", ctypes.c_void_p),
                ("uv_triangles", ctypes.c_void_p), ("shaderids", ctypes.c_void_p),
                ("meshids", ctypes.c_void_p), ("mesh_surfacearea", ctypes.c_void_p),
                ("bvh_nodes", ctypes.c_void_p), ("bvh_indices", ctypes.c_void_p),
                ("lightprims", ctypes.c_void_p), ("shader_is_light", ctypes.c_void_p),
                ("eye", ctypes.c_float * 3), ("dir", ctypes.c_float * 3), ("up", ctypes.c_float * 3),
                ("fov", ctypes.c_float), ("cx", ctypes.c_float * 3), ("cy", ctypes.c_float * 3),
                ("invw", ctypes.c_float), ("invh", ctypes.c_float),
                ("xres", ctypes.c_int), ("yres", ctypes.c_int),
                ("aa", ctypes.c_int), ("max_bounces", ctypes.c_int), ("rr_depth", ctypes.c_int),
                ("no_jitter", ctypes.c_int), ("show_globals", ctypes.c_int),
                ("background_shader", ctypes.c_int), ("background_resolution", ctypes.c_int)]


def fill_render_scene(RS, scene, arrays, xres, yres, aa, max_bounces=None, rr_depth=None,
                      no_jitter=False, show_globals=0):
    """Fill a RenderScene-shaped ctypes struct (oracle and product share the layout)."""
    if max_bounces is None:
        max_bounces = scene.options.get("max_bounces", 1000000)
    if rr_depth is None:
        rr_depth = scene.options.get("rr_depth", 5)
    keep = []
    rs = RS()

    def ptr(name, dtype):
        a = np.ascontiguousarray(arrays[name], dtype)
        keep.append(a)
        return a.ctypes.data
    rs.nverts, rs.ntris = len(arrays["verts"]), len(arrays["triangles"])
    rs.nnodes, rs.nlightprims = len(arrays["bvh_nodes"]), len(arrays["lightprims"])
    rs.nshaders, rs.nmeshes = len(scene.materials), len(arrays["mesh_surfacearea"])
    rs.verts, rs.normals, rs.uvs = ptr("verts", np.float32), ptr("normals", np.float32), ptr("uvs", np.float32)
    rs.triangles, rs.n_triangles = ptr("triangles", np.int32), ptr("n_triangles", np.int32)
    rs.uv_triangles, rs.shaderids = ptr("uv_triangles", np.int32), ptr("shaderids", np.int32)
    rs.meshids, rs.mesh_surfacearea = ptr("meshids", np.int32), ptr("mesh_surfacearea", np.float32)
    rs.bvh_nodes, rs.bvh_indices = ptr("bvh_nodes", np.float32), ptr("bvh_indices", np.uint32)
    rs.lightprims = ptr("lightprims", np.uint32) if len(arrays["lightprims"]) else None
    rs.shader_is_light = ptr("shader_is_light", np.int32)
    for i in range(3):
        rs.eye[i], rs.dir[i], rs.up[i] = float(scene.eye[i]), float(scene.dir[i]), float(scene.up[i])
    rs.fov = float(scene.fov)
    rs.xres, rs.yres, rs.aa = xres, yres, aa
    rs.max_bounces, rs.rr_depth = max_bounces, rr_depth
    rs.no_jitter, rs.show_globals = int(no_jitter), show_globals
    rs.background_shader, rs.background_resolution = scene.background_shader, scene.background_resolution
    return rs, keep


def material_groups(scene, oso_lookup):
    """[(layers, connections)] with .oso text resolved, ready for a generator."""
    mats = []
    for layers, conns in scene.materials:
        mats.append(([dict(oso=oso_lookup(l["shader"]), name=l["name"], params=l["params"]) for l in layers],
                     conns))
    return mats


def load_hdr(path):
    """Radiance RGBE (.hdr) -> float32 [h, w, 3], top scanline first.  Texel value =
    mantissa * 2^(e-136), zero when e == 0 (Ward's rgbe2float, which OIIO's hdr reader
    follows); handles the new-style per-channel run-length coding and flat pixels."""
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while True:
        end = data.index(b"\n", pos)
        line = data[pos:end]
        pos = end + 1
        if line == b"":
            break
    end = data.index(b"\n", pos)
    res = data[pos:end].split()
    pos = end + 1
    if res[0] != b"-Y" or res[2] != b"+X":
        raise ValueError("unsupported .hdr orientation %r" % res)
    h, w = int(res[1]), int(res[3])
    buf = np.frombuffer(data, np.uint8)
    rgbe = np.zeros((h, w, 4), np.uint8)
    for y in range(h):
        if 8 <= w < 32768 and buf[pos] == 2 and buf[pos + 1] == 2 and (int(buf[pos + 2]) << 8 | int(buf[pos + 3])) == w:
            pos += 4
            for c in range(4):
                x = 0
                while x < w:
                    n = int(buf[pos])
                    if n > 128:
                        n -= 128
                        rgbe[y, x:x + n, c] = buf[pos + 1]
                        pos += 2
                    else:
                        rgbe[y, x:x + n, c] = buf[pos + 1:pos + 1 + n]
                        pos += 1 + n
                    x += n
        else:
            rgbe[y] = buf[pos:pos + 4 * w].reshape(w, 4)
            pos += 4 * w
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return (rgbe[..., :3].astype(np.float32) * scale[..., None]).astype(np.float32)


def scene_textures(scene):
    """{name as written in the shader parameter: float32 [h, w, nch]} for every string
    parameter of the scene's groups that names an image file next to the scene."""
    out = {}
    for layers, _ in scene.materials:
        for l in layers:
            for v in (l.get("params") or {}).values():
                if v and isinstance(v[0], str) and v[0].lower().endswith(".hdr"):
                    p = os.path.join(getattr(scene, "basedir", "."), v[0])
                    if os.path.exists(p):
                        out[v[0]] = load_hdr(p)
    return out


def register_textures(lib, textures):
    lib.oracle_texture_add.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    for name, img in (textures or {}).items():
        img = np.ascontiguousarray(img, np.float32)
        if img.ndim == 2:
            img = img[..., None]
        lib.oracle_texture_add(name.encode(), img.shape[1], img.shape[0], img.shape[2], img.ctypes.data)


class OracleGroupWide(OracleGroup):
    """The 16-wide batched restatement (oso2cpp.WideGen): same interface, same results up to FMA
    contraction; bench.py's batched CPU baseline."""

    def __init__(self, layers, connections=(), outputs=(), textures=None):
        ls = [oso2cpp.Layer(l["oso"], l["name"], l.get("params")) for l in layers]
        self.group = oso2cpp.Group(ls, connections, outputs)
        self.so = oso2cpp.build_group_wide(self.group)
        self.lib = ctypes.CDLL(self.so)
        self.lib.oracle_run_mt.argtypes = [ctypes.POINTER(Launch), ctypes.c_longlong, ctypes.c_int]
        self.lib.oracle_run_capture.argtypes = [ctypes.POINTER(Launch), ctypes.c_longlong,
                                                ctypes.c_longlong]
        self.lib.oracle_run_capture.restype = ctypes.c_char_p
        register_textures(self.lib, textures)


class OracleRender:
    def __init__(self, scene, arrays, oso_lookup, opt="-O2"):
        self.scene, self.arrays = scene, arrays
        self.textures = scene_textures(scene)
        groups = []
        for layers, conns in material_groups(scene, oso_lookup):
            ls = [oso2cpp.Layer(l["oso"], l["name"], l["params"]) for l in layers]
            groups.append(oso2cpp.Group(ls, conns, ()))
        self.so = oso2cpp.build_render(groups, opt=opt)
        self.lib = ctypes.CDLL(self.so)
        self.lib.oracle_render.argtypes = [ctypes.POINTER(RenderScene), ctypes.c_void_p, ctypes.c_int]
        register_textures(self.lib, self.textures)
        if os.path.exists(BSDL_LUTS):
            self._luts = bsdl_luts()
            self.lib.oracle_set_bsdl_luts.argtypes = [ctypes.c_void_p]
            self.lib.oracle_set_bsdl_luts(self._luts.ctypes.data)

    def render(self, xres, yres, aa, nthreads=None, rows=None, **kw):
        """The whole image, or with rows=(y0, y1) only that band of it (the other rows stay 0;
        pixels are independent, so a band equals the same rows of a full render)."""
        rs, keep = fill_render_scene(RenderScene, self.scene, self.arrays, xres, yres, aa, **kw)
        out = np.zeros((yres, xres, 3), np.float32)
        y0, y1 = rows if rows else (0, yres)
        self.lib.oracle_render_band(ctypes.c_void_p(ctypes.addressof(rs)), ctypes.c_void_p(out.ctypes.data),
                                    int(nthreads or (os.cpu_count() or 1)), int(y0), int(y1))
        return out
