"""ctypes binding of include/osl_b200.h (host-side plumbing, not the product).

Mirrors the reference call sequence a renderer makes
(ShaderGroupBegin / Parameter / Shader / ConnectShaders / ShaderGroupEnd, then
execute with ShaderGlobals + output arena; src/include/OSL/oslexec.h:634-1033)
one-to-one on the C ABI.  Device memory and streams come from PyTorch.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libosl_b200.so")

SG_FIELDS = ["P", "dPdx", "dPdy", "dPdz", "I", "dIdx", "dIdy", "N", "Ng", "u", "dudx",
             "dudy", "v", "dvdx", "dvdy", "dPdu", "dPdv", "time", "dtime", "dPdtime",
             "Ps", "dPsdx", "dPsdy", "surfacearea", "raytype", "flipHandedness",
             "backfacing"]
SG_VEC = {"P", "dPdx", "dPdy", "dPdz", "I", "dIdx", "dIdy", "N", "Ng", "dPdu", "dPdv",
          "dPdtime", "Ps", "dPsdx", "dPsdy"}
SG_INT = {"raytype", "flipHandedness", "backfacing"}
NF = len(SG_FIELDS)


class B200Error(RuntimeError):
    pass


class _Transform(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("m", ctypes.c_float * 16)]


class _Globals(ctypes.Structure):
    _fields_ = [("varying", ctypes.c_void_p * NF),
                ("uniform", (ctypes.c_float * 4) * NF),
                ("plane_stride", ctypes.c_longlong),
                ("ntransforms", ctypes.c_int),
                ("transforms", ctypes.POINTER(_Transform))]


class _Param(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("type", ctypes.c_int), ("nvalues", ctypes.c_int),
                ("values", ctypes.c_void_p)]


class _Layer(ctypes.Structure):
    _fields_ = [("oso_text", ctypes.c_char_p), ("layername", ctypes.c_char_p),
                ("nparams", ctypes.c_int), ("params", ctypes.POINTER(_Param))]


class _Connection(ctypes.Structure):
    _fields_ = [("srclayer", ctypes.c_char_p), ("srcparam", ctypes.c_char_p),
                ("dstlayer", ctypes.c_char_p), ("dstparam", ctypes.c_char_p)]


class _SymLoc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("offset", ctypes.c_longlong),
                ("stride", ctypes.c_longlong), ("derivs", ctypes.c_int)]


class _UserData(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("ncomp", ctypes.c_int), ("is_int", ctypes.c_int),
                ("offset", ctypes.c_longlong), ("stride", ctypes.c_longlong), ("derivs", ctypes.c_int),
                ("valid_offset", ctypes.c_longlong), ("valid_stride", ctypes.c_longlong)]


class _GroupDesc(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("nlayers", ctypes.c_int),
                ("layers", ctypes.POINTER(_Layer)), ("nconnections", ctypes.c_int),
                ("connections", ctypes.POINTER(_Connection)), ("noutputs", ctypes.c_int),
                ("outputs", ctypes.POINTER(_SymLoc)), ("options", ctypes.c_char_p),
                ("nuserdata", ctypes.c_int), ("userdata", ctypes.POINTER(_UserData)),
                ("nattributes", ctypes.c_int), ("attributes", ctypes.POINTER(_Param))]   # b200_attribute = b200_param layout


def pack_userdata(entries):
    """Lay named per-point userdata out as one arena (b200_userdata, include/osl_b200.h).
    entries: list of dict(name, data [n, ncomp] float32 (or [n, 3*ncomp] = val,dx,dy with
    derivs=True) or int32 [n]; valid=int32[n] marks the points that have the value).
    -> (arena uint8 numpy array, list of descriptor dicts for ShaderGroup(userdata=...)).
    One dense array per entry (stride = its row size), so a warp's loads coalesce."""
    blobs, descs, off = [], [], 0
    for e in entries:
        a = np.ascontiguousarray(e["data"])
        is_int = a.dtype.kind in "iu"
        a = a.astype(np.int32 if is_int else np.float32).reshape(len(a), -1)
        derivs = bool(e.get("derivs"))
        d = dict(name=e["name"], ncomp=a.shape[1] // (3 if derivs else 1), is_int=int(is_int), offset=off,
                 stride=a.shape[1] * 4, derivs=int(derivs), valid_offset=-1, valid_stride=0)
        b = a.tobytes()
        b = b.ljust((len(b) + 15) // 16 * 16, b"\0")
        blobs.append(b)
        off += len(b)
        if e.get("valid") is not None:
            v = np.ascontiguousarray(e["valid"], np.int32).tobytes()
            v = v.ljust((len(v) + 15) // 16 * 16, b"\0")
            d["valid_offset"], d["valid_stride"] = off, 4
            blobs.append(v)
            off += len(v)
        descs.append(d)
    return np.frombuffer(b"".join(blobs) or b"\0" * 16, np.uint8).copy(), descs


_lib = None


def library_path():
    return _LIBPATH


def lib():
    """Load libosl_b200.so; raise if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise B200Error("libosl_b200.so is not built: run `python -m openshadinglanguage_b200.build` "
                        "(or __graft_entry__.build()); the product has no CPU fallback")
    L = ctypes.CDLL(_LIBPATH)
    L.b200_last_error.restype = ctypes.c_char_p
    L.b200_group_compile.argtypes = [ctypes.POINTER(_GroupDesc), ctypes.POINTER(ctypes.c_void_p)]
    L.b200_group_destroy.argtypes = [ctypes.c_void_p]
    L.b200_group_cuda_source.argtypes = [ctypes.c_void_p]
    L.b200_group_cuda_source.restype = ctypes.c_char_p
    L.b200_group_cubin.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
    L.b200_group_cubin.restype = ctypes.c_void_p
    L.b200_group_num_warnings.argtypes = [ctypes.c_void_p]
    L.b200_group_warning.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.b200_group_warning.restype = ctypes.c_char_p
    L.b200_group_reads_global.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.b200_group_execute.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong,
                                     ctypes.POINTER(_Globals), ctypes.c_void_p, ctypes.c_void_p,
                                     ctypes.c_void_p]
    L.b200_group_execute_host.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                          ctypes.POINTER(_Globals), ctypes.c_void_p]
    L.b200_launch_count.restype = ctypes.c_longlong
    L.b200_shadeop_noise.argtypes = [ctypes.c_int] * 5 + [ctypes.c_longlong, ctypes.c_void_p,
                                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    L.b200_shadeop_hash.argtypes = [ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_void_p]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise B200Error("libosl_b200 error %d: %s" % (rc, lib().b200_last_error().decode(errors="replace")))


def launch_count():
    return int(lib().b200_launch_count())


def add_texture(name, pixels):
    """Register a decoded image ([h, w] or [h, w, nch] float32, top scanline first) under the
    file name shaders pass to texture() (b200_texture_add)."""
    import numpy as np
    a = np.ascontiguousarray(pixels, dtype=np.float32)
    if a.ndim == 2:
        a = a[..., None]
    L = lib()
    L.b200_texture_add.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    _check(L.b200_texture_add(name.encode(), a.shape[1], a.shape[0], a.shape[2], a.ctypes.data))


def _ptr(x):
    """raw address of a torch tensor / numpy array / int / None"""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ctypes.data


def _fill_globals(n, varying, uniform, plane_stride=None):
    g = _Globals()
    for i, f in enumerate(SG_FIELDS):
        g.varying[i] = _ptr(varying.get(f)) if varying else None
        vals = list((uniform or {}).get(f, []))
        for c in range(4):
            if f in SG_INT:
                iv = int(vals[0]) if vals else 0
                g.uniform[i][c] = float(np.array([iv], np.int32).view(np.float32)[0]) if c == 0 else 0.0
            else:
                g.uniform[i][c] = float(vals[c]) if c < len(vals) else 0.0
    g.plane_stride = n if plane_stride is None else plane_stride
    # named coordinate systems ("shader", "object", renderer-named spaces): {name: 16 floats}
    xf = (uniform or {}).get("transforms") or {}
    arr = (_Transform * max(1, len(xf)))()
    for k, (name, m) in enumerate(xf.items()):
        arr[k].name = name.encode()
        for c in range(16):
            arr[k].m[c] = float(m[c])
    g.ntransforms = len(xf)
    g.transforms = arr
    g._keep_transforms = arr   # keep the array alive as long as the struct
    return g


class ShaderGroup:
    """One compiled shader group.

    layers:       [dict(oso=<.oso text>, name=<layer name>, params={name: value(s)})]
    connections:  [(srclayer, srcparam, dstlayer, dstparam)]
    outputs:      [dict(name='layer.param'|'param', offset=, stride=, derivs=False)]
    options:      'fma=0' selects strict IEEE evaluation (bit-parity mode)
    """

    def __init__(self, layers, connections=(), outputs=(), options="", name="group", userdata=(), attributes=None):
        """userdata: descriptor dicts (pack_userdata) of the per-point values the renderer supplies
        for interpolated ([[ int lockgeom = 0 ]]) parameters.
        attributes: {name: value(s)} uniform renderer attributes getattribute() can return
        (RendererServices::get_attribute for values that do not vary over the batch; b200_attribute)."""
        L = lib()
        keep = []

        def cs(s):
            b = s.encode() if isinstance(s, str) else s
            keep.append(b)
            return b
        clayers = (_Layer * len(layers))()
        for i, l in enumerate(layers):
            params = l.get("params") or {}
            cp = (_Param * max(1, len(params)))()
            for j, (k, v) in enumerate(params.items()):
                if not isinstance(v, (list, tuple, np.ndarray)):
                    v = [v]
                if isinstance(v[0], str):
                    arr = (ctypes.c_char_p * len(v))(*[cs(x) for x in v])
                    t = 2
                elif isinstance(v[0], (int, np.integer)) and not isinstance(v[0], bool):
                    arr = (ctypes.c_int * len(v))(*[int(x) for x in v])
                    t = 0
                else:
                    arr = (ctypes.c_float * len(v))(*[float(x) for x in v])
                    t = 1
                keep.append(arr)
                cp[j].name, cp[j].type, cp[j].nvalues = cs(k), t, len(v)
                cp[j].values = ctypes.cast(arr, ctypes.c_void_p)
            keep.append(cp)
            clayers[i].oso_text, clayers[i].layername = cs(l["oso"]), cs(l["name"])
            clayers[i].nparams, clayers[i].params = len(params), cp
        cconn = (_Connection * max(1, len(connections)))()
        for i, (a, b, c, d) in enumerate(connections):
            cconn[i].srclayer, cconn[i].srcparam, cconn[i].dstlayer, cconn[i].dstparam = cs(a), cs(b), cs(c), cs(d)
        cout = (_SymLoc * max(1, len(outputs)))()
        for i, o in enumerate(outputs):
            cout[i].name, cout[i].offset, cout[i].stride = cs(o["name"]), int(o["offset"]), int(o["stride"])
            cout[i].derivs = 1 if o.get("derivs") else 0
        cud = (_UserData * max(1, len(userdata)))()
        for i, u in enumerate(userdata):
            cud[i].name, cud[i].ncomp, cud[i].is_int = cs(u["name"]), int(u["ncomp"]), int(u["is_int"])
            cud[i].offset, cud[i].stride, cud[i].derivs = int(u["offset"]), int(u["stride"]), int(u["derivs"])
            cud[i].valid_offset, cud[i].valid_stride = int(u["valid_offset"]), int(u["valid_stride"])
        attributes = attributes or {}
        cat = (_Param * max(1, len(attributes)))()
        for j, (k, v) in enumerate(attributes.items()):
            if not isinstance(v, (list, tuple, np.ndarray)):
                v = [v]
            if isinstance(v[0], str):
                arr, t = (ctypes.c_char_p * len(v))(*[cs(x) for x in v]), 2
            elif isinstance(v[0], (int, np.integer)) and not isinstance(v[0], bool):
                arr, t = (ctypes.c_int * len(v))(*[int(x) for x in v]), 0
            else:
                arr, t = (ctypes.c_float * len(v))(*[float(x) for x in v]), 1
            keep.append(arr)
            cat[j].name, cat[j].type, cat[j].nvalues = cs(k), t, len(v)
            cat[j].values = ctypes.cast(arr, ctypes.c_void_p)
        desc = _GroupDesc(cs(name), len(layers), clayers, len(connections), cconn, len(outputs), cout,
                          cs(options), len(userdata), cud, len(attributes), cat)
        h = ctypes.c_void_p()
        _check(L.b200_group_compile(ctypes.byref(desc), ctypes.byref(h)))
        self._h = h
        self.outputs = list(outputs)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().b200_group_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def cuda_source(self):
        return lib().b200_group_cuda_source(self._h).decode()

    @property
    def cubin(self):
        n = ctypes.c_longlong()
        p = lib().b200_group_cubin(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value)

    @property
    def warnings(self):
        L = lib()
        return [L.b200_group_warning(self._h, i).decode() for i in range(L.b200_group_num_warnings(self._h))]

    def reads_global(self, name):
        return bool(lib().b200_group_reads_global(self._h, SG_FIELDS.index(name)))

    def execute(self, n, varying, uniform, output, shadeindex=None, device=0, stream=None,
                plane_stride=None, userdata=None):
        """Device-pointer path (b200_group_execute).  varying values, output and the userdata
        arena are CUDA tensors (or raw device addresses).  Asynchronous."""
        g = _fill_globals(n, varying, uniform, plane_stride)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(device).cuda_stream
        _check(lib().b200_group_execute(self._h, device, ctypes.c_void_p(stream), n, ctypes.byref(g),
                                        _ptr(shadeindex), _ptr(userdata), _ptr(output)))

    def bind(self, n, varying, uniform, output, shadeindex=None, device=0, stream=None, plane_stride=None):
        """-> zero-argument callable that issues b200_group_execute with the globals block,
        pointers and stream bound once - a renderer reusing its ShaderGlobals batch between
        launches.  The per-call cost is one foreign call."""
        g = _fill_globals(n, varying, uniform, plane_stride)
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(device).cuda_stream
        fn, h = lib().b200_group_execute, self._h
        args = (h, device, ctypes.c_void_p(stream), n, ctypes.byref(g), _ptr(shadeindex), None, _ptr(output))

        def launch(_keep=(g, varying, output, shadeindex)):
            rc = fn(*args)
            if rc:
                _check(rc)
        return launch

    def execute_host(self, n, varying, uniform, output, device=0, plane_stride=None, userdata=None):
        """Host-pointer path (b200_group_execute_host[_userdata]): numpy arrays or pinned
        CPU tensors in, host output arena out.  Synchronous."""
        g = _fill_globals(n, varying, uniform, plane_stride)
        if userdata is None:
            _check(lib().b200_group_execute_host(self._h, device, n, ctypes.byref(g), _ptr(output)))
        else:
            L = lib()
            L.b200_group_execute_host_userdata.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong,
                                                           ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong,
                                                           ctypes.c_void_p]
            _check(L.b200_group_execute_host_userdata(self._h, device, n, ctypes.byref(g), _ptr(userdata),
                                                      int(userdata.nbytes), _ptr(output)))

    def journal(self, device=0):
        """Text printed by the group's printf() ops since the previous call, in shade index
        order (b200_group_journal).  Synchronises the device."""
        L = lib()
        L.b200_group_journal.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.b200_group_journal.restype = ctypes.c_char_p
        return L.b200_group_journal(self._h, device).decode(errors="replace")


def shadeop_noise(kind, outdim, indim, n, inp, out, period=None, derivs=False, stream=None):
    """Batch noise over device SoA planes (b200_shadeop_noise)."""
    kinds = {"noise": 0, "snoise": 1, "cellnoise": 2, "hashnoise": 3, "simplex": 4, "usimplex": 5}
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _check(lib().b200_shadeop_noise(kinds[kind], outdim, indim, 1 if derivs else 0, 0, n, _ptr(inp),
                                    _ptr(period), _ptr(out), ctypes.c_void_p(stream)))


def shadeop_hash(indim, n, inp, out, stream=None):
    if stream is None:
        import torch
        stream = torch.cuda.current_stream().cuda_stream
    _check(lib().b200_shadeop_hash(indim, n, _ptr(inp), _ptr(out), ctypes.c_void_p(stream)))


# ---------------------------------------------------------------------------
# wavefront path tracer (b200_render_*)
# ---------------------------------------------------------------------------
class _RenderScene(ctypes.Structure):
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


class _RenderStats(ctypes.Structure):
    _fields_ = [("paths", ctypes.c_longlong), ("launches", ctypes.c_longlong),
                ("bounce_iterations", ctypes.c_longlong), ("device_ms", ctypes.c_double),
                ("tail_ms", ctypes.c_double), ("slots", ctypes.c_longlong), ("rounds", ctypes.c_longlong)]


def _group_desc(layers, connections, name, options, keep):
    """Build a _GroupDesc (shared by ShaderGroup and Renderer)."""
    def cs(s):
        b = s.encode() if isinstance(s, str) else s
        keep.append(b)
        return b
    clayers = (_Layer * len(layers))()
    for i, l in enumerate(layers):
        params = l.get("params") or {}
        cp = (_Param * max(1, len(params)))()
        for j, (k, v) in enumerate(params.items()):
            if not isinstance(v, (list, tuple, np.ndarray)):
                v = [v]
            if isinstance(v[0], str):
                arr = (ctypes.c_char_p * len(v))(*[cs(x) for x in v])
                t = 2
            elif isinstance(v[0], (int, np.integer)) and not isinstance(v[0], bool):
                arr = (ctypes.c_int * len(v))(*[int(x) for x in v])
                t = 0
            else:
                arr = (ctypes.c_float * len(v))(*[float(x) for x in v])
                t = 1
            keep.append(arr)
            cp[j].name, cp[j].type, cp[j].nvalues = cs(k), t, len(v)
            cp[j].values = ctypes.cast(arr, ctypes.c_void_p)
        keep.append(cp)
        clayers[i].oso_text, clayers[i].layername = cs(l["oso"]), cs(l["name"])
        clayers[i].nparams, clayers[i].params = len(params), cp
    cconn = (_Connection * max(1, len(connections)))()
    for i, (a, b, c, d) in enumerate(connections):
        cconn[i].srclayer, cconn[i].srcparam, cconn[i].dstlayer, cconn[i].dstparam = cs(a), cs(b), cs(c), cs(d)
    keep += [clayers, cconn]
    return _GroupDesc(cs(name), len(layers), clayers, len(connections), cconn, 0, None, cs(options), 0, None, 0, None)


class Renderer:
    """testrender mirror: scene (openshadinglanguage_b200.render.scene.Scene),
    its prepared arrays, and a lookup  shader name -> .oso text."""

    def __init__(self, scene, arrays, oso_lookup, xres, yres, aa, max_bounces=None, rr_depth=None,
                 no_jitter=False, show_globals=0, options=""):
        # testrender defaults (simpleraytracer.cpp:1226), overridden by the scene's <Option>
        if max_bounces is None:
            max_bounces = scene.options.get("max_bounces", 1000000)
        if rr_depth is None:
            rr_depth = scene.options.get("rr_depth", 5)
        L = lib()
        L.b200_render_create.argtypes = [ctypes.POINTER(_RenderScene), ctypes.c_int, ctypes.POINTER(_GroupDesc),
                                         ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.b200_render_destroy.argtypes = [ctypes.c_void_p]
        L.b200_render_cuda_source.argtypes = [ctypes.c_void_p]
        L.b200_render_cuda_source.restype = ctypes.c_char_p
        L.b200_render_rows.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.POINTER(_RenderStats)]
        L.b200_render_tiles.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                        ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_RenderStats)]
        keep = []
        rs = _RenderScene()

        def ptr(name, dtype):
            a = np.ascontiguousarray(arrays[name], dtype)
            keep.append(a)
            return a.ctypes.data if a.size else None
        rs.nverts, rs.ntris = len(arrays["verts"]), len(arrays["triangles"])
        rs.nnodes, rs.nlightprims = len(arrays["bvh_nodes"]), len(arrays["lightprims"])
        rs.nshaders, rs.nmeshes = len(scene.materials), len(arrays["mesh_surfacearea"])
        rs.verts, rs.normals, rs.uvs = ptr("verts", np.float32), ptr("normals", np.float32), ptr("uvs", np.float32)
        rs.triangles, rs.n_triangles = ptr("triangles", np.int32), ptr("n_triangles", np.int32)
        rs.uv_triangles, rs.shaderids = ptr("uv_triangles", np.int32), ptr("shaderids", np.int32)
        rs.meshids, rs.mesh_surfacearea = ptr("meshids", np.int32), ptr("mesh_surfacearea", np.float32)
        rs.bvh_nodes, rs.bvh_indices = ptr("bvh_nodes", np.float32), ptr("bvh_indices", np.uint32)
        rs.lightprims, rs.shader_is_light = ptr("lightprims", np.uint32), ptr("shader_is_light", np.int32)
        for i in range(3):
            rs.eye[i], rs.dir[i], rs.up[i] = float(scene.eye[i]), float(scene.dir[i]), float(scene.up[i])
        rs.fov = float(scene.fov)
        rs.xres, rs.yres, rs.aa = xres, yres, aa
        rs.max_bounces, rs.rr_depth = max_bounces, rr_depth
        rs.no_jitter, rs.show_globals = int(no_jitter), show_globals
        rs.background_shader, rs.background_resolution = scene.background_shader, scene.background_resolution
        mats = (_GroupDesc * len(scene.materials))()
        for k, (layers, conns) in enumerate(scene.materials):
            ls = [dict(oso=oso_lookup(l["shader"]), name=l["name"], params=l["params"]) for l in layers]
            mats[k] = _group_desc(ls, conns, "material%d" % k, "", keep)
        h = ctypes.c_void_p()
        basedir = getattr(scene, "basedir", None)
        if basedir and "texturepath=" not in options:     # texture files are relative to the scene file
            options = (options + "," if options else "") + "texturepath=" + basedir
        _check(L.b200_render_create(ctypes.byref(rs), len(scene.materials), mats, options.encode(), ctypes.byref(h)))
        self._h, self._keep = h, keep
        self.xres, self.yres, self.aa = xres, yres, aa
        self.stats = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().b200_render_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def cuda_source(self):
        return lib().b200_render_cuda_source(self._h).decode()

    @property
    def cubin(self):
        L = lib()
        L.b200_render_cubin.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_longlong)]
        L.b200_render_cubin.restype = ctypes.c_void_p
        n = ctypes.c_longlong()
        p = L.b200_render_cubin(self._h, ctypes.byref(n))
        return ctypes.string_at(p, n.value)

    def render(self, y0=0, y1=None, device=0):
        """-> float32 [(y1-y0), xres, 3]"""
        y1 = self.yres if y1 is None else y1
        out = np.zeros((y1 - y0, self.xres, 3), np.float32)
        st = _RenderStats()
        _check(lib().b200_render_rows(self._h, device, y0, y1, out.ctypes.data, ctypes.byref(st)))
        self._set_stats(st)
        return out

    def _set_stats(self, st):
        self.stats = dict(paths=st.paths, launches=st.launches, bounce_iterations=st.bounce_iterations,
                          device_ms=st.device_ms, tail_ms=st.tail_ms, slots=st.slots, rounds=st.rounds)

    def render_tiles(self, tiles, device=0, out=None):
        """Render a work set of tiles [(x0, y0, w, h), ...] -> float32 [npix, 3], pixels tile after
        tile (row-major inside a tile).  `out`: optional CUDA torch tensor [npix, 3] float32 on
        `device`; the result then stays in device memory (framebuffer gather without a host copy)."""
        t = np.ascontiguousarray(np.asarray(tiles, np.int32).reshape(-1, 4))
        npix = int((t[:, 2].astype(np.int64) * t[:, 3]).sum())
        st = _RenderStats()
        if out is None:
            res = np.zeros((npix, 3), np.float32)
            _check(lib().b200_render_tiles(self._h, device, len(t), t.ctypes.data, res.ctypes.data, 0, ctypes.byref(st)))
        else:
            assert out.is_cuda and out.is_contiguous() and out.numel() == npix * 3 and out.element_size() == 4
            res = out
            _check(lib().b200_render_tiles(self._h, device, len(t), t.ctypes.data, out.data_ptr(), 1, ctypes.byref(st)))
        self._set_stats(st)
        return res


def write_scene_blob(path, scene, arrays, oso_lookup, xres, yres, aa, max_bounces=None, rr_depth=None,
                     no_jitter=False, show_globals=0):
    """The prepared scene as the flat binary examples/testrender_b200.cpp reads: the arrays of
    b200_render_scene, the camera and options, then one group description per material."""
    import struct
    if max_bounces is None:
        max_bounces = scene.options.get("max_bounces", 1000000)
    if rr_depth is None:
        rr_depth = scene.options.get("rr_depth", 5)
    out = [struct.pack("<i", 0x42323030)]

    def arr(name, dtype):
        a = np.ascontiguousarray(arrays[name], dtype).ravel()
        out.append(struct.pack("<q", a.size))
        out.append(a.tobytes())

    def s(text):
        b = text.encode()
        out.append(struct.pack("<i", len(b)) + b)
    for name, dt in (("verts", np.float32), ("normals", np.float32), ("uvs", np.float32), ("triangles", np.int32),
                     ("n_triangles", np.int32), ("uv_triangles", np.int32), ("shaderids", np.int32),
                     ("meshids", np.int32), ("mesh_surfacearea", np.float32), ("bvh_nodes", np.float32),
                     ("bvh_indices", np.uint32), ("lightprims", np.uint32), ("shader_is_light", np.int32)):
        arr(name, dt)
    out.append(struct.pack("<10f", *[float(x) for x in list(scene.eye) + list(scene.dir) + list(scene.up)], float(scene.fov)))
    out.append(struct.pack("<9i", xres, yres, aa, max_bounces, rr_depth, int(no_jitter), show_globals,
                           scene.background_shader, scene.background_resolution))
    out.append(struct.pack("<i", len(scene.materials)))
    for layers, conns in scene.materials:
        out.append(struct.pack("<i", len(layers)))
        for l in layers:
            s(oso_lookup(l["shader"]))
            s(l["name"])
            params = l["params"] or {}
            out.append(struct.pack("<i", len(params)))
            for k, v in params.items():
                if not isinstance(v, (list, tuple, np.ndarray)):
                    v = [v]
                s(k)
                if isinstance(v[0], str):
                    out.append(struct.pack("<ii", 2, len(v)))
                    for x in v:
                        s(x)
                elif isinstance(v[0], (int, np.integer)) and not isinstance(v[0], bool):
                    out.append(struct.pack("<ii", 0, len(v)) + struct.pack("<%di" % len(v), *[int(x) for x in v]))
                else:
                    out.append(struct.pack("<ii", 1, len(v)) + struct.pack("<%df" % len(v), *[float(x) for x in v]))
        out.append(struct.pack("<i", len(conns)))
        for c in conns:
            for x in c:
                s(x)
    with open(path, "wb") as f:
        f.write(b"".join(out))


def tile_list(xres, yres, tile=64):
    """All tiles of an image, row-major, as an int32 [n, 4] array of (x0, y0, w, h)."""
    ts = [(x, y, min(tile, xres - x), min(tile, yres - y)) for y in range(0, yres, tile) for x in range(0, xres, tile)]
    return np.asarray(ts, np.int32)


def tile_pixels(tiles):
    """Image coordinates (ys, xs) of a tile work set's pixels, in the order render_tiles returns them."""
    ys, xs = [], []
    for x0, y0, w, h in np.asarray(tiles, np.int64).reshape(-1, 4):
        yy, xx = np.mgrid[y0:y0 + h, x0:x0 + w]
        ys.append(yy.ravel())
        xs.append(xx.ravel())
    return np.concatenate(ys), np.concatenate(xs)


def device_displacer(oso_lookup, device=0, options="fma=1"):
    """-> callable for render.scene.Scene.prepare(displace=...): runs a displacement group over the vertex
    batch on the GPU (one b200_group_execute; the ShaderGlobals field P comes back as a renderer output),
    the counterpart of SimpleRaytracer::prepare_geometry's per-vertex ShadingSystem::execute
    (src/testrender/simpleraytracer.cpp:1365-1384)."""
    import torch
    dev = torch.device("cuda", device) if isinstance(device, int) else device

    def run(layers, conns, g, n):
        ls = [dict(oso=oso_lookup(l["shader"]), name=l["name"], params=l["params"]) for l in layers]
        grp = ShaderGroup(ls, conns, [dict(name="P", offset=0, stride=12, derivs=False)], options=options)
        planes = {k: np.ascontiguousarray(v.T).reshape(-1) if v.ndim == 2 else np.ascontiguousarray(v)
                  for k, v in g.items()}
        dvar = {k: torch.from_numpy(v).to(dev) for k, v in planes.items()}
        out = torch.zeros((n, 3), dtype=torch.float32, device=dev)
        grp.execute(n, dvar, {}, out, device=dev.index or 0)
        torch.cuda.synchronize(dev)
        return out.cpu().numpy()
    return run

