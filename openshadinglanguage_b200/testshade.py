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


def harness_attributes(xres, yres):
    """What testshade's SimpleRenderer answers to get_attribute for every point (simplerend.cpp:246-266,
    330-346: camera set up as perspective, fov 90, hither 0.1, yon 1000 by testshade.cpp): the uniform
    renderer attributes handed to the group as b200_attribute entries."""
    aspect = float(np.float32(xres) / np.float32(yres))
    return {"camera:resolution": [int(xres), int(yres)], "camera:projection": "perspective", "camera:fov": 90.0,
            "camera:pixelaspect": 1.0, "camera:clip_near": 0.1, "camera:clip_far": 1000.0, "camera:clip": [0.1, 1000.0],
            "camera:shutter_open": 0.0, "camera:shutter_close": 1.0, "camera:shutter": [0.0, 1.0],
            "camera:screen_window": [-aspect, -1.0, aspect, 1.0]}


# ---------------------------------------------------------------------------------------------
# The testshade command line (src/testshade/testshade.cpp:705-905, 495-700) -> a group run
# description, so that the reference's testsuite commands can be replayed through this library
# (tools/testsuite_b200.py, tests/test_testsuite_b200.py).  Only what describes the group and the
# grid is interpreted; flags that tune the reference's optimizer / JIT are accepted and ignored,
# flags this harness does not implement are listed in spec["unsupported"].
# ---------------------------------------------------------------------------------------------
import os
import re
import shlex

_IGNORED_FLAGS = {"-t": 1, "--threads": 1, "-O0": 0, "-O1": 0, "-O2": 0, "--llvm_opt": 1, "--batched": 0,
                  "--stats": 0, "--runstats": 0, "--debugnan": 0, "--debuguninit": 0, "--no-output-placement": 0,
                  "--shadeimage": 0, "--noshadeimage": 0, "--jbufferMB": 1,
                  "--warmup": 0, "--locale": 1, "--raytype_opt": 0, "--groupoutputs": 0, "--use_rs_bitcode": 0, "--texoptions": 1, "-texoptions": 1}
_UNSUPPORTED_FLAGS = {"-v": 0, "--debug": 0, "--debug2": 0, "--archivegroup": 1,
                      "--entry": 1, "--entryoutput": 1, "--oslquery": 0, "--print-groupdata": 0,
                      "--print-group-stats": 0, "--inbuffer": 0, "--expr": 1, "-expr": 1, "--profile": 0}


def _float_list(s, n):
    parts = [p for p in re.split(r"[,\s]+", s.strip()) if p]
    if len(parts) != n:
        return None
    try:
        return [float(p) for p in parts]
    except ValueError:
        return None


def parse_param_value(command, value):
    """add_param (testshade.cpp:494-600): the value's type from its spelling, or from a
    ":type=" hint on the flag.  -> (value, hints dict)"""
    hints = {}
    typ = None
    for opt in command.split(":")[1:]:
        k, _, v = opt.partition("=")
        if k == "type":
            typ = v.strip()
        elif k == "lockgeom":
            hints["interpolated"] = not int(v)
        elif k == "interpolated":
            hints["interpolated"] = bool(int(v))
        elif k == "interactive":
            hints["interactive"] = bool(int(v))
    triple = typ in ("color", "point", "vector", "normal")
    if typ in (None, "matrix"):
        f = _float_list(value, 16)
        if f is not None:
            return f, hints
    if typ is None or triple:
        f = _float_list(value, 3)
        if f is not None:
            return f, hints
    if typ in (None, "int") and re.fullmatch(r"\s*[-+]?\d+\s*", value):
        return int(value), hints
    if typ in (None, "float"):
        try:
            return float(value), hints
        except ValueError:
            pass
    if typ is not None:
        m = re.fullmatch(r"(float|int|string|color|point|vector|normal)(?:\[(\d+)\])?", typ)
        if m and m.group(1) != "string":
            conv = int if m.group(1) == "int" else float
            return [conv(p) for p in re.split(r"[,\s]+", value.strip()) if p], hints
        if m and m.group(2):
            return (value.split(",") + [""] * int(m.group(2)))[:int(m.group(2))], hints
    return value, hints


def parse_command(argstr):
    """One `testshade ...` command line -> spec dict (see run_command)."""
    argv = shlex.split(argstr)
    spec = dict(xres=1, yres=1, center=False, layers=[], connections=[], outputs=[], dataformat=None,
                vary_pdxdy=False, vary_udxdy=False, vary_vdxdy=False, raytype="camera", iters=1,
                uscale=1.0, vscale=1.0, uoffset=0.0, voffset=0.0, userdata=[], options="",
                userdata_isconnected=False, unsupported=[], groupname="", colorspace="", print=False, reparams=[])
    pending, layername = {}, None
    i = 0

    def take(n):
        nonlocal i
        vals = argv[i + 1:i + 1 + n]
        if len(vals) != n:
            raise ValueError("testshade: %s needs %d argument(s)" % (argv[i], n))
        i += n
        return vals

    def declare(shader, name):
        nonlocal pending, layername
        spec["layers"].append(dict(shader=shader, name=name or "%s_%d" % (shader, len(spec["layers"])),
                                   params=pending))
        pending, layername = {}, None

    while i < len(argv):
        a = argv[i]
        base = a.split(":")[0]
        if base in ("-g", "--res"):
            spec["xres"], spec["yres"] = (int(v) for v in take(2))
        elif base in ("--center", "-center"):
            spec["center"] = True
        elif base in ("-od", "-d"):
            spec["dataformat"] = take(1)[0]
        elif base == "-o":
            name, fn = take(2)
            spec["outputs"].append((name, fn))
        elif base in ("--layer", "-layer"):
            layername = take(1)[0]
        elif base in ("--param", "-param"):
            name, value = take(2)
            v, hints = parse_param_value(a, value)
            if hints.get("interpolated"):
                spec["unsupported"].append("param hint lockgeom=0")
            pending[name] = v
        elif base in ("--shader", "-shader"):
            shader, name = take(2)
            declare(shader, name)
        elif base in ("--connect", "-connect"):
            spec["connections"].append(tuple(take(4)))
        elif base in ("--options", "-options"):
            spec["options"] = take(1)[0]
        elif base in ("--vary_pdxdy", "--vary_udxdy", "--vary_vdxdy"):
            spec[base[2:]] = True
        elif base in ("--raytype", "-raytype"):
            spec["raytype"] = take(1)[0]
        elif base in ("--group", "-group"):
            # a serialized group: "param type name values ; shader name layer ; connect a.b c.d ;"
            # (ShaderGroupBegin(name, usage, groupspec), oslexec.h:634-650); the text itself or a file holding it
            from .render.scene import parse_group_spec
            text = take(1)[0]
            if os.path.exists(text):
                text = open(text).read()
            if re.search(r"\[\d+\]\s*(?:[;,]|$)|\.\w+\[\d+\]", text):
                spec["unsupported"].append("-group with component connections")
            else:
                glayers, gconns = parse_group_spec(text)
                for l in glayers:
                    spec["layers"].append(dict(shader=l["shader"], name=l["name"], params=l["params"]))
                spec["connections"] += [tuple(c) for c in gconns]
        elif base in ("--reparam", "-reparam"):
            # ShadingSystem::ReParameter after the first iteration (testshade.cpp:2245-2255)
            layer, name, value = take(3)
            spec["reparams"].append((layer, name, parse_param_value(a, value)[0]))
        elif base == "--print":
            spec["print"] = True                  # print every output value per pixel (save_outputs, testshade.cpp:1246-1290)
        elif base in ("--colorspace", "-colorspace"):
            spec["colorspace"] = take(1)[0]       # ShadingSystem attribute "colorspace"
        elif base in ("--groupname", "-groupname"):
            spec["groupname"] = take(1)[0]
        elif base in ("--iters", "-iters"):
            spec["iters"] = int(take(1)[0])
        elif base in ("--scaleuv", "-scaleuv", "--scalest"):
            spec["uscale"], spec["vscale"] = (float(v) for v in take(2))
        elif base in ("--offsetuv", "-offsetuv", "--offsetst"):
            spec["uoffset"], spec["voffset"] = (float(v) for v in take(2))
        elif base in ("--userdata", "-userdata"):
            name, value = take(2)
            spec["userdata"].append((name, parse_param_value(a, value)[0]))
        elif base == "--userdata_isconnected":
            spec["userdata_isconnected"] = True
            spec["unsupported"].append(base)
        elif base in _IGNORED_FLAGS:
            take(_IGNORED_FLAGS[base])
        elif base in _UNSUPPORTED_FLAGS:
            take(_UNSUPPORTED_FLAGS[base])
            spec["unsupported"].append(base)
        elif a.startswith("-") and not re.fullmatch(r"-?[\d.]+", a):
            spec["unsupported"].append(a)
        else:
            declare(a[:-4] if a.endswith(".oso") else a, layername)
        i += 1
    return spec


_OSO_PARAM = re.compile(r"^(param|oparam)\s+(closure color|\w+)(\[\d*\])?\s+(\S+)", re.M)
_NCOMP = {"float": 1, "int": 1, "color": 3, "point": 3, "vector": 3, "normal": 3, "matrix": 16}


def oso_params(oso_text):
    """[(name, base type, array length or 0, is output)] of a compiled shader (the OSLQuery view
    testshade uses to type its -o outputs, testshade.cpp:1094-1118)."""
    out = []
    for m in _OSO_PARAM.finditer(oso_text):
        arr = m.group(3)
        out.append((m.group(4), m.group(2), int(arr[1:-1] or 0) if arr else 0, m.group(1) == "oparam"))
    return out


def run_command(spec, oso, make_group, globals_fn, userdata_fn=None):
    """Replay one parsed testshade command.

    oso(shader) -> .oso text; make_group(layers, connections, outputs, userdata_descs) -> object with
    run(n, varying, uniform, out, userdata_arena) -> journal text; globals_fn = grid_globals.
    -> dict(text=<what testshade prints>, images={output name: float32 [yres, xres, nchan]})"""
    import numpy as np
    if spec["unsupported"]:
        raise NotImplementedError("testshade flags not replayed: " + " ".join(sorted(set(spec["unsupported"]))))
    if not spec["layers"]:
        raise ValueError("testshade: no shader given")
    layers = [dict(oso=oso(l["shader"]), name=l["name"], params=l["params"]) for l in spec["layers"]]
    lines = []
    for (sl, sp, dl, dp) in spec["connections"]:
        lines.append("Connect %s.%s to %s.%s" % (sl, sp, dl, dp))
    # outputs: -o VAR FILE (default Cout -> null); typed by looking the name up back to front
    wanted = spec["outputs"] or [("Cout", "null")]
    if spec["outputs"]:
        lines.append("")          # testshade.cpp:2113-2114
    outputs, images, offset = [], [], 0
    for var, fn in wanted:
        layer, _, pname = var.rpartition(".")
        for l in reversed(layers):
            if layer and l["name"] != layer:
                continue
            hit = [p for p in oso_params(l["oso"]) if p[3] and p[0] == pname]
            if hit:
                _, base, arr, _ = hit[0]
                if base not in _NCOMP or arr:
                    raise NotImplementedError("output %s of type %s%s" % (var, base, "[%d]" % arr if arr else ""))
                nch = _NCOMP[base]
                if fn != "null":
                    lines.append("Output %s to %s" % (var, fn))
                images.append((var, fn, offset, nch, base == "int"))
                offset += 4 * nch
                break
    n = spec["xres"] * spec["yres"]
    raybits = {"camera": 1, "shadow": 2, "reflection": 4, "refraction": 8, "diffuse": 16, "glossy": 32,
               "subsurface": 64, "displacement": 128}
    var, uni = globals_fn(spec["xres"], spec["yres"], center=spec["center"], vary_udxdy=spec["vary_udxdy"],
                          vary_vdxdy=spec["vary_vdxdy"], vary_pdxdy=spec["vary_pdxdy"], uscale=spec["uscale"],
                          vscale=spec["vscale"], uoffset=spec["uoffset"], voffset=spec["voffset"],
                          raytype_bit=raybits.get(spec["raytype"], 1))
    # one planar arena: output k occupies [offset_k * n, (offset_k + 4 nch) * n), stride = its own size,
    # exactly how testshade places its ImageBufs behind one another (testshade.cpp:1135-1155)
    arena = np.zeros(max(1, offset // 4 * n), np.float32)
    outs = [dict(name=v, offset=off * n, stride=4 * nch, derivs=False) for v, _, off, nch, _ in images]
    g = make_group(layers, spec["connections"], outs, spec)
    text = g.run(n, var, uni, arena)
    body = text[:-1] if text.endswith("\n") else text     # keep the empty lines the error handler adds
    out = "\n".join(lines + ([body] if text.strip("\n") else [])) * 1
    if spec["iters"] > 1 and spec["reparams"]:
        # ReParameter between the first and the second iteration: the later iterations (whose outputs
        # are the ones saved) see the new instance values.  The group is rebuilt with them here; a
        # renderer with "interactive" parameters does the same through ShadingSystem::ReParameter.
        layers2 = [dict(l, params=dict(l["params"])) for l in layers]
        for lname, pname, value in spec["reparams"]:
            for l in layers2:
                if l["name"] == lname:
                    l["params"][pname] = value
        text2 = make_group(layers2, spec["connections"], outs, spec).run(n, var, uni, arena)
        body2 = text2[:-1] if text2.endswith("\n") else text2
        out = "\n".join(lines + ([body] if text.strip("\n") else [])
                         + ([body2] * (spec["iters"] - 1) if text2.strip("\n") else []))
    elif spec["iters"] > 1 and text.strip("\n"):
        out = "\n".join(lines + [body] * spec["iters"])
    if spec["print"]:
        if text.strip("\n"):
            raise NotImplementedError("--print together with shader printf output (interleaved per point)")
        plines = []
        for y in range(spec["yres"]):
            for x in range(spec["xres"]):
                plines.append("Pixel (%d, %d):" % (x, y))
                for v, fn, off, nch, is_int in images:
                    a = arena[off // 4 * n:(off // 4 + nch) * n].reshape(n, nch)[y * spec["xres"] + x]
                    vals = a.view(np.int32) if is_int else a
                    plines.append("  %s :%s" % (v, "".join((" %d" % t) if is_int else (" %g" % t) for t in vals)))
        out = "\n".join([out] + plines) if out else "\n".join(plines)
    if not spec["outputs"]:
        # without -o testshade shades the default output "Cout" into a null image and its run ends with an
        # empty line (visible in goldens of several commands: testsuite/error-dupes, getattribute-shader)
        out += "\n"
    imgs = {}
    for v, fn, off, nch, is_int in images:
        a = arena[off // 4 * n:(off // 4 + nch) * n].reshape(n, nch)
        if is_int:
            a = a.view(np.int32).astype(np.float32)
        # keyed by (variable, file): the same variable may be written to several files (testsuite/shaderglobals)
        imgs[v if fn == "null" else v + "|" + fn] = (fn, a.reshape(spec["yres"], spec["xres"], nch).copy())
    return dict(text=out, images=imgs)


def compare_image(got, ref, kind, failthresh=0.004, failpercent=0.02, hardfail=0.012):
    """idiff as runtest.py uses it: a pixel fails when it is off by more than failthresh; the image
    fails when more than failpercent % of its pixels do, or any is off by more than hardfail.
    `kind`: the file's data format ("uint8" / "uint16": `got` is quantised like the image writer does).
    -> None or a description of the difference."""
    import numpy as np
    if got.shape[:2] != ref.shape[:2]:
        return "image size %s != %s" % (got.shape, ref.shape)
    nch = min(got.shape[2], ref.shape[2])
    g = got[..., :nch]
    if kind == "uint8":
        g = np.round(np.clip(g, 0, 1) * 255.0) / 255.0
    elif kind == "uint16":
        g = np.round(np.clip(g, 0, 1) * 65535.0) / 65535.0
    elif kind == "float":
        with np.errstate(over="ignore"):
            if np.array_equal(ref, ref.astype(np.float16).astype(np.float32)):
                g = g.astype(np.float16).astype(np.float32)      # a half-float file (-od half): written through half
    d = np.abs(g - ref[..., :nch]).max(axis=2)
    bad = (d > failthresh).mean() * 100.0
    if bad > failpercent or d.max() > max(hardfail, failthresh):
        return "image differs: %.3f %% of pixels beyond %g, max %g" % (bad, failthresh, d.max())
    return None
