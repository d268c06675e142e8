"""oso2cpp — CPU ORACLE (test infrastructure, NOT product code).

Turns a shader group (parsed .oso layers + instance params + connections +
renderer outputs) into plain scalar C++ that is compiled with
`g++ -O2 -ffp-contract=off` and run one shading point at a time, the way the
reference's scalar LLVM back end does.  The product's generator (C++ ->
CUDA text, openshadinglanguage_b200/csrc/host) is a separate implementation;
nothing under openshadinglanguage_b200/ imports this module.

What it restates (paths relative to /root/reference):
  * .oso grammar                      src/liboslexec/osogram.y:88-318, osolex.l
  * group assembly / connections      src/liboslexec/shadingsys.cpp:3039-3300
  * derivative propagation            src/liboslexec/runtimeoptimize.cpp:2542-2720, 3223-3231
  * lazy layers / run_lazily rules    src/liboslexec/oslexec_pvt.h:1344-1372,
                                      src/liboslexec/llvm_gen.cpp:133-283
  * layer function structure          src/liboslexec/llvm_instance.cpp:1548-1870
  * per-op semantics                  src/liboslexec/llvm_gen.cpp (cited per emitter)
"""
import hashlib
import os
import re
import struct
import subprocess

TRIPLES = ("color", "point", "vector", "normal")


class OType:
    __slots__ = ("base", "arr")

    def __init__(self, base, arr=0):
        self.base, self.arr = base, arr

    @property
    def triple(self):
        return self.base in TRIPLES

    @property
    def floatbased(self):
        return self.base in ("float", "matrix") + TRIPLES

    @property
    def ncomp(self):
        return 3 if self.triple else (16 if self.base == "matrix" else 1)

    def __repr__(self):
        return self.base + ("[%d]" % self.arr if self.arr else "")


class OSym:
    def __init__(self, name, symtype, t, vals):
        self.name, self.symtype, self.t, self.vals = name, symtype, t, vals
        self.has_derivs = False
        self.initexpr = False
        self.connected_from = None   # (layer index, sym) feeding this param
        self.connected_down = False
        self.renderer_output = None  # dict(offset, stride, derivs) when placed
        self.override = None
        self.written = False

    @property
    def isconst(self):
        return self.symtype == "const"

    @property
    def constval(self):
        """True when the runtime optimizer would have turned this symbol into
        a constant: a real const, or a lockgeom param that is neither
        connected, computed by init ops, nor written
        (runtimeoptimize.cpp find_params_holding_globals / param->const)."""
        if self.symtype == "const":
            return True
        return (self.symtype == "param" and self.connected_from is None
                and not self.initexpr and not self.written and not getattr(self, "interpolated", False))


class OOp:
    __slots__ = ("name", "args", "jumps", "rw", "derivs", "method")

    def __init__(self, name, args, jumps, rw, derivs, method):
        self.name, self.args, self.jumps, self.rw = name, args, jumps, rw
        self.derivs, self.method = derivs, method


TOK = re.compile(r'"(?:\\.|[^"\\])*"|%\w+\{(?:"(?:\\.|[^"\\])*"|[^}"])*\}|%\w+|\S+')   # a %hint{...} may quote braces


def _unescape(s):
    return (s.replace("\\n", "\n").replace("\\t", "\t").replace('\\"', '"')
            .replace("\\\\", "\\"))


class Master:
    def __init__(self):
        self.shadertype = self.name = None
        self.syms = []
        self.byname = {}
        self.ops = []
        self.methods = {}   # name -> (begin, end)


def parse_oso(text):
    m = Master()
    lines = text.split("\n")
    if not lines[0].startswith("OpenShadingLanguage"):
        raise ValueError("not an OSO file")
    method = None
    begin = 0
    for ln in lines[1:]:
        if not ln.strip() or ln.lstrip().startswith("#"):
            continue
        if ln.startswith("%preprocessed_source"):
            break
        toks = TOK.findall(ln)
        if m.name is None:
            m.shadertype, m.name = toks[0], toks[1]
            continue
        if toks[0] == "code":
            if method is not None:
                m.methods[method] = (begin, len(m.ops))
            method, begin = toks[1], len(m.ops)
            continue
        if method is None:
            # symbol declaration
            symtype = toks[0]
            i = 1
            if toks[i] == "closure":
                base = "closure color"
                i += 2
            else:
                base = toks[i]
                i += 1
            arr = 0
            mm = re.match(r"(\w+)\[(\d*)\]$", base)
            if mm:
                base, arr = mm.group(1), int(mm.group(2) or -1)
            name = toks[i]
            mm = re.match(r"(.*)\[(\d*)\]$", name)
            if mm:   # closure color x[3] style
                name, arr = mm.group(1), int(mm.group(2) or -1)
            i += 1
            vals = []
            hints = []
            for t in toks[i:]:
                if t.startswith("%"):
                    hints.append(t)
                elif t.startswith('"'):
                    vals.append(_unescape(t[1:-1]))
                elif base == "int":
                    vals.append(int(t))
                elif base == "string":
                    vals.append(t)
                else:
                    vals.append(struct.unpack("f", struct.pack("f", float(t)))[0])
            unsized = arr < 0
            if unsized:     # "type[]": as long as its default list until an instance value or a connection resizes it
                arr = max(1, len(vals) // {"matrix": 16, "color": 3, "point": 3, "vector": 3, "normal": 3}.get(base, 1))
            s = OSym(name, symtype, OType(base, arr), vals)
            s.unsized = unsized
            s.initexpr = "%initexpr" in hints
            s.interpolated = "%meta{int,lockgeom,0}" in hints   # [[ int lockgeom = 0 ]]
            m.byname[name] = s
            m.syms.append(s)
            continue
        # op line
        name = toks[0]
        if name == "end":
            continue
        args, jumps, rw, derivs = [], [], None, ()
        for t in toks[1:]:
            if t.startswith("%argrw"):
                rw = t[t.index('"') + 1:t.rindex('"')]
            elif t.startswith("%argderivs"):
                derivs = tuple(int(x) for x in t[t.index("{") + 1:-1].split(",") if x)
            elif t.startswith("%"):
                pass
            elif re.match(r"-?\d+$", t):
                jumps.append(int(t))
            else:
                args.append(m.byname[t])
        if rw is None:
            rw = "w" + "r" * (len(args) - 1) if args else ""
        m.ops.append(OOp(name, args, jumps, rw, derivs, method))
    if method is not None:
        m.methods[method] = (begin, len(m.ops))
    return m


# ----------------------------------------------------------------------------
# group
# ----------------------------------------------------------------------------
class Layer:
    def __init__(self, master_text, layername, params=None):
        self.m = parse_oso(master_text)   # private copy per instance
        self.name = layername
        self.idx = -1
        for k, v in (params or {}).items():
            s = self.m.byname[k]
            if not isinstance(v, (list, tuple)):
                v = [v]
            if s.t.base not in ("int", "string"):
                v = [struct.unpack("f", struct.pack("f", float(x)))[0] for x in v]
            if s.t.base in ("color", "point", "vector", "normal") and not s.t.arr and len(v) == 1:
                v = list(v) * 3         # a float value for a triple parameter (shadingsys.cpp:2880-3035)
            if getattr(s, "unsized", False):
                # an unsized array parameter takes the length of the instance value (instance.cpp:250-330)
                nc = {"matrix": 16, "color": 3, "point": 3, "vector": 3, "normal": 3}.get(s.t.base, 1)
                s.t = OType(s.t.base, max(1, len(v) // nc))
            s.vals = list(v)
            s.initexpr = False
            s.override = True
        self.lazy = False
        self.unused = False


class Group:
    """layers: list[Layer]; connections: (srclayer, srcparam, dstlayer, dstparam);
    outputs: list of dict(name='layer.param' or 'param', offset, stride, derivs)."""

    def __init__(self, layers, connections=(), outputs=(), name="group", attributes=None):
        self.layers = list(layers)
        self.name = name
        # uniform renderer attributes getattribute() can return: {name: int | float | str | list of those}
        self.attributes = dict(attributes or {})
        for i, l in enumerate(self.layers):
            l.idx = i
        byname = {l.name: l for l in self.layers}
        self.connections = []
        for (sl, sp, dl, dp) in connections:
            s, d = byname[sl], byname[dl]
            ssym, dsym = s.m.byname[sp], d.m.byname[dp]
            if getattr(dsym, "unsized", False) and ssym.t.arr:
                dsym.t = OType(dsym.t.base, ssym.t.arr)     # ... or the length of what is connected to it
            dsym.connected_from = (s.idx, ssym)
            ssym.connected_down = True
            self.connections.append((s.idx, ssym, d.idx, dsym))
        self.outputs = []
        for o in outputs:
            nm = o["name"]
            if "." in nm:
                ln, pn = nm.split(".", 1)
                lay = byname[ln]
            else:
                pn = nm
                lay = None
                for l in reversed(self.layers):   # last layer that has it wins
                    if pn in l.m.byname and l.m.byname[pn].symtype in ("param", "oparam"):
                        lay = l
                        break
                if lay is None and pn in self.layers[-1].m.byname and \
                        self.layers[-1].m.byname[pn].symtype == "global":
                    # a ShaderGlobals field as the entry layer left it (globals are in / out of execute():
                    # a displacement shader's P, simpleraytracer.cpp:1365-1384)
                    lay = self.layers[-1]
                if lay is None:
                    raise KeyError("renderer output %s not found" % nm)
            sym = lay.m.byname[pn]
            sym.renderer_output = dict(offset=o["offset"], stride=o["stride"],
                                       derivs=bool(o.get("derivs", False)))
            self.outputs.append((lay.idx, sym))
        self.analyze()

    # -- analysis -----------------------------------------------------------
    def analyze(self):
        n = len(self.layers)
        for l in self.layers:
            for op in l.m.ops:
                for a, c in zip(op.args, op.rw):
                    if c in "wW":
                        a.written = True
        # unused(): not last, no downstream connection, no renderer output
        # (llvm_instance.cpp:2306-2318)
        for l in self.layers:
            has_out = any(s.renderer_output for s in l.m.syms)
            has_down = any(s.connected_down for s in l.m.syms)
            l.unused = (l.idx != n - 1) and not has_out and not has_down
            # run_lazily (oslexec_pvt.h:1344-1372, defaults lazylayers=1,
            # lazyglobals=1, lazyunconnected=1, lazyerror=1)
            l.lazy = (l.idx != n - 1) and not has_out
        # derivative needs, last layer first so requirements flow upstream
        for l in reversed(self.layers):
            self.track_derivs(l)
            for s in l.m.syms:
                if s.connected_from and s.has_derivs:
                    s.connected_from[1].has_derivs = True

    DERIV_GLOBALS = ("P", "I", "u", "v", "Ps")

    def track_derivs(self, l):
        deps = {}
        need = set()
        for op in l.m.ops:
            reads = [a for a, c in zip(op.args, op.rw) if c in "rW"]
            writes = [a for a, c in zip(op.args, op.rw) if c in "wW"]
            for w in writes:
                for r in reads:
                    if not r.isconst:
                        deps.setdefault(w, set()).add(r)
            for ai in op.derivs:
                s = op.args[ai]
                if s.isconst or not s.t.floatbased or s.t.base == "matrix":
                    continue
                if s.symtype == "global" and s.name not in self.DERIV_GLOBALS:
                    continue
                need.add(s)
        for s in l.m.syms:
            if s.symtype == "global" and s.written and s.t.floatbased and s.name != "N":
                s.has_derivs = True
            ro = s.renderer_output
            if ro and ro["derivs"] and s.written and s.t.floatbased:
                s.has_derivs = True
            if s.has_derivs:
                need.add(s)
        seen = set()
        stack = list(need)
        while stack:
            s = stack.pop()
            if s in seen:
                continue
            seen.add(s)
            if s.t.floatbased and s.t.base != "matrix" and not s.isconst:
                s.has_derivs = True
            for r in deps.get(s, ()):
                stack.append(r)
        for s in l.m.syms:
            if s.symtype == "global" and s.name not in self.DERIV_GLOBALS:
                s.has_derivs = False
            if not s.t.floatbased or s.t.base == "matrix":
                s.has_derivs = False


# ----------------------------------------------------------------------------
# code generation
# ----------------------------------------------------------------------------
def cfloat(f):
    if f != f:
        return "NAN"
    if f in (float("inf"), float("-inf")):
        return "INFINITY" if f > 0 else "(-INFINITY)"
    s = "%.9g" % f
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s + "f"


def cstr(s):
    return '"' + s.replace("\\", "\\\\").replace('"', '\\"').replace("\n", "\\n") \
        .replace("\t", "\\t") + '"'


UNARY = {"sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "log2",
         "log10", "exp", "exp2", "expm1", "erf", "erfc", "cbrt", "sqrt", "inversesqrt",
         "abs", "fabs", "floor", "ceil", "round", "trunc", "sign", "logb", "neg"}
BINARY = {"add", "sub", "mul", "atan2", "pow", "fmod", "step", "min", "max"}
TERNARY = {"mix", "clamp", "smoothstep", "select"}
CMP = {"eq": "==", "neq": "!=", "lt": "<", "gt": ">", "le": "<=", "ge": ">="}
INTBIN = {"bitand": "&", "bitor": "|", "xor": "^", "shl": "<<", "shr": ">>"}
NOISE_KIND = {"noise": "N_NOISE", "uperlin": "N_NOISE", "snoise": "N_SNOISE",
              "perlin": "N_SNOISE", "cellnoise": "N_CELL", "cell": "N_CELL",
              "hashnoise": "N_HASH", "hash": "N_HASH", "simplex": "N_SIMPLEX",
              "simplexnoise": "N_SIMPLEX", "usimplex": "N_USIMPLEX",
              "usimplexnoise": "N_USIMPLEX", "gabor": "N_GABOR"}
PNOISE_KIND = {"pnoise": "N_NOISE", "psnoise": "N_SNOISE", "pcellnoise": "N_CELL", "gabor": "N_GABOR",
               "phashnoise": "N_HASH", "noise": "N_NOISE", "uperlin": "N_NOISE",
               "snoise": "N_SNOISE", "perlin": "N_SNOISE", "cell": "N_CELL",
               "cellnoise": "N_CELL", "hash": "N_HASH", "hashnoise": "N_HASH"}
SG_GLOBALS = {"P": "sg.P", "I": "sg.I", "N": "sg.N", "Ng": "sg.Ng", "u": "sg.u",
              "v": "sg.v", "dPdu": "sg.dPdu", "dPdv": "sg.dPdv", "Ps": "sg.Ps",
              "time": "sg.time", "dtime": "sg.dtime", "dPdtime": "sg.dPdtime", "Ci": "sg.Ci"}

SPLINE_TYPES = {"catmull-rom": 0, "catmullrom": 0, "bezier": 1, "bspline": 2, "hermite": 3,
                "linear": 4, "constant": 5}

# closure registry of the testrender renderer (src/testrender/shading.cpp:194-297):
# (name, number of positional params) -> closure id constant
CLOSURES = {("emission", 0): "EMISSION_ID", ("background", 0): "BACKGROUND_ID",
            ("diffuse", 1): "DIFFUSE_ID", ("oren_nayar", 2): "OREN_NAYAR_ID",
            ("translucent", 1): "TRANSLUCENT_ID", ("phong", 2): "PHONG_ID",
            ("ward", 4): "WARD_ID", ("microfacet", 7): "MICROFACET_ID",
            ("reflection", 1): "REFLECTION_ID", ("reflection", 2): "FRESNEL_REFLECTION_ID",
            ("refraction", 2): "REFRACTION_ID", ("transparent", 0): "TRANSPARENT_ID",
            ("transparent_bsdf", 0): "MX_TRANSPARENT_ID",
            # MaterialX closures registered from libbsdl lobes (BSDLtoOSL, shading.cpp:156-182)
            ("oren_nayar_diffuse_bsdf", 3): "MX_OREN_NAYAR_DIFFUSE_ID",
            ("burley_diffuse_bsdf", 3): "MX_BURLEY_DIFFUSE_ID",
            ("sheen_bsdf", 3): "MX_SHEEN_ID", ("layer", 2): "MX_LAYER_ID",
            ("uniform_edf", 1): "MX_UNIFORM_EDF_ID",
            ("conductor_bsdf", 7): "MX_CONDUCTOR_ID", ("dielectric_bsdf", 8): "MX_DIELECTRIC_ID",
            ("generalized_schlick_bsdf", 10): "MX_GENERALIZED_SCHLICK_ID",
            ("translucent_bsdf", 2): "MX_TRANSLUCENT_ID",
            ("subsurface_bssrdf", 4): "MX_SUBSURFACE_ID",
            # spi::ThinLayerLobe (SpiThinLayer, shading.cpp:119-152; Data: SPI/bsdf_thinlayer_decl.h:112-124)
            ("thinlayer", 9): "SPI_THINLAYER",
            # participating media (shading.cpp:265-284)
            ("anisotropic_vdf", 3): "MX_ANISOTROPIC_VDF_ID", ("medium_vdf", 6): "MX_MEDIUM_VDF_ID"}
# keyword parameters per closure id, in slot order after the positional words; value
# type "int" / "float".  Unspecified keywords are zero (llvm_gen_closure memsets the
# parameter block when there is no prepare callback).  String keywords ("label") have
# no effect on this path and take no slot.
CLOSURE_KEYS = {"MX_OREN_NAYAR_DIFFUSE_ID": [("energy_compensation", "int")],
                "MX_SHEEN_ID": [("mode", "int")],
                # libbsdl Data structs: bsdf_conductor_decl.h:43-46, bsdf_dielectric_decl.h:112-120
                "MX_CONDUCTOR_ID": [("thinfilm_thickness", "float"), ("thinfilm_ior", "float")],
                "MX_DIELECTRIC_ID": [("thinfilm_thickness", "float"), ("thinfilm_ior", "float"),
                                     ("absorption", "color"), ("dispersion", "float")]}
KEY_WORDS = {"int": 1, "float": 1, "color": 3}


class Gen:
    def __init__(self, group):
        self.g = group
        self.out = []
        self.ind = 1
        self.label = 0
        self.messages = {}

    def w(self, s):
        self.out.append("    " * self.ind + s)

    def ident(self, name):
        return re.sub(r"[^A-Za-z0-9_]", "_", name.replace("$", "S_"))

    def ctype(self, s):
        b = s.t.base
        if b == "int":
            return "int"
        if b == "float":
            return "Df" if s.has_derivs else "float"
        if s.t.triple:
            return "Dv" if s.has_derivs else "V3"
        if b == "string":
            return "const char*"
        if b == "matrix":
            return "M44"
        if b == "closure color":
            return "const Clos*"
        raise NotImplementedError("type %s" % s.t)

    def ref(self, l, s):
        """C++ lvalue/rvalue expression for symbol s inside layer l."""
        if s.symtype == "const":
            return self.constexpr(s)
        if s.symtype in ("param", "oparam"):
            return "gd.L%d_%s" % (l.idx, self.ident(s.name))
        if s.symtype == "global":
            if s.name not in SG_GLOBALS:
                raise NotImplementedError("global %s" % s.name)
            e = SG_GLOBALS[s.name]
            # globals that carry derivs in SG but are used without them
            if s.name in Group.DERIV_GLOBALS and not s.has_derivs:
                return "%s.val" % e
            return e
        return self.ident(s.name)

    def constexpr(self, s):
        b = s.t.base
        if s.t.arr:
            return "K_" + self.ident(s.name)
        if b == "int":
            v = int(s.vals[0]) & 0xffffffff          # OSL ints are 32 bit: 0xffffffff is -1
            v = v - (1 << 32) if v & 0x80000000 else v
            return "(-2147483647 - 1)" if v == -(1 << 31) else str(v)
        if b == "float":
            return cfloat(s.vals[0])
        if s.t.triple:
            return "V3(%s, %s, %s)" % tuple(cfloat(v) for v in s.vals[:3])
        if b == "string":
            return cstr(s.vals[0])
        if b == "matrix":
            return "M44(%s)" % ", ".join(cfloat(v) for v in (list(s.vals) + [0.0] * 16)[:16])
        raise NotImplementedError(b)

    def initval(self, s):
        """C++ initializer for a param/local from its default values."""
        b = s.t.base
        n = max(1, s.t.arr)
        per = s.t.ncomp
        vals = list(s.vals) + [0] * (n * per - len(s.vals)) if b != "string" else \
            list(s.vals) + [""] * (n - len(s.vals))

        def one(v):
            if b == "int":
                return str(int(v[0]))
            if b == "float":
                return ("Df(%s)" if s.has_derivs else "%s") % cfloat(v[0])
            if s.t.triple:
                e = "V3(%s, %s, %s)" % tuple(cfloat(x) for x in v)
                return "Dv(%s)" % e if s.has_derivs else e
            if b == "string":
                return cstr(v[0])
            if b == "matrix":
                return "M44(%s)" % ", ".join(cfloat(x) for x in v)
            if b == "closure color":
                return "nullptr"
            raise NotImplementedError(b)
        return [one(vals[i * per:(i + 1) * per]) for i in range(n)]

    # -- group --------------------------------------------------------------
    def generate_material(self, ns):
        """Group as a callable `void <ns>::entry(SG&)` (the renderer executes
        one point at a time: simpleraytracer.cpp:1034)."""
        g = self.g
        o = self.out
        self.scan_messages()
        o.append("namespace %s {" % ns)
        o.append("struct GD {")
        o.append("    bool ran[%d];" % max(1, len(g.layers)))
        o.extend(self.message_fields())
        for l in g.layers:
            if l.unused:
                continue
            for s in l.m.syms:
                if s.symtype in ("param", "oparam"):
                    arr = "[%d]" % s.t.arr if s.t.arr else ""
                    o.append("    %s L%d_%s%s;" % (self.ctype(s), l.idx, self.ident(s.name), arr))
        o.append("};")
        used = [l for l in g.layers if not l.unused]
        for l in used:
            o.append("static void layer_%d(SG& sg, GD& gd, const Launch* L);" % l.idx)
        for l in used:
            self.gen_layer(l)
        o.append("static void entry(SG& sg)")
        o.append("{")
        o.append("    GD gd;")
        o.append("    for (int k = 0; k < %d; ++k) gd.ran[k] = false;" % max(1, len(g.layers)))
        o.append("    layer_%d(sg, gd, nullptr);" % (len(g.layers) - 1))
        o.append("}")
        o.append("}  // namespace %s" % ns)
        return "\n".join(o) + "\n"

    def generate(self):
        g = self.g
        o = self.out
        o.append("// generated by oracle/oso2cpp.py — CPU oracle, test infrastructure only")
        o.append('#include "osl_oracle_closure.h"')
        o.append('#include "osl_oracle_runtime.h"')
        o.append("using namespace oslo;")
        o.append("namespace {")
        self.scan_messages()
        o.append("struct GD {")
        o.append("    bool ran[%d];" % max(1, len(g.layers)))
        o.extend(self.message_fields())
        for l in g.layers:
            if l.unused:
                continue
            for s in l.m.syms:
                if s.symtype in ("param", "oparam"):
                    arr = "[%d]" % s.t.arr if s.t.arr else ""
                    o.append("    %s L%d_%s%s;" % (self.ctype(s), l.idx, self.ident(s.name), arr))
        o.append("};")
        used = [l for l in g.layers if not l.unused]
        for l in used:
            o.append("static void layer_%d(SG& sg, GD& gd, const Launch* L);" % l.idx)
        for l in used:
            self.gen_layer(l)
        o.append("}  // namespace")
        o.append('extern "C" void oracle_run(const Launch* L, long long begin, long long end, '
                 'std::string* pf)')
        o.append("{")
        o.append("    Ctx ctx{pf};")
        o.append("    for (long long i = begin; i < end; ++i) {")
        o.append("        SG sg; GD gd;")
        o.append("        ClosurePool pool_; pool_.reset(); sg.pool = &pool_;   // per-point closure arena")
        o.append("        load_sg(sg, L, i, &ctx);")
        o.append("        for (int k = 0; k < %d; ++k) gd.ran[k] = false;" % max(1, len(g.layers)))
        o.append("        layer_%d(sg, gd, L);" % (len(g.layers) - 1))
        o.append("    }")
        o.append("}")
        o.append(RUNNER_TAIL)
        return "\n".join(o) + "\n"

    def gen_layer(self, l):
        g = self.g
        m = l.m
        self.l = l
        self.ind = 0
        self.w("static void layer_%d(SG& sg, GD& gd, const Launch* L)" % l.idx)
        self.w("{")
        self.ind = 1
        self.w("gd.ran[%d] = true;" % l.idx)
        self.w("(void)L;")
        # declare locals / temps
        for s in m.syms:
            if s.symtype in ("local", "temp"):
                arr = "[%d]" % s.t.arr if s.t.arr else ""
                init = " = {}" if s.t.arr else (" = nullptr" if s.t.base in ("string", "closure color")
                                                else " = {}")
                self.w("%s %s%s%s;" % (self.ctype(s), self.ident(s.name), arr, init))
            elif s.symtype == "const" and s.t.arr:
                vals = self.initval(s)
                self.w("static const %s K_%s[%d] = {%s};" % (
                    self.ctype(s), self.ident(s.name), s.t.arr, ", ".join(vals)))
        # entry layer: run earlier non-lazy layers unconditionally
        # (llvm_instance.cpp:1693-1720)
        if l.idx == len(g.layers) - 1:
            for e in g.layers[:-1]:
                if not e.unused and not e.lazy:
                    self.w("if (!gd.ran[%d]) layer_%d(sg, gd, L);" % (e.idx, e.idx))
        # param initialisation (llvm_instance.cpp:1666-1690, 703-1000)
        for s in m.syms:
            if s.symtype not in ("param", "oparam"):
                continue
            if s.connected_from is not None:
                continue   # value arrives from upstream (possibly lazily)
            vals = self.initval(s)
            r = self.ref(l, s)
            bound = getattr(s, "interpolated", False) and not s.t.arr
            if bound:
                # interpolated parameter: the renderer's userdata wins, else default / init ops
                self.w('if (!bind_userdata(L, sg, "%s", %s)) {' % (s.name, r))
                self.ind += 1
            if s.t.arr:
                for i, v in enumerate(vals):
                    self.w("%s[%d] = %s;" % (r, i, v))
            else:
                self.w("%s = %s;" % (r, vals[0]))
            if s.initexpr and s.name in m.methods:
                b, e = m.methods[s.name]
                self.ensured = set()
                self.emit_block(b, e, None)
            if bound:
                self.ind -= 1
                self.w("}")
        self.ensured = set()
        b, e = m.methods.get("___main___", (0, 0))
        self.emit_block(b, e, None)
        self.w("layer_end:;")
        # copy outputs to downstream params (llvm_instance.cpp:1738-1802)
        for (si, ssym, di, dsym) in g.connections:
            if si != l.idx or g.layers[di].unused:
                continue
            self.emit_copy(g.layers[di], dsym, l, ssym)
        # renderer outputs (llvm_instance.cpp:1807-1848)
        for (li, s) in g.outputs:
            if li != l.idx:
                continue
            ro = s.renderer_output
            fn = "wrd" if ro["derivs"] else "wr"
            self.w("%s(outp(L, sg, %d, %d), %s);" % (fn, ro["offset"], ro["stride"], self.ref(l, s)))
        self.ind = 0
        self.w("}")

    def emit_copy(self, dl, dsym, sl, ssym):
        d, s = self.ref(dl, dsym), self.ref(sl, ssym)
        if dsym.t.arr:
            for i in range(dsym.t.arr):
                self.w("assign(%s[%d], %s[%d]);" % (d, i, s, i))
        elif dsym.t.base in ("matrix", "closure color"):
            self.w("%s = %s;" % (d, s))
        else:
            self.w("assign(%s, %s);" % (d, s))

    # -- control flow -------------------------------------------------------
    def emit_block(self, b, e, ctx):
        """ctx = dict(ret=label or None, brk=label, cont=label)"""
        ops = self.l.m.ops
        i = b
        while i < e:
            op = ops[i]
            n = op.name
            if n == "if":
                self.useparams(op)
                self.w("if (%s) {" % self.ref(self.l, op.args[0]))
                self.ind += 1
                saved = set(self.ensured)
                self.emit_block(i + 1, op.jumps[0], ctx)
                self.ind -= 1
                self.ensured = set(saved)
                if op.jumps[1] > op.jumps[0]:
                    self.w("} else {")
                    self.ind += 1
                    self.emit_block(op.jumps[0], op.jumps[1], ctx)
                    self.ind -= 1
                    self.ensured = set(saved)
                self.w("}")
                i = op.jumps[1]
            elif n in ("for", "while", "dowhile"):
                cl, bl, il, dl = op.jumps
                self.label += 1
                lab = self.label
                c2 = dict(ctx or {})
                c2.update(brk="brk_%d" % lab, cont="cont_%d" % lab)
                self.emit_block(i + 1, cl, ctx)
                saved = set(self.ensured)
                cond = self.ref(self.l, op.args[0])
                self.w("for (;;) {")
                self.ind += 1
                if n == "dowhile":
                    self.emit_block(bl, il, c2)
                    self.w("cont_%d:;" % lab)
                    self.emit_block(cl, bl, c2)
                    self.w("if (!(%s)) break;" % cond)
                    self.emit_block(il, dl, c2)
                else:
                    self.emit_block(cl, bl, c2)
                    self.w("if (!(%s)) break;" % cond)
                    self.emit_block(bl, il, c2)
                    self.w("cont_%d:;" % lab)
                    self.emit_block(il, dl, c2)
                self.ind -= 1
                self.w("}")
                self.w("brk_%d:;" % lab)
                self.ensured = set(saved)
                i = dl
            elif n == "functioncall":
                self.label += 1
                lab = self.label
                c2 = dict(ctx or {})
                c2["ret"] = "ret_%d" % lab
                self.w("{")
                self.ind += 1
                saved = set(self.ensured)
                self.emit_block(i + 1, op.jumps[0], c2)
                self.ensured = set(saved)
                self.ind -= 1
                self.w("}")
                self.w("ret_%d:;" % lab)
                i = op.jumps[0]
            elif n == "break":
                self.w("goto %s;" % ctx["brk"])
                i += 1
            elif n == "continue":
                self.w("goto %s;" % ctx["cont"])
                i += 1
            elif n == "return":
                self.w("goto %s;" % ((ctx or {}).get("ret") or "layer_end"))
                i += 1
            elif n == "exit":
                self.w("goto layer_end;")
                i += 1
            elif n in ("nop", "end", "useparam"):
                i += 1
            else:
                self.useparams(op)
                self.w("{")
                self.ind += 1
                self.emit_op(op)
                self.ind -= 1
                self.w("}")
                i += 1

    def useparams(self, op):
        """Lazy upstream evaluation before reading a connected param
        (llvm_gen.cpp:133-283)."""
        for a, c in zip(op.args, op.rw):
            if c in "rW" and a.connected_from is not None:
                up = a.connected_from[0]
                if up in self.ensured:
                    continue
                self.ensured.add(up)
                self.w("if (!gd.ran[%d]) layer_%d(sg, gd, L);" % (up, up))

    # -- ops ----------------------------------------------------------------
    def R(self, s):
        return self.ref(self.l, s)

    def comp(self, s, c, derivs):
        """expression for component c of s as float (derivs=False) or as the
        natural scalar (float/Df)."""
        t = s.t
        if s.isconst:
            if t.base == "int":
                return "%s" % cfloat(float(s.vals[0]))
            if t.base == "float":
                return cfloat(s.vals[0])
            if t.triple:
                return cfloat(s.vals[c])
        e = "getc(%s, %d)" % (self.R(s), c)
        if not derivs and s.has_derivs:
            e = "nd(%s)" % e
        return e

    def emit_op(self, op):
        n = op.name
        A = op.args
        h = getattr(self, "op_" + n, None)
        if h is not None:
            return h(op)
        if n in UNARY or n in BINARY or n in TERNARY or n in ("div", "log"):
            return self.op_percomp(op)
        if n in CMP:
            return self.op_cmp(op)
        if n in INTBIN:
            return self.w("%s = %s %s %s;" % (self.R(A[0]), self.R(A[1]), INTBIN[n], self.R(A[2])))
        if n in TRIPLES:
            return self.op_triple(op)
        if n in NOISE_KIND and n != "hash":
            return self.noise_impl(op, False)
        if n in ("pnoise", "psnoise", "pcellnoise", "phashnoise"):
            return self.noise_impl(op, True)
        raise NotImplementedError("oracle: op '%s' not supported" % n)

    def op_percomp(self, op):
        n = op.name
        A = op.args
        d = A[0]
        if d.t.base == "closure color":
            return self.op_closure_arith(op)
        if d.t.base == "matrix":
            return self.op_matrix_arith(op)
        isint = d.t.base == "int"
        dv = d.has_derivs and any(a.has_derivs for a in A[1:])
        fn = "o_" + n
        if n == "log" and len(A) == 3:
            raise NotImplementedError("log(x,b) is an OSL-source function")
        if n == "div" and A[2].isconst and not any(v == 0 for v in A[2].vals):
            fn = "o_divc"
        for c in range(d.t.ncomp):
            if isint:
                args = [self.R(a) for a in A[1:]]
            else:
                args = [self.comp(a, c, dv) for a in A[1:]]
            self.w("setc(%s, %d, %s(%s));" % (self.R(d), c, fn, ", ".join(args)))

    def op_mod(self, op):
        A = op.args
        self.w("%s = o_mod(%s, %s);" % (self.R(A[0]), self.R(A[1]), self.R(A[2])))

    def op_compl(self, op):
        self.w("%s = ~%s;" % (self.R(op.args[0]), self.R(op.args[1])))

    def op_cmp(self, op):
        A = op.args
        a, b = A[1], A[2]
        cop = CMP[op.name]
        if a.t.base == "string":
            e = "str_eq(%s, %s)" % (self.R(a), self.R(b))
            if op.name == "neq":
                e = "!" + e
        elif a.t.base == "matrix" or b.t.base == "matrix":
            am = self.R(a) if a.t.base == "matrix" else "m44_diag(%s)" % self.fl(a)
            bm = self.R(b) if b.t.base == "matrix" else "m44_diag(%s)" % self.fl(b)
            e = "%s(%s == %s)" % ("!" if op.name == "neq" else "", am, bm)
        elif a.t.base == "closure color":
            e = "(%s %s nullptr)" % (self.R(a), cop)
        elif a.t.base == "int" and b.t.base == "int":
            e = "(%s %s %s)" % (self.R(a), cop, self.R(b))
        else:
            nc = max(a.t.ncomp, b.t.ncomp)
            parts = ["(%s %s %s)" % (self.comp(a, c if a.t.ncomp > 1 else 0, False), cop,
                                     self.comp(b, c if b.t.ncomp > 1 else 0, False))
                     for c in range(nc)]
            e = (" || " if op.name == "neq" else " && ").join(parts)
        self.w("%s = (%s) ? 1 : 0;" % (self.R(A[0]), e))

    def op_assign(self, op):
        d, s = op.args
        if d.t.arr:
            # array = array copies as many elements as both have (llvm_assign_impl, llvm_gen.cpp:770-800;
            # testsuite/array-copy assigns an int[2] to an int[3])
            for i in range(min(d.t.arr, s.t.arr or d.t.arr)):
                self.w("assign(%s[%d], %s[%d]);" % (self.R(d), i, self.R(s), i))
        elif d.t.base == "matrix":
            if s.t.base == "matrix":
                self.w("%s = %s;" % (self.R(d), self.R(s)))
            else:
                self.w("%s = m44_diag(%s);" % (self.R(d), self.comp(s, 0, False)))
        elif d.t.base == "closure color":
            self.w("%s = %s;" % (self.R(d), "nullptr" if s.t.base == "int" else self.R(s)))
        else:
            self.w("assign(%s, %s);" % (self.R(d), self.R(s)))

    def op_triple(self, op):
        A = op.args
        d = A[0]
        if len(A) == 5 and op.name != "color":
            # llvm_gen_construct_triple (llvm_gen.cpp:1871-1933): build the triple (with the
            # derivatives of the components), then osl_transform_triple(space -> "common")
            dv = d.has_derivs and any(a.has_derivs for a in A[2:])
            for c in range(3):
                self.w("setc(%s, %d, %s);" % (self.R(d), c, self.comp(A[2 + c], 0, dv)))
            sp = A[1]
            if sp.constval and sp.vals[0] in ("common", "world"):
                return
            vt = {"point": 0, "vector": 1, "normal": 2}[op.name]
            T = "Dv" if d.has_derivs else "V3"
            self.w("{ %s i_ = %s, o_; xf_transform_triple_err(sg, xf_set(L), %s, \"common\", i_, o_, %d); %s = o_; }" % (
                T, self.R(d), self.R(sp), vt, self.R(d)))
            return
        if len(A) == 5:
            if op.name != "color":
                raise NotImplementedError("triple constructor with a space name")
            # llvm_gen_construct_color (llvm_gen.cpp:1826-1865): osl_prepend_color_from on
            # the value, derivatives of the result are zeroed
            self.w("{ V3 c_(%s, %s, %s); assign(%s, color_to_rgb(colorsystem(), %s, c_)); }" % (
                self.comp(A[2], 0, False), self.comp(A[3], 0, False), self.comp(A[4], 0, False),
                self.R(d), self.R(A[1])))
            return
        dv = d.has_derivs and any(a.has_derivs for a in A[1:])
        for c in range(3):
            self.w("setc(%s, %d, %s);" % (self.R(d), c, self.comp(A[1 + c], 0, dv)))

    def op_compref(self, op):
        d, s, i = op.args
        if i.isconst:
            self.w("setc(%s, 0, %s);" % (self.R(d), self.comp(s, int(i.vals[0]), d.has_derivs)))
        else:
            # out of range: osl_range_check_err clamps (>= 3 -> 2, < 0 -> 0)
            self.w("switch (%s < 0 ? 0 : %s) {" % (self.R(i), self.R(i)))
            for c in range(3):
                self.w("%s %d: setc(%s, 0, %s); break;" % (
                    "default: case" if c == 2 else "case", c, self.R(d),
                    self.comp(s, c, d.has_derivs)))
            self.w("}")

    def op_compassign(self, op):
        d, i, s = op.args
        v = self.comp(s, 0, d.has_derivs)
        if i.isconst:
            self.w("setc(%s, %d, %s);" % (self.R(d), int(i.vals[0]), v))
        else:
            self.w("switch (%s < 0 ? 0 : %s) {" % (self.R(i), self.R(i)))
            for c in range(3):
                self.w("%s %d: setc(%s, %d, %s); break;" % (
                    "default: case" if c == 2 else "case", c, self.R(d), c, v))
            self.w("}")

    def op_aref(self, op):
        d, s, i = op.args
        # osl_range_check_err (shadingsys.cpp:4981-4984): >= length -> length - 1, negative -> 0
        self.w("{ int ix_ = %s; ix_ = ix_ < 0 ? 0 : (ix_ >= %d ? %d - 1 : ix_); assign(%s, %s[ix_]); }" % (
            self.R(i), s.t.arr, s.t.arr, self.R(d), self.R(s)))

    def op_aassign(self, op):
        d, i, s = op.args
        self.w("{ int ix_ = %s; ix_ = ix_ < 0 ? 0 : (ix_ >= %d ? %d - 1 : ix_); assign(%s[ix_], %s); }" % (
            self.R(i), d.t.arr, d.t.arr, self.R(d), self.R(s)))

    def op_arraylength(self, op):
        self.w("%s = %d;" % (self.R(op.args[0]), op.args[1].t.arr))

    def op_sincos(self, op):
        x, s, c = op.args
        for k in range(x.t.ncomp):
            self.w("{ auto x_ = %s; setc(%s, %d, o_sin(x_)); setc(%s, %d, o_cos(x_)); }" % (
                self.comp(x, k, (s.has_derivs or c.has_derivs) and x.has_derivs),
                self.R(s), k, self.R(c), k))

    def op_vec(self, op):
        # dot cross length distance normalize: pass whole values; the C++
        # overloads pick the Dv form only when an operand carries derivs
        A = op.args
        d = A[0]
        dv = d.has_derivs and any(a.has_derivs for a in A[1:])
        args = []
        for a in A[1:]:
            e = self.R(a)
            if a.has_derivs and not dv:
                e = "nd(%s)" % e
            args.append(e)
        self.w("assign(%s, o_%s(%s));" % (self.R(d), op.name, ", ".join(args)))

    op_dot = op_cross = op_length = op_distance = op_normalize = op_luminance = op_vec

    def op_deriv(self, op):
        d, s = op.args
        self.w("assign(%s, o_%s(%s));" % (self.R(d), op.name, self.R(s)))

    op_Dx = op_Dy = op_filterwidth = op_deriv

    def op_Dz(self, op):
        d, s = op.args
        if s.symtype == "global" and s.name == "P":
            self.w("assign(%s, sg.dPdz);" % self.R(d))
        else:
            self.w("assign(%s, 0.0f);" % self.R(d))

    def op_area(self, op):
        self.w("assign(%s, o_area(%s));" % (self.R(op.args[0]), self.R(op.args[1])))

    def op_calculatenormal(self, op):
        self.w("assign(%s, o_calculatenormal(%s, sg.flipHandedness != 0));" % (
            self.R(op.args[0]), self.R(op.args[1])))

    def op_isnan(self, op):
        self.w("%s = std::isnan(%s) ? 1 : 0;" % (self.R(op.args[0]), self.comp(op.args[1], 0, False)))

    def op_isinf(self, op):
        self.w("%s = std::isinf(%s) ? 1 : 0;" % (self.R(op.args[0]), self.comp(op.args[1], 0, False)))

    def op_isfinite(self, op):
        self.w("%s = std::isfinite(%s) ? 1 : 0;" % (self.R(op.args[0]), self.comp(op.args[1], 0, False)))

    def op_surfacearea(self, op):
        self.w("assign(%s, sg.surfacearea);" % self.R(op.args[0]))

    def op_backfacing(self, op):
        self.w("%s = sg.backfacing;" % self.R(op.args[0]))

    def op_raytype(self, op):
        d, nm = op.args
        if not nm.constval:
            raise NotImplementedError("raytype(non-constant)")
        self.w("%s = (sg.raytype & %d) != 0;" % (self.R(d), raytype_bit(nm.vals[0])))

    def op_isconnected(self, op):
        d, s = op.args
        # (up ? 1 : 0) + ((down || renderer output) ? 2 : 0)  (runtimeoptimize.cpp:2495-2499)
        v = (1 if s.connected_from is not None else 0) + (2 if (s.connected_down or s.renderer_output) else 0)
        self.w("%s = %d;" % (self.R(d), v))

    FOLDABLE = {"assign", "add", "sub", "mul", "div", "mod", "neg", "abs", "fabs", "sqrt", "inversesqrt", "pow",
                "min", "max", "floor", "ceil", "round", "trunc", "sign", "clamp", "mix", "color", "point", "vector",
                "normal", "float", "int", "compref", "dot", "cross", "length", "normalize", "sin", "cos", "tan",
                "exp", "exp2", "log", "log2", "eq", "neq", "lt", "gt", "le", "ge", "and", "or", "not", "bitand",
                "bitor", "xor", "shl", "shr", "compl", "step", "smoothstep", "select"}

    def folds_to_constant(self, s, depth=0):
        """What the reference's constant folder (constfold.cpp) knows at optimisation time: constants,
        instance values, and top-level temporaries computed from those by foldable ops."""
        if s.constval:
            return True
        if depth > 16 or s.symtype not in ("temp", "local") or s.t.arr:
            return False
        ops = self.l.m.ops
        writers = [i for i, op in enumerate(ops) for a, c in zip(op.args, op.rw) if a is s and c in "wW"]
        if len(writers) != 1:
            return False
        w = writers[0]
        if ops[w].name not in self.FOLDABLE or ops[w].jumps:
            return False
        if any(j < w < max(op.jumps) for j, op in enumerate(ops) if op.jumps):
            return False
        return all(self.folds_to_constant(a, depth + 1) for a, c in zip(ops[w].args, ops[w].rw) if c == "r")

    def op_isconstant(self, op):
        self.w("%s = %d;" % (self.R(op.args[0]), 1 if self.folds_to_constant(op.args[1]) else 0))

    def op_hash(self, op):
        A = op.args
        d = A[0]
        ins = A[1:]
        if len(ins) == 1 and ins[0].t.base == "int":
            e = "hash_i(%s)" % self.R(ins[0])
        elif len(ins) == 1 and ins[0].t.base == "float":
            e = "hash_f(%s)" % self.comp(ins[0], 0, False)
        elif len(ins) == 2 and ins[0].t.base == "float":
            e = "hash_ff(%s, %s)" % (self.comp(ins[0], 0, False), self.comp(ins[1], 0, False))
        elif len(ins) == 1:
            e = "hash_v(nd(%s))" % self.R(ins[0])
        else:
            e = "hash_vf(nd(%s), %s)" % (self.R(ins[0]), self.comp(ins[1], 0, False))
        self.w("%s = %s;" % (self.R(d), e))

    def noise_impl(self, op, periodic, runtime_name=None):
        """llvm_gen_noise (llvm_gen.cpp:3117-3299): resolve the name at gen
        time, pick float/Dual form from has_derivs of result and inputs.  A name that is
        not a compile-time constant goes through osl_genericnoise / osl_genericpnoise in the
        reference (GenericNoise / GenericPNoise, opnoise.cpp:704-900: string compares at run
        time); here: one branch per accepted name."""
        A = list(op.args)
        d = A[0]
        rest = A[1:]
        name = op.name
        if rest and rest[0].t.base == "string":
            if not rest[0].constval and runtime_name is None:
                names = [("perlin", "snoise"), ("uperlin", "noise"), ("simplex", "simplexnoise"),
                         ("usimplex", "usimplexnoise"), ("cell",), ("hash",), ("gabor",)]
                if periodic:
                    names = [n for n in names if "simplex" not in n[0]]
                self.w("const char* nm_ = %s;" % self.R(rest[0]))
                for k, alts in enumerate(names):
                    cond = " || ".join('str_eq(nm_, "%s")' % a for a in alts)
                    self.w("%sif (%s) {" % ("} else " if k else "", cond))
                    self.ind += 1
                    self.noise_impl(op, periodic, runtime_name=alts[0])
                    self.ind -= 1
                self.w("}")   # unknown name: "Unknown noise type" error in the reference, result untouched
                return
            name = runtime_name if runtime_name is not None else rest[0].vals[0]
            rest = rest[1:]
        # strip optional token/value pairs
        coords = []
        for a in rest:
            if a.t.base == "string":
                break
            coords.append(a)
        opts = rest[len(coords):]
        table = PNOISE_KIND if periodic else NOISE_KIND
        if name not in table:
            raise NotImplementedError("noise type '%s'" % name)
        kind = table[name]
        if periodic:
            half = len(coords) // 2
            pers = coords[half:]
            coords = coords[:half]
        if kind == "N_GABOR":
            return self.gabor_impl(d, coords, pers if periodic else None, opts)
        # flatten coordinates
        ins = []
        for a in coords:
            for c in range(a.t.ncomp):
                ins.append((a, c))
        dim = len(ins)
        nc = d.t.ncomp
        hashy = kind in ("N_CELL", "N_HASH")
        dv = (not hashy) and d.has_derivs and any(a.has_derivs for a in coords)
        S = "Df" if dv else "float"
        self.w("%s in_[4] = {%s};" % (S, ", ".join(self.comp(a, c, dv) for a, c in ins)))
        self.w("%s out_[3];" % S)
        if periodic:
            pin = []
            for a in pers:
                for c in range(a.t.ncomp):
                    pin.append(self.comp(a, c, False))
            if hashy:
                self.w("float per_[4] = {%s};" % ", ".join(pin))
                self.w("for (int k_ = 0; k_ < %d; ++k_) in_[k_] = pwrap(in_[k_], per_[k_]);" % dim)
                self.w("ihnoise_core<%s, %d>(out_, %d, in_);" % (kind, nc, dim))
            else:
                self.w("int per_[4] = {%s};" % ", ".join("iperiod(%s)" % p for p in pin))
                self.w("perlin_nd<%s, %d, %s>(out_, %d, in_, per_);" % (
                    S, nc, "true" if kind == "N_SNOISE" else "false", dim))
        elif hashy:
            self.w("ihnoise_core<%s, %d>(out_, %d, in_);" % (kind, nc, dim))
        else:
            self.w("noise_core<%s, %s, %d>(out_, %d, in_);" % (kind, S, nc, dim))
        for c in range(nc):
            self.w("setc(%s, %d, out_[%d]);" % (self.R(d), c, c))


    def gabor_impl(self, d, coords, pers, opts):
        """Gabor branch of llvm_gen_noise (llvm_gen.cpp:3196-3203): derivs are
        always taken (zero when the coordinate has none), options go through
        NoiseParams (llvm_gen_noise_options, llvm_gen.cpp:3057-3103); 1-D/2-D
        slice the 3-D noise and 4-D ignores time (opnoise.cpp:484-632)."""
        ins = [(a, c) for a in coords for c in range(a.t.ncomp)]
        comps = ["Df(%s)" % self.comp(a, c, True) for a, c in ins][:3]
        while len(comps) < 3:
            comps.append("Df(0.0f)")
        self.w("NoiseParams opt_;")
        i = 0
        while i + 1 < len(opts):
            nm, val = opts[i], opts[i + 1]
            i += 2
            if not nm.constval:
                raise NotImplementedError("noise option with a non-constant name")
            key = nm.vals[0]
            e = self.R(val)
            if val.has_derivs:
                e = "nd(%s)" % e
            if key == "":
                continue
            if key in ("anisotropic", "do_filter") and val.t.base == "int" and not val.t.arr:
                self.w("opt_.%s = %s;" % (key, e))
            elif key == "direction" and val.t.triple:
                self.w("assign(opt_.direction, %s);" % e)
            elif key in ("bandwidth", "impulses") and val.t.base in ("float", "int") and not val.t.triple \
                    and not val.t.arr:
                self.w("opt_.%s = (float)(%s);" % (key, e))
            else:
                # the reference reports "Unknown noise optional argument" and carries on
                self.w("// unknown noise option '%s' ignored" % key)
        self.w("Dv P_ = make_dv(%s);" % ", ".join(comps))
        per = "nullptr"
        if pers is not None:
            pc = [self.comp(a, c, False) for a in pers for c in range(a.t.ncomp)][:3]
            while len(pc) < 3:
                pc.append("0.0f")
            self.w("V3 per_(%s);" % ", ".join(pc))
            per = "&per_"
        nc = d.t.ncomp
        self.w("Df out_[3]; gabor_noise<%d>(out_, P_, %s, opt_);" % (nc, per))
        for c in range(nc):
            self.w("setc(%s, %d, out_[%d]%s);" % (self.R(d), c, c, "" if d.has_derivs else ".val"))

    # ---- colour shadeops (opcolor.cpp:434-518) -------------------------------------
    def op_luminance(self, op):
        d, c = op.args
        dv = d.has_derivs and c.has_derivs
        e = self.R(c) if (dv or not c.has_derivs) else "nd(%s)" % self.R(c)
        if c.isconst:
            e = "V3(%s)" % ", ".join(self.comp(c, k, False) for k in range(3))
        self.w("assign(%s, color_luminance(colorsystem(), %s));" % (self.R(d), e))

    def op_blackbody(self, op):
        d, t = op.args
        self.w("assign(%s, color_blackbody(colorsystem(), %s));" % (self.R(d), self.comp(t, 0, False)))

    def op_wavelength_color(self, op):
        d, t = op.args
        self.w("assign(%s, color_wavelength(colorsystem(), %s));" % (self.R(d), self.comp(t, 0, False)))

    def op_transformc(self, op):
        """osl_transformc (opcolor.cpp:487-518): Dual form only when both sides carry derivs."""
        d, frm, to, c = op.args
        dv = d.has_derivs and c.has_derivs
        e = self.R(c) if (dv or not c.has_derivs) else "nd(%s)" % self.R(c)
        if c.isconst:
            e = "V3(%s)" % ", ".join(self.comp(c, k, False) for k in range(3))
        self.w("assign(%s, color_transformc(colorsystem(), %s, %s, %s));" % (
            self.R(d), self.R(frm), self.R(to), e))

    # ---- matrix shadeops (opmatrix.cpp; llvm_gen_matrix / _getmatrix / _transform) ----
    def fl(self, s):
        """scalar float expression of an int/float symbol (no derivs)"""
        return self.comp(s, 0, False)

    def op_matrix_arith(self, op):
        n, A = op.name, op.args
        d = self.R(A[0])
        if n == "neg":
            return self.w("%s = -%s;" % (d, self.R(A[1])))
        a, b = A[1], A[2]
        am, bm = a.t.base == "matrix", b.t.base == "matrix"
        if n == "mul":
            if am and bm:
                return self.w("%s = %s * %s;" % (d, self.R(a), self.R(b)))
            m, f = (a, b) if am else (b, a)
            return self.w("%s = %s * %s;" % (d, self.R(m), self.fl(f)))
        if n == "div":
            if am and bm:       # osl_div_mmm
                return self.w("%s = %s * m44_inverse(%s);" % (d, self.R(a), self.R(b)))
            if am:              # osl_div_mmf
                return self.w("%s = %s * (1.0f / %s);" % (d, self.R(a), self.fl(b)))
            if bm:              # osl_div_mfm
                return self.w("%s = %s * m44_inverse(%s);" % (d, self.fl(a), self.R(b)))
            # osl_div_m_ff
            return self.w("{ float b_ = %s; %s = m44_diag(b_ == 0 ? 0.0f : (%s / b_)); }" % (self.fl(b), d, self.fl(a)))
        raise NotImplementedError("matrix %s" % n)

    def op_matrix(self, op):
        """llvm_gen_matrix (llvm_gen.cpp:2295-2370)"""
        A = op.args
        d = self.R(A[0])
        nargs = len(A)
        using_space = nargs in (3, 18) and A[1].t.base == "string"
        two_spaces = nargs == 3 and A[2].t.base == "string"
        if two_spaces:
            return self.w("xf_get_from_to_matrix_err(sg, xf_set(L), %s, %s, %s);" % (self.R(A[1]), self.R(A[2]), d))
        vals = A[1 + using_space:]
        if len(vals) == 1:
            self.w("%s = m44_diag(%s);" % (d, self.fl(vals[0])))
        elif len(vals) == 16:
            self.w("%s = M44(%s);" % (d, ", ".join(self.fl(v) for v in vals)))
        else:
            raise NotImplementedError("matrix constructor with %d values" % len(vals))
        if using_space:   # osl_prepend_matrix_from
            self.w("{ M44 f_; if (xf_get_matrix_err(sg, xf_set(L), %s, f_)) %s = f_ * %s; }" % (self.R(A[1]), d, d))

    def op_getmatrix(self, op):
        r, frm, to, m = op.args
        self.w("%s = xf_get_from_to_matrix_err(sg, xf_set(L), %s, %s, %s) ? 1 : 0;" % (
            self.R(r), self.R(frm), self.R(to), self.R(m)))

    def op_mxcompref(self, op):
        d, m, i, j = op.args
        self.w("assign(%s, %s.x[%s][%s]);" % (self.R(d), self.R(m), self.R(i), self.R(j)))

    def op_mxcompassign(self, op):
        m, i, j, v = op.args
        self.w("%s.x[%s][%s] = %s;" % (self.R(m), self.R(i), self.R(j), self.fl(v)))

    def op_transpose(self, op):
        self.w("%s = m44_transposed(%s);" % (self.R(op.args[0]), self.R(op.args[1])))

    def op_texture(self, op):
        """llvm_gen_texture (llvm_gen.cpp:2715-2830) -> osl_texture (optexture.cpp:235-310):
        result, filename, s, t, [dsdx, dtdx, dsdy, dtdy,] then "name", value option pairs
        (llvm_gen_texture_options, llvm_gen.cpp:2480-2713).  Derivatives of s and t come
        from the Dual2 arguments unless given explicitly."""
        A = op.args
        d, fn, s, t = A[:4]
        i = 4

        def dpart(sym, which):
            return "(%s).%s" % (self.R(sym), which) if sym.has_derivs else "0.0f"
        if len(A) >= 8 and all(a.t.base in ("float", "int") for a in A[4:8]):
            dd = [self.fl(a) for a in A[4:8]]
            i = 8
        else:
            dd = [dpart(s, "dx"), dpart(t, "dx"), dpart(s, "dy"), dpart(t, "dy")]
        self.w("{")
        self.w("    TexOpt o_;")
        missing = {}
        while i < len(A):
            key, val = A[i], A[i + 1]
            i += 2
            if not key.constval:
                raise NotImplementedError("texture option with a non-constant name")
            k = key.vals[0]
            if k in ("wrap", "swrap", "twrap"):
                for f in (("swrap", "twrap") if k == "wrap" else (k,)):
                    self.w("    o_.%s = tex_wrap_code(%s);" % (f, self.R(val)))
            elif k in ("width", "swidth", "twidth", "blur", "sblur", "tblur"):
                for f in (("s" + k, "t" + k) if k in ("width", "blur") else (k,)):
                    self.w("    o_.%s = %s;" % (f, self.fl(val)))
            elif k == "fill":
                self.w("    o_.fill = %s;" % self.fl(val))
            elif k == "interp":
                self.w("    o_.interp = tex_interp_code(%s);" % self.R(val))
            elif k in ("missingcolor", "missingalpha", "alpha"):
                missing[k] = val
            else:
                raise NotImplementedError("texture option '%s'" % k)
        nch = 3 if d.t.triple else 1
        self.w("    float r_[4];")
        self.w("    const bool found_ = texture_lookup(%s, o_, %s, %s, %s, %s, %s, %s, %d, r_);" % (
            self.R(fn), self.fl(s), self.fl(t), dd[0], dd[1], dd[2], dd[3], nch))
        if missing:
            # a file that cannot be had, with "missingcolor" / "missingalpha" given: no error, the result is
            # the missing colour (zero when only the alpha was given) and alpha the missing alpha
            # (llvm_gen_texture_options, llvm_gen.cpp:2611-2640; osl_texture, optexture.cpp:283-300)
            mc = missing.get("missingcolor")
            self.w("    if (!found_) {")
            for c in range(nch):
                self.w("        r_[%d] = %s;" % (c, self.comp(mc, c if mc.t.triple else 0, False) if mc is not None else "0.0f"))
            if "alpha" in missing:
                self.w("        assign(%s, %s);" % (self.R(missing["alpha"]), self.fl(missing["missingalpha"])
                                                   if "missingalpha" in missing else "0.0f"))
            self.w("    }")
            if "alpha" in missing:
                self.w('    else unsupported_at_runtime("texture option \'alpha\' of an existing image");')
        elif False:
            pass
        self.w("    (void)found_;")
        self.w("    assign(%s, %s);" % (self.R(d), "V3(r_[0], r_[1], r_[2])" if nch == 3 else "r_[0]"))
        self.w("}")

    def op_determinant(self, op):
        self.w("assign(%s, m44_determinant(%s));" % (self.R(op.args[0]), self.R(op.args[1])))

    def op_transform(self, op):
        """llvm_gen_transform (llvm_gen.cpp:2376-2466): transform / transformv / transformn with
        a matrix, or with (from, to) space names through osl_transform_triple."""
        A = op.args
        vt = {"transform": 0, "transformv": 1, "transformn": 2}[op.name]
        d, p = A[0], A[-1]
        dv = d.has_derivs and p.has_derivs
        pe = self.R(p)
        if p.isconst:
            pe = "V3(%s)" % ", ".join(self.comp(p, k, False) for k in range(3))
        elif p.has_derivs and not dv:
            pe = "nd(%s)" % pe
        if len(A) == 3 and A[1].t.base == "matrix":
            return self.w("assign(%s, m44_transform(%s, %s, %d));" % (self.R(d), self.R(A[1]), pe, vt))
        fs = None if len(A) == 3 else A[1]
        ts = A[1] if len(A) == 3 else A[2]
        if (fs is None or fs.constval) and ts.constval:
            norm = lambda s: "common" if s in ("common", "world") else s
            if norm(fs.vals[0] if fs is not None else "common") == norm(ts.vals[0]):
                return self.w("assign(%s, %s);" % (self.R(d), pe))   # identity: just copy
        if len(A) == 3:      # transform("to", p): from = "common"
            frm, to = '"common"', self.R(A[1])
        else:
            frm, to = self.R(A[1]), self.R(A[2])
        T = "Dv" if dv else "V3"
        self.w("{ %s o_; xf_transform_triple_err(sg, xf_set(L), %s, %s, %s(%s), o_, %d); assign(%s, o_); }" % (
            T, frm, to, T, pe, vt, self.R(d)))

    op_transformv = op_transform
    op_transformn = op_transform

    def op_spline(self, op):
        """llvm_gen_spline (llvm_gen.cpp:3673-3740) -> osl_spline_* (opspline.cpp)."""
        A = op.args
        d, basis, x = A[0], A[1], A[2]
        knots = A[4] if len(A) == 5 else A[3]
        count = self.R(A[3]) if len(A) == 5 else str(knots.t.arr)
        dv = d.has_derivs and (x.has_derivs or knots.has_derivs)
        bt = str(SPLINE_TYPES.get(basis.vals[0], 4)) if basis.constval else "spline_type(%s)" % self.R(basis)
        xe = self.R(x)
        if x.has_derivs and not dv:
            xe = "nd(%s)" % xe
        if op.name == "splineinverse":
            self.w("{ float k_[%d]; for (int i_ = 0; i_ < %d; ++i_) k_[i_] = nd(%s[i_]);" % (
                knots.t.arr, knots.t.arr, self.R(knots)))
            # derivatives only flow from x (osl_splineinverse_dfdff / _dfdfdf ignore knot derivs)
            xi = self.R(x) if (d.has_derivs and x.has_derivs) else "nd(%s)" % self.R(x)
            self.w("  assign(%s, spline_inverse(%s, k_, %s, %s)); }" % (self.R(d), xi, count, bt))
            return
        if knots.has_derivs and not dv:
            kt = "V3" if knots.t.triple else "float"
            self.w("%s k_[%d]; for (int i_ = 0; i_ < %d; ++i_) k_[i_] = nd(%s[i_]);" % (
                kt, knots.t.arr, knots.t.arr, self.R(knots)))
            ke = "k_"
        else:
            ke = self.R(knots)
        self.w("spline_eval(%s, %s, %s, %s, %s);" % (self.R(d), xe, ke, count, bt))

    op_splineinverse = op_spline

    def op_printf(self, op):
        A = op.args
        fmt = A[0]
        if not fmt.constval:
            raise NotImplementedError("printf with non-constant format")
        self.emit_format(fmt.vals[0], A[1:])

    def scan_messages(self):
        """one group-data slot per constant message name, typed by the first setmessage of that
        name in layer order (opmessage.cpp keeps a per-execution list; same observable behaviour
        for int / float / triple values)"""
        self.messages = {}
        for l in self.g.layers:
            if l.unused:
                continue
            for op in l.m.ops:
                if op.name != "setmessage" or len(op.args) != 2:
                    continue
                nm, v = op.args
                if nm.t.base != "string" or not nm.constval or v.t.arr:
                    continue
                if not (v.t.base in ("int", "float") or v.t.triple):
                    continue
                if nm.vals[0] not in self.messages and len(self.messages) < 31:
                    ctype = "int" if v.t.base == "int" else ("V3" if v.t.triple else "float")
                    self.messages[nm.vals[0]] = (len(self.messages), ctype)

    def message_fields(self):
        out = []
        if self.messages:
            out.append("    unsigned msgset = 0u;")
            for name, (k, ctype) in self.messages.items():
                out.append("    %s M%d;" % (ctype, k))
        return out

    def _msg_ctype(self, s):
        return "int" if s.t.base == "int" else ("V3" if s.t.triple else ("float" if s.t.base == "float" else None))

    def op_setmessage(self, op):
        nm, v = op.args
        slot = self.messages.get(nm.vals[0]) if (nm.t.base == "string" and nm.constval) else None
        if slot and not v.t.arr and self._msg_ctype(v) == slot[1]:
            e = self.R(v)
            if v.has_derivs:
                e = "nd(%s)" % e
            self.w("if (!(gd.msgset & %du)) { gd.M%d = %s; gd.msgset |= %du; }" % (1 << slot[0], slot[0], e, 1 << slot[0]))

    def op_getmessage(self, op):
        A = op.args
        res, dst = self.R(A[0]), A[-1]
        nm = A[1]
        slot = self.messages.get(nm.vals[0]) if (len(A) == 3 and nm.t.base == "string" and nm.constval) else None
        if slot and not dst.t.arr and self._msg_ctype(dst) == slot[1]:
            val = "gd.M%d" % slot[0]
            if dst.has_derivs:
                val = ("Dv(%s)" if dst.t.triple else "Df(%s)") % val
            self.w("if (gd.msgset & %du) { %s = %s; %s = 1; } else %s = 0;" % (1 << slot[0], self.R(dst), val, res, res))
        else:
            self.w("%s = 0;" % res)

    def op_getattribute(self, op):
        """llvm_gen_getattribute (llvm_gen.cpp:3303-3420) -> RendererServices::get_attribute.  The
        harness renderer answers "osl:version" / "shading:index" and otherwise falls back to
        per-point userdata of that name and type (simplerend.cpp:480-503); object and array
        index lookups find nothing here."""
        A = op.args
        res, dst = self.R(A[0]), A[-1]
        nm = A[1]

        def attr_stores(val):
            """statements handing a uniform renderer attribute to dst, or None when the type asked for differs"""
            v = list(val) if isinstance(val, (list, tuple)) else [val]
            per = 3 if dst.t.triple else 1
            cnt = dst.t.arr or 1
            d = self.R(dst)
            el = (lambda e: "%s[%d]" % (d, e)) if dst.t.arr else (lambda e: d)
            if dst.t.base == "int":
                if len(v) != cnt or not all(isinstance(x, int) and not isinstance(x, bool) for x in v):
                    return None
                return " ".join("%s = %d;" % (el(e), v[e]) for e in range(cnt))
            if dst.t.base == "string":
                if len(v) != cnt or not all(isinstance(x, str) for x in v):
                    return None
                return " ".join("%s = %s;" % (el(e), cstr(v[e])) for e in range(cnt))
            if dst.t.base == "float" or dst.t.triple:
                if len(v) != cnt * per or not all(isinstance(x, float) for x in v):
                    return None
                if per == 1:
                    return " ".join("assign(%s, %s);" % (el(e), cfloat(v[e])) for e in range(cnt))
                return " ".join("assign(%s, V3(%s, %s, %s));" % (el(e), cfloat(v[3 * e]), cfloat(v[3 * e + 1]),
                                                                 cfloat(v[3 * e + 2])) for e in range(cnt))
            return None

        if len(A) == 3 and nm.t.base == "string" and self.g.attributes:
            if nm.constval:
                st = attr_stores(self.g.attributes[nm.vals[0]]) if nm.vals[0] in self.g.attributes else None
                if st:
                    self.w("%s %s = 1;" % (st, res))
                    return
            elif not nm.t.arr:
                # a name the shader computes: compared with the renderer's table at run time
                self.w("{ const char* nm_ = %s; %s = 0;" % (self.R(nm), res))
                for k, val in self.g.attributes.items():
                    st = attr_stores(val)
                    if st:
                        self.w("  if (nm_ && !strcmp(nm_, %s)) { %s %s = 1; }" % (cstr(k), st, res))
                self.w("}")
                return
        if len(A) == 3 and nm.t.base == "string" and nm.constval and not dst.t.arr:
            name = nm.vals[0]
            if name == "osl:version" and dst.t.base == "int":
                self.w("%s = 11600; %s = 1;" % (self.R(dst), res))
                return
            if name == "shading:index" and dst.t.base == "int":
                self.w("%s = sg.shadeindex; %s = 1;" % (self.R(dst), res))
                return
            # folded at optimisation time when the name is a constant (constfold_getattribute,
            # constfold.cpp: "shader:shadername" / "shader:layername" / "shader:groupname")
            shader_attr = {"shader:shadername": self.l.m.name, "shader:layername": self.l.name,
                           "shader:groupname": self.g.name}
            if name in shader_attr and dst.t.base == "string":
                self.w("%s = %s; %s = 1;" % (self.R(dst), cstr(shader_attr[name]), res))
                return
            if dst.t.base in ("int", "float") or dst.t.triple:
                self.w("%s = bind_userdata(L, sg, %s, %s) ? 1 : 0;" % (res, cstr(name), self.R(dst)))
                return
        self.w("%s = 0;" % res)

    def op_error(self, op, kind="error"):
        """osl_error / osl_warning (llvm_gen_printf with the error / warning flavours): the
        formatted message goes to the error handler, which testshade prints as
        "ERROR: Shader error [<shader>]: <message>" followed by a newline; a message already
        reported is dropped (ShadingSystem attribute error_repeats = 0, the default)."""
        A = op.args
        if not A[0].constval:
            raise NotImplementedError("%s with non-constant format" % kind)
        self.w("if (sg.ctx && sg.ctx->out) {")
        self.ind += 1
        self.w("std::string msg_; std::string* save_ = sg.ctx->out; sg.ctx->out = &msg_;")
        self.emit_format(A[0].vals[0], A[1:])
        self.w("sg.ctx->out = save_;")
        self.w("report_message(sg, %s, msg_);" % cstr("%s: Shader %s [%s]: " % (kind.upper(), kind, self.l.m.name)))
        self.ind -= 1
        self.w("}")

    def op_warning(self, op):
        return self.op_error(op, "warning")

    def emit_format(self, f, args):
        """Split an OSL format string at gen time (reference: llvm_gen_printf,
        llvm_gen.cpp:350-700: one C conversion per component, space separated)."""
        ai = 0
        i = 0
        lit = ""
        while i < len(f):
            ch = f[i]
            if ch != "%":
                lit += ch
                i += 1
                continue
            if f[i + 1:i + 2] == "%":
                lit += "%"
                i += 2
                continue
            j = i + 1
            while j < len(f) and f[j] not in "cdefgimnopsuvxXEG":
                j += 1
            spec, conv = f[i:j + 1], f[j]
            i = j + 1
            if lit:
                self.w("pf_lit(sg, %s);" % cstr(lit))
                lit = ""
            a = args[ai]
            ai += 1
            n = max(1, a.t.arr)
            for e in range(n):
                r = self.R(a) + ("[%d]" % e if a.t.arr else "")
                if e:
                    self.w('pf_lit(sg, " ");')
                if a.t.base == "string":
                    self.w("pf_s(sg, %s, %s);" % (cstr(spec[:-1] + "s"), r))
                elif a.t.base == "int":
                    if conv in "dioxXuc":   # integer conversions keep the int (llvm_gen_printf)
                        self.w("pf_i(sg, %s, %s);" % (cstr(spec), r))
                    else:
                        self.w("pf_f(sg, %s, %s);" % (cstr(spec), r))
                elif a.t.base == "float":
                    if conv in "dixX":
                        self.w("pf_i(sg, %s, nd(%s));" % (cstr(spec), r))
                    else:
                        self.w("pf_f(sg, %s, %s);" % (cstr(spec), r))
                elif a.t.triple or a.t.base == "matrix":
                    self.w("pf_v(sg, %s, %s);" % (cstr(spec), r))
                else:
                    raise NotImplementedError("printf of %s" % a.t)
        if lit:
            self.w("pf_lit(sg, %s);" % cstr(lit))

    def op_closure_arith(self, op):
        """mul / add on closures (llvm_gen_mul / llvm_gen_add closure branches ->
        osl_mul_closure_{float,color}, osl_add_closure_closure)."""
        d, a, b = op.args
        if op.name == "add":
            self.w("%s = clos_add(sg.pool, %s, %s);" % (self.R(d), self.R(a), self.R(b)))
            return
        if op.name != "mul":
            raise NotImplementedError("closure op %s" % op.name)
        if a.t.base != "closure color":
            a, b = b, a
        w = self.R(b)
        if b.has_derivs:
            w = "nd(%s)" % w
        if b.t.base == "int":
            w = "(float)%s" % w
        self.w("%s = clos_mul(sg.pool, %s, %s);" % (self.R(d), self.R(a), w))

    def op_closure(self, op):
        """llvm_gen_closure (llvm_gen.cpp:3786-3903): allocate a component,
        optionally weighted, and copy the positional params into its block."""
        A = list(op.args)
        d = A[0]
        rest = A[1:]
        weight = None
        if rest[0].t.base != "string":
            weight = rest[0]
            rest = rest[1:]
        name = rest[0].vals[0]
        params = []
        for a in rest[1:]:
            if a.t.base == "string" and not params and False:
                break
            params.append(a)
        # keyword params ("label", value) are trailing string/value pairs: drop them
        pos = []
        i = 0
        while i < len(params):
            a = params[i]
            if a.t.base == "string" and a.constval and i + 1 < len(params) and \
                    (name, len(pos)) in CLOSURES:
                break
            pos.append(a)
            i += 1
        key = (name, len(pos))
        if key not in CLOSURES:
            raise NotImplementedError("closure %s with %d params is not registered" % key)
        # closure-typed parameters (layer's top / base) are host pointers: two words each
        wsize = lambda a: 2 if a.t.base == "closure color" else a.t.ncomp
        nwords = sum(wsize(a) for a in pos)
        keys = CLOSURE_KEYS.get(CLOSURES[key], [])
        given = {}
        kw = params[len(pos):]
        for j in range(0, len(kw) - 1, 2):
            if kw[j].t.base == "string" and kw[j].constval:
                given[kw[j].vals[0]] = kw[j + 1]
        key_base = nwords
        nwords += sum(KEY_WORDS[t] for _, t in keys)
        wexpr = "nullptr"
        if weight is not None:
            self.w("V3 w_; assign(w_, %s);" % self.R(weight))
            wexpr = "&w_"
        self.w("ClosComp* c_ = clos_component(sg.pool, %s, %d, %s);" % (CLOSURES[key], nwords, wexpr))
        self.w("if (c_) {")
        off = 0
        for a in pos:
            e = self.R(a)
            if a.has_derivs:
                e = "nd(%s)" % e
            self.w("    putp(c_->params + %d, %s);" % (off, e))
            off += wsize(a)
        koff = key_base
        for kname, ktype in keys:
            a = given.get(kname)
            if a is not None and a.t.base == ktype and not a.t.arr:
                e = self.R(a)
                if a.has_derivs:
                    e = "nd(%s)" % e
                self.w("    putp(c_->params + %d, %s);" % (koff, e))
            else:
                zero = {"int": "0", "float": "0.0f", "color": "V3(0.0f)"}[ktype]
                self.w("    putp(c_->params + %d, %s);" % (koff, zero))
            koff += KEY_WORDS[ktype]
        self.w("}")
        self.w("%s = c_;" % self.R(d))


RAYTYPES = ["camera", "shadow", "reflection", "refraction", "diffuse", "glossy",
            "subsurface", "displacement"]


def raytype_bit(name):
    """testshade/simplerend.cpp raytype names -> bit (1<<index)."""
    return (1 << RAYTYPES.index(name)) if name in RAYTYPES else 0


RUNNER_TAIL = r"""
extern "C" void oracle_texture_add(const char* name, int w, int h, int nch, const float* px)
{
    oracle_texture_add_impl(name, w, h, nch, px);
}
extern "C" void oracle_run_mt(const Launch* L, long long n, int nthreads)
{
    if (nthreads <= 1) { oracle_run(L, 0, n, nullptr); return; }
    std::vector<std::thread> th;
    long long chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        long long b = t * chunk, e = std::min(n, b + chunk);
        if (b >= e) break;
        th.emplace_back([=] { oracle_run(L, b, e, nullptr); });
    }
    for (auto& t : th) t.join();
}
extern "C" void oracle_set_error_repeats(int v) { oracle_error_repeats() = v; }
extern "C" const char* oracle_run_capture(const Launch* L, long long begin, long long end)
{
    static std::string buf;
    buf.clear();
    oracle_run(L, begin, end, &buf);
    return buf.c_str();
}
"""

HERE = os.path.dirname(os.path.abspath(__file__))


RENDER_TAIL = r"""
extern "C" void oracle_texture_add(const char* name, int w, int h, int nch, const float* px)
{
    oracle_texture_add_impl(name, w, h, nch, px);
}
// energy tables of the libbsdl microfacet lobes (data; see osl_oracle_mxlobes.h)
extern "C" void oracle_set_bsdl_luts(const float* p) { bsdl_luts() = p; }
static const Background* g_background = nullptr;
static void oracle_render_rows(const RenderScene* S, int y0, int y1, float* out, std::string* pf)
{
    Ctx ctx{pf};
    Renderer R{*S, g_shaders, &ctx};
    R.background = g_background;
    for (int y = y0; y < y1; ++y)
        for (int x = 0; x < S->xres; ++x) {
            V3 c = R.antialias_pixel(x, y);
            float* p = out + 3 * ((long long)y * S->xres + x);
            p[0] = c.x; p[1] = c.y; p[2] = c.z;
        }
}
// rows [y0, y1) of the image (pixels are independent: a band equals the same rows of a full render)
extern "C" void oracle_render_band(const RenderScene* S0, float* out, int nthreads, int y0, int y1)
{
    RenderScene S1 = *S0;
    camera_finalize(S1);
    const RenderScene* S = &S1;
    // background importance table (simpleraytracer.cpp:1232-1249)
    Background bg;
    g_background = nullptr;
    if (S->background_resolution > 0 && S->background_shader >= 0) {
        Ctx ctx{nullptr};
        Renderer R{*S, g_shaders, &ctx};
        ClosurePool pool;
        bg.prepare(S->background_resolution, [&](const Dv& d) { return R.eval_background(d, -1, pool); });
        g_background = &bg;
    }
    // scanline-parallel like the reference (parallel_for_chunked, simpleraytracer.cpp:1428)
    y0 = std::max(y0, 0);
    y1 = std::min(y1, S->yres);
    if (nthreads <= 1) { oracle_render_rows(S, y0, y1, out, nullptr); return; }
    std::vector<std::thread> th;
    std::atomic<int> next{y0};
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&] {
            for (;;) {
                int y = next.fetch_add(2);
                if (y >= y1) break;
                oracle_render_rows(S, y, std::min(y1, y + 2), out, nullptr);
            }
        });
    for (auto& t : th) t.join();
}
extern "C" void oracle_render(const RenderScene* S0, float* out, int nthreads)
{
    oracle_render_band(S0, out, nthreads, 0, S0->yres);
}
"""



# ---------------------------------------------------------------------------
# 16-wide batched restatement (the reference's BatchedExecutor<16> path:
# testshade.cpp:1829-1922 batched_shade_region, batched_llvm_gen.cpp).  Every
# symbol is a block of WIDTH lanes, every op is a loop over the lanes enabled in
# the current execution mask (the reference's wide ops are OMP-simd lane loops
# too: wide/*.cpp), `if` splits the mask, loops run while any lane is active,
# `return` retires lanes until the end of the function, a lazily evaluated
# upstream layer runs for (mask & ~already-run lanes) (batched_llvm_gen.cpp:116-137).
# Used as the batched CPU baseline of bench.py and pinned against the scalar
# oracle by tests/test_oracle_wide.py.  printf text is dropped (no journal); closures /
# break / continue / exit are not restated in this mode (NotImplementedError).
# ---------------------------------------------------------------------------
WIDTH = 16

WIDE_PRELUDE = r"""
#define OSLO_W %d
#define OSLO_LANES(m) _Pragma("omp simd") for (int l_ = 0; l_ < OSLO_W; ++l_) if (((m) >> l_) & 1u)
""" % WIDTH


class WideGen(Gen):
    def ref(self, l, s):
        e = Gen.ref(self, l, s)
        if s.symtype in ("const", "global"):
            return e
        return e + "[l_]"

    def eff(self, ctx):
        r = ctx.get("ret")
        return "(%s & ~%s)" % (ctx["mask"], r) if r else ctx["mask"]

    def generate(self):
        g = self.g
        o = self.out
        o.append("// generated by oracle/oso2cpp.py (wide mode) — CPU oracle, test infrastructure only")
        o.append('#include "osl_oracle_closure.h"')
        o.append('#include "osl_oracle_runtime.h"')
        o.append(WIDE_PRELUDE)
        o.append("using namespace oslo;")
        o.append("namespace {")
        o.append("struct GD {")
        o.append("    unsigned ran[%d];   // lanes for which the layer has run" % max(1, len(g.layers)))
        for l in g.layers:
            if l.unused:
                continue
            for s in l.m.syms:
                if s.symtype in ("param", "oparam"):
                    arr = "[%d]" % s.t.arr if s.t.arr else ""
                    o.append("    %s L%d_%s[OSLO_W]%s;" % (self.ctype(s), l.idx, self.ident(s.name), arr))
        o.append("};")
        used = [l for l in g.layers if not l.unused]
        for l in used:
            o.append("static void layer_%d(SG* sgw, GD& gd, const Launch* L, unsigned m0);" % l.idx)
        for l in used:
            self.gen_layer(l)
        o.append("}  // namespace")
        o.append('extern "C" void oracle_run(const Launch* L, long long begin, long long end, std::string* pf)')
        o.append("{")
        o.append("    Ctx ctx{pf};")
        o.append("    for (long long i0 = begin; i0 < end; i0 += OSLO_W) {")
        o.append("        SG sgw[OSLO_W]; GD gd;")
        o.append("        const int nb = (int)std::min<long long>(OSLO_W, end - i0);")
        o.append("        const unsigned m0 = nb >= 32 ? ~0u : ((1u << nb) - 1u);")
        o.append("        for (int l = 0; l < nb; ++l) load_sg(sgw[l], L, i0 + l, &ctx);")
        o.append("        for (int k = 0; k < %d; ++k) gd.ran[k] = 0u;" % max(1, len(g.layers)))
        o.append("        layer_%d(sgw, gd, L, m0);" % (len(g.layers) - 1))
        o.append("    }")
        o.append("}")
        o.append(RUNNER_TAIL)
        return "\n".join(o) + "\n"

    def lanes(self, mask):
        self.w("OSLO_LANES(%s) {" % mask)
        self.ind += 1
        self.w("SG& sg = sgw[l_]; (void)sg;")

    def end_lanes(self):
        self.ind -= 1
        self.w("}")

    def gen_layer(self, l):
        g = self.g
        m = l.m
        self.l = l
        self.ind = 0
        self.w("static void layer_%d(SG* sgw, GD& gd, const Launch* L, unsigned m0)" % l.idx)
        self.w("{")
        self.ind = 1
        self.w("gd.ran[%d] |= m0;" % l.idx)
        self.w("(void)L;")
        for s in m.syms:
            if s.symtype in ("local", "temp"):
                arr = "[%d]" % s.t.arr if s.t.arr else ""
                self.w("%s %s[OSLO_W]%s = {};" % (self.ctype(s), self.ident(s.name), arr))
            elif s.symtype == "const" and s.t.arr:
                vals = self.initval(s)
                self.w("static const %s K_%s[%d] = {%s};" % (
                    self.ctype(s), self.ident(s.name), s.t.arr, ", ".join(vals)))
        ctx = dict(mask="m0")
        if l.idx == len(g.layers) - 1:
            for e in g.layers[:-1]:
                if not e.unused and not e.lazy:
                    self.w("if (m0 & ~gd.ran[%d]) layer_%d(sgw, gd, L, m0 & ~gd.ran[%d]);" % (e.idx, e.idx, e.idx))
        for s in m.syms:
            if s.symtype not in ("param", "oparam"):
                continue
            if s.connected_from is not None:
                continue
            vals = self.initval(s)
            r = self.ref(l, s)
            self.lanes("m0")
            if s.t.arr:
                for i, v in enumerate(vals):
                    self.w("%s[%d] = %s;" % (r, i, v))
            else:
                self.w("%s = %s;" % (r, vals[0]))
            self.end_lanes()
            if s.initexpr and s.name in m.methods:
                b, e = m.methods[s.name]
                self.ensured = set()
                self.emit_block(b, e, ctx)
        self.ensured = set()
        b, e = m.methods.get("___main___", (0, 0))
        self.emit_block(b, e, ctx)
        self.lanes("m0")
        for (si, ssym, di, dsym) in g.connections:
            if si != l.idx or g.layers[di].unused:
                continue
            self.emit_copy(g.layers[di], dsym, l, ssym)
        for (li, s) in g.outputs:
            if li != l.idx:
                continue
            ro = s.renderer_output
            fn = "wrd" if ro["derivs"] else "wr"
            self.w("%s(outp(L, sg, %d, %d), %s);" % (fn, ro["offset"], ro["stride"], self.ref(l, s)))
        self.end_lanes()
        self.ind = 0
        self.w("}")

    def mask_where(self, var, mask, cond, negate=False):
        """var = lanes of `mask` whose condition symbol is true (false with negate)"""
        self.w("for (int l_ = 0; l_ < OSLO_W; ++l_) if (((%s) >> l_) & 1u) { if (%s(%s)) %s |= 1u << l_; }"
               % (mask, "!" if negate else "", cond, var))

    def emit_block(self, b, e, ctx):
        ops = self.l.m.ops
        i = b
        while i < e:
            op = ops[i]
            n = op.name
            if n == "if":
                self.useparams(op, ctx)
                self.label += 1
                k = self.label
                self.w("unsigned mt_%d = 0u;" % k)
                self.mask_where("mt_%d" % k, self.eff(ctx), self.ref(self.l, op.args[0]))
                saved = set(self.ensured)
                self.w("if (mt_%d) {" % k)
                self.ind += 1
                self.emit_block(i + 1, op.jumps[0], dict(ctx, mask="mt_%d" % k))
                self.ind -= 1
                self.w("}")
                self.ensured = set(saved)
                if op.jumps[1] > op.jumps[0]:
                    self.w("const unsigned me_%d = %s & ~mt_%d;" % (k, self.eff(ctx), k))
                    self.w("if (me_%d) {" % k)
                    self.ind += 1
                    self.emit_block(op.jumps[0], op.jumps[1], dict(ctx, mask="me_%d" % k))
                    self.ind -= 1
                    self.w("}")
                    self.ensured = set(saved)
                i = op.jumps[1]
            elif n in ("for", "while", "dowhile"):
                cl, bl, il, dl = op.jumps
                self.label += 1
                k = self.label
                self.emit_block(i + 1, cl, ctx)
                saved = set(self.ensured)
                cond = self.ref(self.l, op.args[0])
                self.w("unsigned ml_%d = %s;" % (k, self.eff(ctx)))
                c2 = dict(ctx, mask="ml_%d" % k, loop=True)
                self.w("for (;;) {")
                self.ind += 1
                if n == "dowhile":
                    self.emit_block(bl, il, c2)
                    self.emit_block(cl, bl, c2)
                else:
                    self.emit_block(cl, bl, c2)
                self.w("{ unsigned keep_ = 0u;")
                self.mask_where("keep_", self.eff(c2), cond)
                self.w("  ml_%d = keep_; }" % k)
                self.w("if (!ml_%d) break;" % k)
                if n == "dowhile":
                    self.emit_block(il, dl, c2)
                else:
                    self.emit_block(bl, il, c2)
                    self.emit_block(il, dl, c2)
                self.ind -= 1
                self.w("}")
                self.ensured = set(saved)
                i = dl
            elif n == "functioncall":
                self.label += 1
                k = self.label
                # lanes that executed `return` inside the body stay off until its end
                self.w("unsigned mr_%d = %s;" % (k, ("~" + self.eff(ctx)) if ctx.get("ret") else "0u"))
                saved = set(self.ensured)
                self.w("{")
                self.ind += 1
                self.emit_block(i + 1, op.jumps[0], dict(ctx, ret="mr_%d" % k))
                self.ind -= 1
                self.w("}")
                self.ensured = set(saved)
                i = op.jumps[0]
            elif n == "return":
                if not ctx.get("ret"):
                    raise NotImplementedError("wide oracle: 'return' outside a function")
                self.w("%s |= %s;" % (ctx["ret"], ctx["mask"]))
                i += 1
            elif n in ("break", "continue", "exit"):
                raise NotImplementedError("wide oracle: op '%s' is not restated in batched mode" % n)
            elif n in ("nop", "end", "useparam"):
                i += 1
            else:
                if n in ("printf", "error", "warning", "fprintf"):
                    i += 1      # no journal in batched mode: the text is dropped (results are unaffected)
                    continue
                if n == "closure":
                    raise NotImplementedError("wide oracle: op '%s' is not restated in batched mode" % n)
                self.useparams(op, ctx)
                self.lanes(self.eff(ctx))
                self.emit_op(op)
                self.end_lanes()
                i += 1

    def useparams(self, op, ctx=None):
        for a, c in zip(op.args, op.rw):
            if c in "rW" and a.connected_from is not None:
                up = a.connected_from[0]
                if up in self.ensured:
                    continue
                self.ensured.add(up)
                m = self.eff(ctx)
                self.w("if (%s & ~gd.ran[%d]) layer_%d(sgw, gd, L, %s & ~gd.ran[%d]);" % (m, up, up, m, up))


def build_group_wide(group, workdir=None):
    """The batched restatement, built the way testshade --batched runs: native ISA (AVX-512 when
    the host has it), FMA contraction allowed (testshade.cpp:294-298 turns llvm_jit_fma on)."""
    return _compile(WideGen(group).generate(), workdir, "-O3",
                    ("-march=native", "-fopenmp-simd", "-ffp-contract=fast"))


def generate_render_module(groups):
    """One translation unit holding every material group of a scene plus the
    restated path tracer (osl_oracle_render.h)."""
    out = ["// generated by oracle/oso2cpp.py — CPU oracle render module, test infrastructure only",
           '#include "osl_oracle_render.h"', "#include <atomic>", "using namespace oslo;"]
    for k, g in enumerate(groups):
        out.append(Gen(g).generate_material("mat%d" % k))
    out.append("static const ShaderFn g_shaders[] = {%s};" % ", ".join(
        "mat%d::entry" % k for k in range(len(groups))))
    out.append(RENDER_TAIL)
    return "\n".join(out) + "\n"


def build_render(groups, workdir=None, opt="-O2", extra_flags=()):
    return _compile(generate_render_module(groups), workdir, opt, extra_flags)


def build_group(group, workdir=None, opt="-O2", extra_flags=()):
    """Generate + compile; returns path of the shared object."""
    return _compile(Gen(group).generate(), workdir, opt, extra_flags)


def _compile(src, workdir=None, opt="-O2", extra_flags=()):
    workdir = workdir or os.path.join(HERE, "_build")
    os.makedirs(workdir, exist_ok=True)
    hdrs = b""
    for h in sorted(f for f in os.listdir(HERE) if f.endswith(".h")):
        with open(os.path.join(HERE, h), "rb") as f:
            hdrs += f.read()
    key = hashlib.sha1(src.encode() + hdrs + opt.encode() + " ".join(extra_flags).encode()).hexdigest()[:16]
    so = os.path.join(workdir, "oracle_%s.so" % key)
    if not os.path.exists(so):
        cpp = os.path.join(workdir, "oracle_%s.cpp" % key)
        with open(cpp, "w") as f:
            f.write(src)
        cmd = ["g++", "-std=c++17", opt, "-ffp-contract=off", "-fPIC", "-shared", "-pthread",
               "-I", HERE, cpp, "-o", so + ".tmp"] + list(extra_flags)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle compile failed:\n" + r.stderr[:6000])
        os.replace(so + ".tmp", so)
    return so
