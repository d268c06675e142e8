// osl_b200_codegen.cpp — shader group -> CUDA C++ text for sm_100a (product code).
//
// This is the B200 target of the reference's code generators: it fills the
// role of BackendLLVM::run + llvm_gen_* for the GPU
// (src/liboslexec/llvm_instance.cpp:1288-1870, 2083-2572; llvm_gen.cpp), with
// the structure of BackendCpp (src/liboslexec/backendcpp.cpp) — a group-data
// struct, one device function per used layer, lazy upstream calls guarded by
// run bits, and an entry kernel — but emitting code against the device
// library in csrc/device/osl_b200_device.cuh.  One thread shades one point;
// only the ShaderGlobals fields the group reads are loaded (coalesced SoA
// planes); group data lives in registers.
#include "../../../include/osl_b200.h"
#include "osl_b200_group.h"
#include "osl_b200_texture.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <sstream>
#include <stdexcept>

namespace oslb200 {

namespace {

std::string
cfloat(float f)
{
    if (std::isnan(f))
        return "__int_as_float(0x7fc00000)";
    if (std::isinf(f))
        return f > 0 ? "__int_as_float(0x7f800000)" : "__int_as_float(0xff800000)";
    char buf[64];
    snprintf(buf, sizeof buf, "%.9g", f);
    std::string s = buf;
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos)
        s += ".0";
    return s + "f";
}

std::string
ident(const std::string& n)
{
    std::string r;
    for (char c : n) {
        if (c == '$')
            r += "S_";
        else if (isalnum((unsigned char)c) || c == '_')
            r += c;
        else
            r += '_';
    }
    return r;
}

struct GlobalInfo {
    const char* expr;
    int field;          // value field
    int dx, dy;         // derivative fields or -1
    bool triple;
};
const std::map<std::string, GlobalInfo>&
global_table()
{
    static const std::map<std::string, GlobalInfo> t = {
        { "P", { "sg.P", B200_SG_P, B200_SG_dPdx, B200_SG_dPdy, true } },
        { "I", { "sg.I", B200_SG_I, B200_SG_dIdx, B200_SG_dIdy, true } },
        { "N", { "sg.N", B200_SG_N, -1, -1, true } },
        { "Ng", { "sg.Ng", B200_SG_Ng, -1, -1, true } },
        { "u", { "sg.u", B200_SG_u, B200_SG_dudx, B200_SG_dudy, false } },
        { "v", { "sg.v", B200_SG_v, B200_SG_dvdx, B200_SG_dvdy, false } },
        { "dPdu", { "sg.dPdu", B200_SG_dPdu, -1, -1, true } },
        { "dPdv", { "sg.dPdv", B200_SG_dPdv, -1, -1, true } },
        { "Ps", { "sg.Ps", B200_SG_Ps, B200_SG_dPsdx, B200_SG_dPsdy, true } },
        { "time", { "sg.time", B200_SG_time, -1, -1, false } },
        { "dtime", { "sg.dtime", B200_SG_dtime, -1, -1, false } },
        { "dPdtime", { "sg.dPdtime", B200_SG_dPdtime, -1, -1, true } },
    };
    return t;
}

const char* RAYTYPES[] = { "camera", "shadow", "reflection", "refraction",
                           "diffuse", "glossy", "subsurface", "displacement" };
int
raytype_bit(const std::string& n)
{
    for (int i = 0; i < 8; ++i)
        if (n == RAYTYPES[i])
            return 1 << i;
    return 0;
}

const std::set<std::string> UNARY
    = { "sin",  "cos",   "tan",   "asin",  "acos", "atan",  "sinh",  "cosh",        "tanh", "log",
        "log2", "log10", "exp",   "exp2",  "expm1", "erf",  "erfc",  "cbrt",        "sqrt",
        "inversesqrt", "abs", "fabs", "floor", "ceil", "round", "trunc", "sign", "logb", "neg" };
const std::set<std::string> BINARY = { "add", "sub", "mul", "div", "atan2", "pow", "fmod", "step", "min", "max" };
const std::set<std::string> TERNARY = { "mix", "clamp", "smoothstep", "select" };
const std::map<std::string, const char*> CMP
    = { { "eq", "==" }, { "neq", "!=" }, { "lt", "<" }, { "gt", ">" }, { "le", "<=" }, { "ge", ">=" } };
const std::map<std::string, const char*> INTBIN
    = { { "bitand", "&" }, { "bitor", "|" }, { "xor", "^" }, { "shl", "<<" }, { "shr", ">>" } };

// noise name -> (kind 0 perlin-unsigned, 1 perlin-signed, 2 cell, 3 hash)
int
noise_kind(const std::string& n, bool periodic)
{
    static const std::map<std::string, int> base
        = { { "noise", 0 },  { "uperlin", 0 },   { "snoise", 1 }, { "perlin", 1 },
            { "cell", 2 },   { "cellnoise", 2 }, { "hash", 3 },   { "hashnoise", 3 },
            { "simplex", 4 }, { "simplexnoise", 4 }, { "usimplex", 5 }, { "usimplexnoise", 5 },
            { "gabor", 6 } };
    static const std::map<std::string, int> per
        = { { "pnoise", 0 }, { "psnoise", 1 }, { "pcellnoise", 2 }, { "phashnoise", 3 } };
    if (periodic) {
        auto it = per.find(n);
        if (it != per.end())
            return it->second;
    }
    auto it = base.find(n);
    return it == base.end() ? -1 : it->second;
}

class Gen {
public:
    explicit Gen(Group& g) : g(g) {}
    std::string run();
    // group as `namespace <ns> { ... __device__ void entry(SG&); }` for the renderer
    std::string run_material(const std::string& ns);

private:
    Group& g;
    std::ostringstream o;
    int ind = 0, label = 0, loop_depth = 0;
    Layer* L = nullptr;
    int li   = 0;
    std::set<int> ensured;
    struct Ctx {
        std::string ret, brk, cont;
    };

    void w(const std::string& s)
    {
        for (int i = 0; i < ind; ++i)
            o << "    ";
        o << s << "\n";
    }
    [[noreturn]] void unsupported(const std::string& what)
    {
        throw std::runtime_error("B200 back end: " + what + " (layer '" + L->layername + "', shader '"
                                 + L->m.shadername + "')");
    }
    Symbol& S(int i) { return L->m.syms[i]; }

    std::string ctype(const Symbol& s)
    {
        switch (s.type.base) {
        case Base::Int: return "int";
        case Base::String: return "int";  // interned string id
        case Base::Closure: return "int";  // word offset into the closure pool
        case Base::Float: return s.has_derivs ? "Df" : "float";
        case Base::Color:
        case Base::Point:
        case Base::Vector:
        case Base::Normal: return s.has_derivs ? "Dv" : "V3";
        case Base::Matrix: return "M44";
        default: break;
        }
        unsupported("symbol '" + s.name + "' has a type the device path does not support yet");
    }
    static std::string m44_literal(const std::vector<float>& v, size_t at)
    {
        std::string r = "m44_make(";
        for (size_t k = 0; k < 16; ++k)
            r += (k ? ", " : "") + cfloat(at + k < v.size() ? v[at + k] : 0.0f);
        return r + ")";
    }
    // slot of a named coordinate system in the launch block's xf table
    int space_slot(const std::string& name)
    {
        for (size_t k = 0; k < g.spaces.size(); ++k)
            if (g.spaces[k] == name)
                return (int)k;
        if ((int)g.spaces.size() >= B200_MAX_SPACES)
            unsupported("more than " + std::to_string(B200_MAX_SPACES) + " named coordinate systems in one group");
        g.spaces.push_back(name);
        return (int)g.spaces.size() - 1;
    }
    // Expressions for osl_get_matrix(from) / osl_get_inverse_matrix(to) with a compile-time
    // space name: {matrix expression, ok expression}.  "common" and its synonym are identity.
    std::pair<std::string, std::string> space_matrix(int si, bool inverse)
    {
        const Symbol& s = S(si);
        if (s.type.base != Base::String || !s.const_value())
            unsupported("coordinate-system name that is not known at compile time");
        std::string name = s.svals.empty() ? "" : s.svals[0];
        if (name == "common" || name == g.commonspace_synonym)
            return { "m44_diag(1.0f)", "1" };
        std::string k = std::to_string(space_slot(name));
        if (journal_ok) {
            // unknown_coordsys_error = 1 (the default): a look-up of a name the renderer does not
            // know reports "Unknown transformation" through the error handler, at the op
            // (opmatrix.cpp:129-134, 160-165, 188-196)
            JournalFormat jf;
            jf.fmt        = "Unknown transformation \"" + name + "\"";
            for (size_t q = 0; q < jf.fmt.size(); ++q)
                if (jf.fmt[q] == '%')
                    jf.fmt.insert(q++, "%");
            jf.kind       = 3;  // an error of the shading system itself: no "Shader error [name]" prefix
            jf.shadername = L->m.shadername;
            int id        = (int)g.jformats.size();
            g.jformats.push_back(jf);
            w("if (!L.xf_ok[" + k + "]) { (void)jr_reserve(L, sg, " + std::to_string(id) + "u, 0u); }");
        }
        return { "m44_load(L.xf[" + k + "][" + (inverse ? "1" : "0") + "])", "L.xf_ok[" + k + "]" };
    }
    bool space_is(int si, const char* what)
    {
        const Symbol& s = S(si);
        return s.type.base == Base::String && s.const_value() && !s.svals.empty() && s.svals[0] == what;
    }
    std::string constexpr_(const Symbol& s)
    {
        if (s.type.arraylen)
            return "K_" + ident(s.name);
        switch (s.type.base) {
        case Base::Int: return std::to_string(s.ivals.empty() ? 0 : s.ivals[0]);
        case Base::Float: return cfloat(s.fvals.empty() ? 0.0f : s.fvals[0]);
        case Base::String: return std::to_string(g.intern(s.svals.empty() ? "" : s.svals[0]));
        default:
            if (s.type.is_triple()) {
                float v[3] = { 0, 0, 0 };
                for (size_t i = 0; i < 3 && i < s.fvals.size(); ++i)
                    v[i] = s.fvals[i];
                return "mkv(" + cfloat(v[0]) + ", " + cfloat(v[1]) + ", " + cfloat(v[2]) + ")";
            }
            if (s.type.base == Base::Matrix)
                return m44_literal(s.fvals, 0);
        }
        unsupported("constant '" + s.name + "' type");
    }
    std::string ref(int layer, const Symbol& s)
    {
        if (s.is_const())
            return constexpr_(s);
        if (s.is_param())
            return "gd.L" + std::to_string(layer) + "_" + ident(s.name);
        if (s.symtype == SymType::Global && s.name == "Ci")
            return "sg.Ci";
        if (s.symtype == SymType::Global) {
            auto it = global_table().find(s.name);
            if (it == global_table().end())
                unsupported("global '" + s.name + "'");
            const GlobalInfo& gi = it->second;
            g.globals_read.insert(gi.field);
            if (gi.dx >= 0) {
                if (s.has_derivs) {
                    g.globals_read.insert(gi.dx);
                    g.globals_read.insert(gi.dy);
                    // a global with derivatives that this layer WRITES ("P += ..." of a displacement
                    // shader) lives in a layer-local Dv, loaded at entry and stored back at the end
                    // (gen_layer); the accessor below is an rvalue
                    if (s.written)
                        return "gw_" + s.name;
                    return std::string(gi.expr) + "_d()";
                }
                return std::string(gi.expr);
            }
            return gi.expr;
        }
        return ident(s.name);
    }
    std::string R(int si) { return ref(li, S(si)); }

    // Would the reference's constant folder (constfold.cpp, runtimeoptimize.cpp) know this value at
    // optimisation time?  Constants and instance values do; so does a temporary or local written exactly
    // once, outside any conditional or loop, by a foldable op whose inputs fold.
    bool folds_to_constant(int si, int depth)
    {
        const Symbol& s = S(si);
        if (s.const_value())
            return true;
        if (depth > 16 || (s.symtype != SymType::Temp && s.symtype != SymType::Local) || s.type.arraylen)
            return false;
        static const std::set<std::string> foldable
            = { "assign", "add",  "sub",  "mul",   "div",   "mod",   "neg",    "abs",    "fabs",  "sqrt", "inversesqrt",
                "pow",    "min",  "max",  "floor", "ceil",  "round", "trunc",  "sign",   "clamp", "mix",  "color",
                "point",  "vector", "normal", "float", "int", "compref", "dot", "cross", "length", "normalize",
                "sin",    "cos",  "tan",  "exp",   "exp2",  "log",   "log2",   "eq",     "neq",   "lt",   "gt",
                "le",     "ge",   "and",  "or",    "not",   "bitand", "bitor", "xor",    "shl",   "shr",  "compl",
                "step",   "smoothstep", "select" };
        const std::vector<Opcode>& ops = L->m.ops;
        int writer = -1, nwrites = 0;
        for (size_t i = 0; i < ops.size(); ++i)
            for (size_t a = 0; a < ops[i].args.size(); ++a)
                if (ops[i].args[a] == si && ops[i].writes((int)a)) {
                    ++nwrites;
                    writer = (int)i;
                }
        if (nwrites != 1 || !foldable.count(ops[writer].name) || !ops[writer].jumps.empty())
            return false;
        for (size_t j = 0; j < ops.size(); ++j) {   // inside the body of an if / loop: not folded
            int last = -1;
            for (int t : ops[j].jumps)
                last = std::max(last, t);
            if ((int)j < writer && writer < last)
                return false;
        }
        for (size_t a = 0; a < ops[writer].args.size(); ++a)
            if (ops[writer].reads((int)a) && !ops[writer].writes((int)a) && !folds_to_constant(ops[writer].args[a], depth + 1))
                return false;
        return true;
    }

    // default-value initialiser expressions, one per array element
    std::vector<std::string> initvals(const Symbol& s)
    {
        int n   = s.type.arraylen ? s.type.arraylen : 1;
        int per = s.type.ncomp();
        std::vector<std::string> out;
        for (int e = 0; e < n; ++e) {
            auto fv = [&](int c) {
                size_t k = (size_t)e * per + c;
                return k < s.fvals.size() ? s.fvals[k] : 0.0f;
            };
            switch (s.type.base) {
            case Base::Int: out.push_back(std::to_string((size_t)e < s.ivals.size() ? s.ivals[e] : 0)); break;
            case Base::String: out.push_back(std::to_string(g.intern((size_t)e < s.svals.size() ? s.svals[e] : ""))); break;
            case Base::Float: out.push_back(s.has_derivs ? "mkd(" + cfloat(fv(0)) + ")" : cfloat(fv(0))); break;
            default:
                if (s.type.is_triple()) {
                    std::string v = "mkv(" + cfloat(fv(0)) + ", " + cfloat(fv(1)) + ", " + cfloat(fv(2)) + ")";
                    out.push_back(s.has_derivs ? "mkdv(" + v + ")" : v);
                } else if (s.type.base == Base::Matrix)
                    out.push_back(m44_literal(s.fvals, (size_t)e * 16));
                else if (s.type.base == Base::Closure)
                    out.push_back("0");   // closure-typed parameters default to the null closure (pool offset 0)
                else
                    unsupported("default value of '" + s.name + "'");
            }
        }
        return out;
    }

    // component c of symbol as float (derivs=false) or natural scalar
    std::string comp(int si, int c, bool derivs)
    {
        const Symbol& s = S(si);
        if (s.is_const()) {
            if (s.type.base == Base::Int)
                return cfloat((float)(s.ivals.empty() ? 0 : s.ivals[0]));
            if (s.type.base == Base::Float)
                return cfloat(s.fvals.empty() ? 0.0f : s.fvals[0]);
            if (s.type.is_triple())
                return cfloat((size_t)c < s.fvals.size() ? s.fvals[c] : 0.0f);
        }
        std::string e = "getc(" + R(si) + ", " + std::to_string(c) + ")";
        if (!derivs && s.has_derivs)
            e = "nd(" + e + ")";
        return e;
    }

    // Colour-space name -> OSLD_CS_* code (osl_b200_color.cuh).  ctor: the name feeds
    // ColorSystem::to_rgb (no "linear"/"sRGB" clauses) instead of transformc.
    std::string space_code(int si, bool ctor)
    {
        struct E {
            const char* name;
            const char* code;
            bool ctor_ok;
        };
        static const E table[] = { { "RGB", "OSLD_CS_RGB", true },   { "rgb", "OSLD_CS_RGB", true },
                                   { "linear", "OSLD_CS_RGB", false }, { "hsv", "OSLD_CS_HSV", true },
                                   { "hsl", "OSLD_CS_HSL", true },   { "YIQ", "OSLD_CS_YIQ", true },
                                   { "XYZ", "OSLD_CS_XYZ", true },   { "xyY", "OSLD_CS_XYY", true },
                                   { "sRGB", "OSLD_CS_SRGB", false } };
        const Symbol& s = S(si);
        if (s.const_value()) {
            std::string v = s.svals.empty() ? "" : s.svals[0];
            if (v == g.colorspace)
                return "OSLD_CS_RGB";
            for (const E& e : table)
                if (v == e.name && (e.ctor_ok || !ctor))
                    return e.code;
            return "OSLD_CS_UNKNOWN";
        }
        // run-time name: compare interned ids
        std::string e = R(si), r = "(";
        r += "(" + e + ") == " + std::to_string(g.intern(g.colorspace)) + " ? OSLD_CS_RGB : ";
        for (const E& t : table)
            if (t.ctor_ok || !ctor)
                r += "(" + e + ") == " + std::to_string(g.intern(t.name)) + " ? " + t.code + " : ";
        return r + "OSLD_CS_UNKNOWN)";
    }

    void emit_matrix_op(const Opcode& op);
    void emit_printf(const Opcode& op);
    bool journal_ok = false;  // grid kernels write printf records to the launch's journal
    bool material_mode = false;  // group compiled as a renderer material: no launch block, no userdata
    bool uses_closures = false;  // a grid kernel then carries a per-point closure pool
    void gen_layer(int layer);
    void emit_block(int b, int e, const Ctx* ctx);
    void useparams(const Opcode& op);
    void emit_op(const Opcode& op);
    void op_percomp(const Opcode& op);
    void op_cmp(const Opcode& op);
    void op_noise(const Opcode& op, bool periodic);
    void op_noise_named(const Opcode& op, bool periodic, const std::string& runtime_name);
    void emit_copy(int dl, const Symbol& d, int sl, const Symbol& s);
    void emit_userdata_load(const UserData& ud, const Symbol& s, const std::string& r, const std::string& flag);
    // messages (opmessage.cpp): one group-data slot per constant message name, typed by the
    // first setmessage of that name in layer order
    struct MsgSlot {
        int id;
        Base base;
        int ncomp;
    };
    std::map<std::string, MsgSlot> messages;
    void scan_messages();
    std::string message_fields();
};

// value of userdata entry `ud` for this point -> r; `flag` says whether the point has it
void
Gen::emit_userdata_load(const UserData& ud, const Symbol& s, const std::string& r, const std::string& flag)
{
    w("bool " + flag + " = L.userdata_base != nullptr;");
    if (ud.valid_offset >= 0)
        w("if (" + flag + ") " + flag + " = __ldg((const int*)((const char*)L.userdata_base + "
          + std::to_string(ud.valid_offset) + "ll + " + std::to_string(ud.valid_stride) + "ll * (long long)sg.shadeindex)) != 0;");
    w("if (" + flag + ") {");
    ++ind;
    w("const float* p_ = (const float*)((const char*)L.userdata_base + " + std::to_string(ud.offset) + "ll + "
      + std::to_string(ud.stride) + "ll * (long long)sg.shadeindex);");
    auto ld = [&](int i) { return "__ldg(p_ + " + std::to_string(i) + ")"; };
    if (ud.is_int)
        w(r + " = __float_as_int(" + ld(0) + ");");
    else if (ud.ncomp == 1) {
        if (s.has_derivs)
            w(r + " = mkd(" + ld(0) + ", " + (ud.derivs ? ld(1) : std::string("0.0f")) + ", "
              + (ud.derivs ? ld(2) : std::string("0.0f")) + ");");
        else
            w(r + " = " + ld(0) + ";");
    } else {
        auto v3 = [&](int b) { return "mkv(" + ld(b) + ", " + ld(b + 1) + ", " + ld(b + 2) + ")"; };
        if (s.has_derivs)
            w(r + " = mkdv(" + v3(0) + ", " + (ud.derivs ? v3(3) : std::string("mkv(0.0f)")) + ", "
              + (ud.derivs ? v3(6) : std::string("mkv(0.0f)")) + ");");
        else
            w(r + " = " + v3(0) + ";");
    }
    --ind;
    w("}");
}

void
Gen::scan_messages()
{
    messages.clear();
    for (Layer& l : g.layers) {
        if (l.unused)
            continue;
        for (const Opcode& op : l.m.ops) {
            if (op.name != "setmessage" || op.args.size() != 2)
                continue;
            const Symbol& nm = l.m.syms[op.args[0]];
            const Symbol& v  = l.m.syms[op.args[1]];
            if (nm.type.base != Base::String || !nm.const_value() || nm.svals.empty() || v.type.arraylen)
                continue;
            if (!(v.type.base == Base::Int || v.type.base == Base::Float || v.type.is_triple()))
                continue;   // closures, strings, matrices and arrays as messages are not built
            if (!messages.count(nm.svals[0]) && messages.size() < 31)
                messages[nm.svals[0]] = MsgSlot { (int)messages.size(), v.type.base, v.type.ncomp() };
        }
    }
}
std::string
Gen::message_fields()
{
    std::string s;
    if (messages.empty())
        return s;
    s += "    unsigned msgset;\n";
    for (auto& kv : messages)
        s += std::string("    ") + (kv.second.base == Base::Int ? "int" : (kv.second.ncomp == 3 ? "V3" : "float")) + " M"
             + std::to_string(kv.second.id) + ";\n";
    return s;
}

void
Gen::emit_copy(int dl, const Symbol& d, int sl, const Symbol& s)
{
    std::string de = ref(dl, d), se = ref(sl, s);
    if (d.type.arraylen) {
        for (int i = 0; i < d.type.arraylen; ++i)
            w("assign(" + de + "[" + std::to_string(i) + "], " + se + "[" + std::to_string(i) + "]);");
    } else if (d.type.base == Base::String || d.type.base == Base::Int) {
        w(de + " = " + se + ";");
    } else
        w("assign(" + de + ", " + se + ");");
}

void
Gen::useparams(const Opcode& op)
{
    // lazy upstream evaluation (llvm_gen.cpp:133-283): run the producing layer
    // once, the first time a connected parameter is read
    for (size_t a = 0; a < op.args.size(); ++a) {
        const Symbol& s = S(op.args[a]);
        if (op.reads((int)a) && s.conn_layer >= 0 && !ensured.count(s.conn_layer)) {
            ensured.insert(s.conn_layer);
            std::string k = std::to_string(s.conn_layer);
            w("if (!(gd.ran & " + std::to_string(1u << s.conn_layer) + "u)) layer_" + k + "(sg, gd, L);");
        }
    }
}

void
Gen::emit_block(int b, int e, const Ctx* ctx)
{
    const std::vector<Opcode>& ops = L->m.ops;
    int i                          = b;
    while (i < e) {
        const Opcode& op     = ops[i];
        const std::string& n = op.name;
        if (n == "if") {
            if (op.jumps.size() < 2)
                unsupported("malformed 'if'");
            useparams(op);
            w("if (" + R(op.args[0]) + ") {");
            ++ind;
            std::set<int> saved = ensured;
            emit_block(i + 1, op.jumps[0], ctx);
            --ind;
            ensured = saved;
            if (op.jumps[1] > op.jumps[0]) {
                w("} else {");
                ++ind;
                emit_block(op.jumps[0], op.jumps[1], ctx);
                --ind;
                ensured = saved;
            }
            w("}");
            i = op.jumps[1];
        } else if (n == "for" || n == "while" || n == "dowhile") {
            if (op.jumps.size() < 4)
                unsupported("malformed loop");
            int cl = op.jumps[0], bl = op.jumps[1], il = op.jumps[2], dl = op.jumps[3];
            int lab = ++label;
            Ctx c2  = ctx ? *ctx : Ctx();
            c2.brk  = "brk_" + std::to_string(lab);
            c2.cont = "cont_" + std::to_string(lab);
            emit_block(i + 1, cl, ctx);
            std::set<int> saved = ensured;
            std::string cond    = R(op.args[0]);
            w("for (;;) {");
            ++ind;
            ++loop_depth;
            if (n == "dowhile") {
                emit_block(bl, il, &c2);
                w(c2.cont + ":;");
                emit_block(cl, bl, &c2);
                w("if (!(" + cond + ")) break;");
                emit_block(il, dl, &c2);
            } else {
                emit_block(cl, bl, &c2);
                w("if (!(" + cond + ")) break;");
                emit_block(bl, il, &c2);
                w(c2.cont + ":;");
                emit_block(il, dl, &c2);
            }
            --loop_depth;
            --ind;
            w("}");
            w(c2.brk + ":;");
            ensured = saved;
            i       = dl;
        } else if (n == "functioncall") {
            int lab = ++label;
            Ctx c2  = ctx ? *ctx : Ctx();
            c2.ret  = "ret_" + std::to_string(lab);
            w("{");
            ++ind;
            std::set<int> saved = ensured;
            emit_block(i + 1, op.jumps[0], &c2);
            ensured = saved;
            --ind;
            w("}");
            w(c2.ret + ":;");
            i = op.jumps[0];
        } else if (n == "break") {
            w("goto " + ctx->brk + ";");
            ++i;
        } else if (n == "continue") {
            w("goto " + ctx->cont + ";");
            ++i;
        } else if (n == "return") {
            w("goto " + ((ctx && !ctx->ret.empty()) ? ctx->ret : std::string("layer_end")) + ";");
            ++i;
        } else if (n == "exit") {
            w("goto layer_end;");
            ++i;
        } else if (n == "nop" || n == "end" || n == "useparam") {
            ++i;
        } else {
            useparams(op);
            w("{");
            ++ind;
            emit_op(op);
            --ind;
            w("}");
            ++i;
        }
    }
}

void
Gen::op_percomp(const Opcode& op)
{
    const Symbol& d = S(op.args[0]);
    if (d.type.base == Base::Closure) {
        // llvm_gen_add / llvm_gen_mul closure branches -> osl_add_closure_closure,
        // osl_mul_closure_{float,color} (opclosure.cpp:18-76)
        if (op.args.size() != 3)
            unsupported("closure op '" + op.name + "'");
        uses_closures = true;
        int a = op.args[1], b = op.args[2];
        g.closure_in_loop |= loop_depth > 0;
        if (op.name == "add") {
            g.pool_words_bound += 3;
            g.closure_adds += 1;
            w(R(op.args[0]) + " = clos_add(*sg.pool, " + R(a) + ", " + R(b) + ");");
            return;
        }
        if (op.name != "mul")
            unsupported("closure op '" + op.name + "'");
        if (S(a).type.base != Base::Closure)
            std::swap(a, b);
        std::string wt = R(b);
        if (S(b).has_derivs)
            wt = "nd(" + wt + ")";
        if (S(b).type.base == Base::Int)
            wt = "(float)" + wt;
        g.pool_words_bound += 5;
        w(R(op.args[0]) + " = clos_mul(*sg.pool, " + R(a) + ", " + wt + ");");
        return;
    }
    if (d.type.base == Base::Matrix)
        unsupported("op '" + op.name + "' on matrices");
    bool isint = d.type.base == Base::Int;
    bool dv    = false;
    if (d.has_derivs)
        for (size_t a = 1; a < op.args.size(); ++a)
            dv |= S(op.args[a]).has_derivs;
    std::string fn = "o_" + op.name;
    if (op.name == "div") {
        const Symbol& b = S(op.args[2]);
        bool nz         = b.is_const();
        for (float f : b.fvals)
            nz &= (f != 0.0f);
        for (int v : b.ivals)
            nz &= (v != 0);
        if (nz)
            fn = "o_divc";
    }
    for (int c = 0; c < d.type.ncomp(); ++c) {
        std::string args;
        for (size_t a = 1; a < op.args.size(); ++a) {
            if (a > 1)
                args += ", ";
            args += isint ? R(op.args[a]) : comp(op.args[a], c, dv);
        }
        w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", " + fn + "(" + args + "));");
    }
}

void
Gen::op_cmp(const Opcode& op)
{
    const Symbol &a = S(op.args[1]), &b = S(op.args[2]);
    const char* cop = CMP.at(op.name);
    std::string e;
    if (a.type.base == Base::String || (a.type.base == Base::Int && b.type.base == Base::Int)) {
        e = "(" + R(op.args[1]) + " " + cop + " " + R(op.args[2]) + ")";
    } else if (a.type.base == Base::Closure) {
        e = "(" + R(op.args[1]) + " " + cop + " 0)";   // only closure == 0 / != 0 exist
    } else {
        int nc = std::max(a.type.ncomp(), b.type.ncomp());
        for (int c = 0; c < nc; ++c) {
            if (c)
                e += op.name == "neq" ? " || " : " && ";
            e += "(" + comp(op.args[1], a.type.ncomp() > 1 ? c : 0, false) + " " + cop + " "
                 + comp(op.args[2], b.type.ncomp() > 1 ? c : 0, false) + ")";
        }
    }
    w(R(op.args[0]) + " = (" + e + ") ? 1 : 0;");
}

void
Gen::op_noise(const Opcode& op, bool periodic)
{
    // A name that is not known at compile time: the reference calls osl_genericnoise /
    // osl_genericpnoise, which compare the name at run time (GenericNoise / GenericPNoise,
    // opnoise.cpp:704-900).  Here every name the reference accepts gets the code the
    // compile-time path would emit for it, selected by the interned name id: a warp whose
    // lanes agree on the name (the usual case: a string parameter) runs exactly one branch.
    if (op.args.size() > 1 && S(op.args[1]).type.base == Base::String && !S(op.args[1]).const_value()) {
        static const char* const names[][2]  = { { "perlin", "snoise" },   { "uperlin", "noise" }, { "simplex", "simplexnoise" },
                                                 { "usimplex", "usimplexnoise" }, { "cell", nullptr }, { "hash", nullptr },
                                                 { "gabor", nullptr } };
        w("const int nm_ = " + R(op.args[1]) + ";");
        bool first = true;
        for (auto& pr : names) {
            if (periodic && (std::string(pr[0]) == "simplex" || std::string(pr[0]) == "usimplex"))
                continue;   // GenericPNoise has no simplex branch
            std::string cond = "nm_ == " + std::to_string(g.intern(pr[0]));
            if (pr[1])
                cond += " || nm_ == " + std::to_string(g.intern(pr[1]));
            w(std::string(first ? "if (" : "} else if (") + cond + ") {");
            first = false;
            ++ind;
            op_noise_named(op, periodic, pr[0]);
            --ind;
        }
        // unknown name: the reference reports "Unknown noise type" and leaves the result alone
        w("}");
        return;
    }
    op_noise_named(op, periodic, "");
}

void
Gen::op_noise_named(const Opcode& op, bool periodic, const std::string& runtime_name)
{
    // llvm_gen_noise (llvm_gen.cpp:3117-3299): the noise name is resolved at
    // code-generation time; the float vs Dual2 entry is chosen from has_derivs
    std::vector<int> rest(op.args.begin() + 1, op.args.end());
    const Symbol& d  = S(op.args[0]);
    std::string name = op.name;
    if (!rest.empty() && S(rest[0]).type.base == Base::String) {
        const Symbol& ns = S(rest[0]);
        if (!runtime_name.empty())
            name = runtime_name;
        else
            name = ns.svals.empty() ? "" : ns.svals[0];
        rest.erase(rest.begin());
    }
    std::vector<int> coords;
    for (int a : rest) {
        if (S(a).type.base == Base::String)
            break;  // optional token/value pairs (only gabor consumes them)
        coords.push_back(a);
    }
    int kind = noise_kind(name, periodic);
    if (kind < 0)
        unsupported("noise type \"" + name + "\"");
    std::vector<int> opts(rest.begin() + coords.size(), rest.end());
    std::vector<int> pers;
    if (periodic) {
        size_t half = coords.size() / 2;
        pers.assign(coords.begin() + half, coords.end());
        coords.resize(half);
    }
    std::vector<std::pair<int, int>> ins;
    for (int a : coords)
        for (int c = 0; c < S(a).type.ncomp(); ++c)
            ins.push_back({ a, c });
    int dim = (int)ins.size(), nc = d.type.ncomp();
    if (dim < 1 || dim > 4)
        unsupported("noise with " + std::to_string(dim) + " input dimensions");
    if (kind == 6) {
        // Gabor branch (llvm_gen.cpp:3196-3203): derivatives are always taken
        // (zero where the coordinate has none); options travel in NoiseParams
        // (llvm_gen_noise_options, llvm_gen.cpp:3057-3103); 1-D / 2-D slice the
        // 3-D noise and the 4-D forms ignore time (opnoise.cpp:484-632).
        w("NoiseParams opt_ = noise_params_default();");
        for (size_t i = 0; i + 1 < opts.size(); i += 2) {
            const Symbol& nm = S(opts[i]);
            const Symbol& v  = S(opts[i + 1]);
            if (nm.type.base != Base::String || !nm.const_value())
                unsupported("noise option with a name that is not known at compile time");
            std::string key = nm.svals.empty() ? "" : nm.svals[0];
            std::string e   = R(opts[i + 1]);
            if (v.has_derivs)
                e = "nd(" + e + ")";
            bool scalar = !v.type.is_triple() && v.type.arraylen == 0;
            if (key.empty())
                continue;
            if ((key == "anisotropic" || key == "do_filter") && v.type.base == Base::Int && scalar)
                w("opt_." + key + " = " + e + ";");
            else if (key == "direction" && v.type.is_triple())
                w("assign(opt_.direction, " + e + ");");
            else if ((key == "bandwidth" || key == "impulses") && scalar
                     && (v.type.base == Base::Float || v.type.base == Base::Int))
                w("opt_." + key + " = (float)(" + e + ");");
            else
                g.warnings.push_back("Unknown noise optional argument: \"" + key + "\" in layer '" + L->layername + "'");
        }
        std::string P[3] = { "mkd(0.0f)", "mkd(0.0f)", "mkd(0.0f)" };
        for (int k = 0; k < dim && k < 3; ++k) {
            P[k] = "as_dual(" + comp(ins[k].first, ins[k].second, true) + ")";
        }
        std::string per = "nullptr";
        if (periodic) {
            std::string pc[3] = { "0.0f", "0.0f", "0.0f" };
            int np = 0;
            for (int a : pers)
                for (int c = 0; c < S(a).type.ncomp() && np < 3; ++c, ++np)
                    pc[np] = comp(a, c, false);
            w("V3 per_ = mkv(" + pc[0] + ", " + pc[1] + ", " + pc[2] + ");");
            per = "&per_";
        }
        w("Df out_[3];");
        w("gabor_noise<" + std::to_string(nc) + ">(out_, " + P[0] + ", " + P[1] + ", " + P[2] + ", " + per + ", opt_);");
        for (int c = 0; c < nc; ++c)
            w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", out_[" + std::to_string(c) + "]);");
        return;
    }
    bool hashy   = kind == 2 || kind == 3;
    bool simplex = kind == 4 || kind == 5;
    if (simplex && periodic)
        unsupported("periodic simplex noise does not exist in OSL");
    bool dv = false;
    if (!hashy && d.has_derivs)
        for (int a : coords)
            dv |= S(a).has_derivs;
    std::string T = dv ? "Df" : "float";
    std::string in = T + " in_[4] = {";
    for (int k = 0; k < dim; ++k)
        in += (k ? ", " : "") + comp(ins[k].first, ins[k].second, dv);
    w(in + "};");
    w(T + " out_[3];");
    std::string sd = std::to_string(dim), sn = std::to_string(nc);
    if (periodic) {
        std::string pin;
        int np = 0;
        for (int a : pers)
            for (int c = 0; c < S(a).type.ncomp(); ++c, ++np)
                pin += (np ? ", " : "") + comp(a, c, false);
        if (hashy) {
            w("float per_[4] = {" + pin + "};");
            w("for (int k_ = 0; k_ < " + sd + "; ++k_) in_[k_] = pwrap(in_[k_], per_[k_]);");
            w(std::string("ihnoise<") + (kind == 2 ? "true" : "false") + ", " + sd + ", " + sn + ">(out_, in_);");
        } else {
            w("float perf_[4] = {" + pin + "};");
            w("int per_[4];");
            w("for (int k_ = 0; k_ < " + sd + "; ++k_) per_[k_] = iperiod(perf_[k_]);");
            w("perlin<" + T + ", " + sd + ", " + sn + ", " + (kind == 1 ? "true" : "false") + ", true>(out_, in_, per_);");
        }
    } else if (hashy) {
        w(std::string("ihnoise<") + (kind == 2 ? "true" : "false") + ", " + sd + ", " + sn + ">(out_, in_);");
    } else if (simplex) {
        w("simplex<" + sd + ", " + sn + ", " + (kind == 5 ? "true" : "false") + ">(out_, in_);");
    } else {
        w("perlin<" + T + ", " + sd + ", " + sn + ", " + (kind == 1 ? "true" : "false") + ", false>(out_, in_, nullptr);");
    }
    for (int c = 0; c < nc; ++c)
        w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", out_[" + std::to_string(c) + "]);");
}

// printf -> one journal record: [words][shadeindex][sequence][format id][argument words].
// The reference's device path does the same through its journal (src/include/OSL/journal.h,
// rs_printfmt) and formats on the host; llvm_gen_printf (llvm_gen.cpp:350-700) fixes the
// per-type conversion rules that the host decoder (osl_b200_cabi.cu) applies.
void
Gen::emit_printf(const Opcode& op)
{
    const Symbol& f = S(op.args[0]);
    if (f.type.base != Base::String || !f.const_value())
        unsupported("printf with a format that is not known at compile time");
    JournalFormat jf;
    jf.fmt        = f.svals.empty() ? "" : f.svals[0];
    jf.kind       = op.name == "error" ? 1 : (op.name == "warning" ? 2 : 0);
    jf.shadername = L->m.shadername;
    // The format is applied on the host with the C library: check every conversion against its
    // argument here, so that a malformed or hand-written .oso cannot make the formatter read a
    // string pointer out of a float or write through %n (the reference's llvm_gen_printf raises
    // "Mismatch between format string and arguments" in the same spot).
    {
        size_t ai = 1;
        for (size_t i = 0; i < jf.fmt.size();) {
            if (jf.fmt[i] != '%') {
                ++i;
                continue;
            }
            if (i + 1 < jf.fmt.size() && jf.fmt[i + 1] == '%') {
                i += 2;
                continue;
            }
            size_t j = i + 1;
            while (j < jf.fmt.size() && strchr("-+ #0123456789.", jf.fmt[j]))
                ++j;
            const char conv = j < jf.fmt.size() ? jf.fmt[j] : 0;
            if (!conv || !strchr("diouxXcfeEgGsm", conv))
                unsupported("printf format \"" + jf.fmt + "\": unsupported conversion");
            if (ai >= op.args.size())
                unsupported("Mismatch between format string and arguments");
            const Base b = S(op.args[ai++]).type.base;
            const bool ok = b == Base::String ? conv == 's'
                                              : (b == Base::Int ? strchr("diouxXcfeEgG", conv) != nullptr
                                                                : strchr("feEgGdixXm", conv) != nullptr);
            if (!ok)
                unsupported("Mismatch between format string and arguments");
            i = j + 1;
        }
        if (ai != op.args.size())
            unsupported("Mismatch between format string and arguments");
    }
    std::vector<std::string> words;
    for (size_t a = 1; a < op.args.size(); ++a) {
        const Symbol& s = S(op.args[a]);
        JournalArg ja;
        ja.base     = s.type.base;
        ja.ncomp    = s.type.ncomp();
        ja.arraylen = s.type.arraylen;
        if (s.type.base == Base::Closure || s.type.base == Base::Void)
            unsupported("printf of a closure");
        jf.args.push_back(ja);
        int n = s.type.arraylen ? s.type.arraylen : 1;
        for (int e = 0; e < n; ++e) {
            std::string el = R(op.args[a]) + (s.type.arraylen ? "[" + std::to_string(e) + "]" : "");
            if (s.type.base == Base::Int || s.type.base == Base::String)
                words.push_back("(unsigned)(" + el + ")");
            else if (s.type.base == Base::Matrix)
                for (int c = 0; c < 16; ++c)
                    words.push_back("__float_as_uint(" + el + ".x[" + std::to_string(c / 4) + "][" + std::to_string(c % 4) + "])");
            else
                for (int c = 0; c < ja.ncomp; ++c)
                    words.push_back("__float_as_uint(nd(getc(" + el + ", " + std::to_string(c) + ")))");
        }
    }
    int id = (int)g.jformats.size();
    g.jformats.push_back(jf);
    w("{ unsigned* j_ = jr_reserve(L, sg, " + std::to_string(id) + "u, " + std::to_string(words.size()) + "u);");
    w("  if (j_) {");
    for (size_t k = 0; k < words.size(); ++k)
        w("    j_[" + std::to_string(k) + "] = " + words[k] + ";");
    w("  } }");
}

// Matrix shadeops (opmatrix.cpp; llvm_gen_matrix, llvm_gen_getmatrix, llvm_gen_transform,
// llvm_gen_mxcompref/assign).  Space names must be compile-time constants; their matrices
// come from the launch block (see space_matrix).
void
Gen::emit_matrix_op(const Opcode& op)
{
    const std::string& n = op.name;
    auto A               = [&](int i) -> Symbol& { return S(op.args[i]); };
    auto fl              = [&](int i) { return comp(op.args[i], 0, false); };
    auto is_m            = [&](int i) { return A(i).type.base == Base::Matrix; };
    if ((n == "printf" || n == "error" || n == "warning") && journal_ok) {
        emit_printf(op);
        return;
    }
    const std::string d  = R(op.args[0]);
    if (n == "printf" || n == "error" || n == "warning" || n == "fprintf") {
        std::string msg = "op '" + n + "' ignored on device (no journal yet) in layer '" + L->layername + "'";
        bool dup = false;
        for (auto& s : g.warnings)
            dup |= (s == msg);
        if (!dup)
            g.warnings.push_back(msg);
    } else if (n == "assign") {
        w(is_m(1) ? d + " = " + R(op.args[1]) + ";" : d + " = m44_diag(" + fl(1) + ");");
    } else if (n == "neg") {
        w(d + " = -" + R(op.args[1]) + ";");
    } else if (n == "mul") {
        if (is_m(1) && is_m(2))
            w(d + " = " + R(op.args[1]) + " * " + R(op.args[2]) + ";");
        else
            w(d + " = " + R(op.args[is_m(1) ? 1 : 2]) + " * " + fl(is_m(1) ? 2 : 1) + ";");
    } else if (n == "div") {
        if (is_m(1) && is_m(2))
            w(d + " = " + R(op.args[1]) + " * m44_inverse(" + R(op.args[2]) + ");");
        else if (is_m(1))
            w(d + " = " + R(op.args[1]) + " * (1.0f / " + fl(2) + ");");
        else if (is_m(2))
            w(d + " = " + fl(1) + " * m44_inverse(" + R(op.args[2]) + ");");
        else
            w("{ float b_ = " + fl(2) + "; " + d + " = m44_diag(b_ == 0 ? 0.0f : (" + fl(1) + " / b_)); }");
    } else if (n == "eq" || n == "neq") {
        std::string a = is_m(1) ? R(op.args[1]) : "m44_diag(" + fl(1) + ")";
        std::string b = is_m(2) ? R(op.args[2]) : "m44_diag(" + fl(2) + ")";
        w(d + " = (" + (n == "neq" ? "!" : "") + "(" + a + " == " + b + ")) ? 1 : 0;");
    } else if (n == "transpose") {
        w(d + " = m44_transposed(" + R(op.args[1]) + ");");
    } else if (n == "determinant") {
        w("assign(" + d + ", m44_determinant(" + R(op.args[1]) + "));");
    } else if (n == "mxcompref") {
        w("assign(" + d + ", " + R(op.args[1]) + ".x[" + R(op.args[2]) + " & 3][" + R(op.args[3]) + " & 3]);");
    } else if (n == "mxcompassign") {
        w(d + ".x[" + R(op.args[1]) + " & 3][" + R(op.args[2]) + " & 3] = " + fl(3) + ";");
    } else if (n == "matrix") {
        // llvm_gen_matrix (llvm_gen.cpp:2295-2370)
        size_t nargs     = op.args.size();
        bool using_space = (nargs == 3 || nargs == 18) && A(1).type.base == Base::String;
        bool two_spaces  = nargs == 3 && A(2).type.base == Base::String;
        if (two_spaces) {  // osl_get_from_to_matrix
            auto mf = space_matrix(op.args[1], false), mt = space_matrix(op.args[2], true);
            w(d + " = " + mf.first + " * " + mt.first + ";");
            return;
        }
        size_t v0 = 1 + (using_space ? 1 : 0), nv = nargs - v0;
        if (nv == 1)
            w(d + " = m44_diag(" + fl((int)v0) + ");");
        else if (nv == 16) {
            std::string e = "m44_make(";
            for (size_t k = 0; k < 16; ++k)
                e += (k ? ", " : "") + fl((int)(v0 + k));
            w(d + " = " + e + ");");
        } else
            unsupported("matrix constructor with " + std::to_string(nv) + " values");
        if (using_space) {  // osl_prepend_matrix_from: only when the space is known
            auto mf = space_matrix(op.args[1], false);
            w("if (" + mf.second + ") " + d + " = " + mf.first + " * " + d + ";");
        }
    } else if (n == "getmatrix") {
        auto mf = space_matrix(op.args[1], false), mt = space_matrix(op.args[2], true);
        w(R(op.args[3]) + " = " + mf.first + " * " + mt.first + ";");
        w(d + " = (" + mf.second + " & " + mt.second + ") ? 1 : 0;");
    } else if (n == "transform" || n == "transformv" || n == "transformn") {
        // llvm_gen_transform (llvm_gen.cpp:2376-2466) -> osl_transform{,v,n}_* / osl_transform_triple
        const char* vt = n == "transform" ? "0" : (n == "transformv" ? "1" : "2");
        int pi         = (int)op.args.size() - 1;
        const Symbol& p = A(pi);
        bool dv        = A(0).has_derivs && p.has_derivs;
        std::string pe = R(op.args[pi]);
        if (p.has_derivs && !dv)
            pe = "nd(" + pe + ")";
        if (op.args.size() == 3 && is_m(1)) {
            w("assign(" + d + ", m44_transform(" + R(op.args[1]) + ", " + pe + ", " + vt + "));");
            return;
        }
        int fi = op.args.size() == 3 ? -1 : 1, ti = op.args.size() == 3 ? 1 : 2;
        {
            // same space on both sides (after the "world" synonym): an identity, just copy
            auto norm = [&](int i) -> std::string {
                if (i < 0)
                    return "common";
                const Symbol& s = S(op.args[i]);
                if (s.type.base != Base::String || !s.const_value())
                    unsupported("coordinate-system name that is not known at compile time");
                std::string v = s.svals.empty() ? "" : s.svals[0];
                return v == g.commonspace_synonym ? "common" : v;
            };
            if (norm(fi) == norm(ti)) {
                w("assign(" + d + ", " + pe + ");");
                return;
            }
        }
        bool from_common = fi < 0 || space_is(op.args[fi], "common");
        bool to_common   = space_is(op.args[ti], "common");
        std::string M, ok;
        if (from_common) {
            auto mt = space_matrix(op.args[ti], true);
            M = mt.first; ok = mt.second;
        } else if (to_common) {
            auto mf = space_matrix(op.args[fi], false);
            M = mf.first; ok = mf.second;
        } else {
            auto mf = space_matrix(op.args[fi], false), mt = space_matrix(op.args[ti], true);
            M  = mf.first + " * " + mt.first;
            ok = "(" + mf.second + " & " + mt.second + ")";
        }
        // unknown space: the value passes through unchanged (osl_transform_triple)
        w("if (" + ok + ") assign(" + d + ", m44_transform(" + M + ", " + pe + ", " + vt + ")); else assign(" + d + ", "
          + pe + ");");
    } else {
        unsupported("op '" + n + "' on a matrix");
    }
}

void
Gen::emit_op(const Opcode& op)
{
    const std::string& n = op.name;
    auto A               = [&](int i) -> Symbol& { return S(op.args[i]); };
    auto need            = [&](size_t k) {
        if (op.args.size() < k)
            unsupported("op '" + n + "' has too few arguments");
    };
    bool any_matrix = false;
    for (int a : op.args)
        any_matrix |= S(a).type.base == Base::Matrix;
    if (any_matrix || n == "transform" || n == "transformv" || n == "transformn" || n == "getmatrix"
        || n == "matrix") {
        emit_matrix_op(op);
    } else if (n == "mod") {
        need(3);
        w(R(op.args[0]) + " = o_mod(" + R(op.args[1]) + ", " + R(op.args[2]) + ");");
    } else if (n == "compl") {
        need(2);
        w(R(op.args[0]) + " = ~" + R(op.args[1]) + ";");
    } else if (UNARY.count(n) || BINARY.count(n) || TERNARY.count(n)) {
        op_percomp(op);
    } else if (CMP.count(n)) {
        need(3);
        op_cmp(op);
    } else if (INTBIN.count(n)) {
        need(3);
        w(R(op.args[0]) + " = " + R(op.args[1]) + " " + INTBIN.at(n) + " " + R(op.args[2]) + ";");
    } else if (n == "assign") {
        need(2);
        const Symbol &d = A(0), &s = A(1);
        if (d.type.arraylen) {
            // array = array copies as many elements as both have (llvm_assign_impl; testsuite/array-copy)
            const int ncopy = s.type.arraylen ? std::min(d.type.arraylen, s.type.arraylen) : d.type.arraylen;
            for (int i = 0; i < ncopy; ++i)
                w("assign(" + R(op.args[0]) + "[" + std::to_string(i) + "], " + R(op.args[1]) + "["
                  + std::to_string(i) + "]);");
        } else if (d.type.base == Base::String) {
            w(R(op.args[0]) + " = " + R(op.args[1]) + ";");
        } else if (d.type.base == Base::Closure) {
            w(R(op.args[0]) + " = " + (s.type.base == Base::Int ? std::string("0") : R(op.args[1])) + ";");
        } else if (d.type.base == Base::Matrix) {
            unsupported("assignment of matrix values");
        } else {
            (void)s;
            w("assign(" + R(op.args[0]) + ", " + R(op.args[1]) + ");");
        }
    } else if (n == "color" && op.args.size() == 5) {
        // llvm_gen_construct_color (llvm_gen.cpp:1826-1865): osl_prepend_color_from on the
        // value; the derivatives of the result are zeroed.  ColorSystem::to_rgb has no
        // "linear"/"sRGB" clause, so those names pass through here (opcolor.cpp:287-305).
        g.uses_colorsystem = true;
        w("assign(" + R(op.args[0]) + ", color_transformc(osl_cs_, " + space_code(op.args[1], true) + ", OSLD_CS_RGB, mkv("
          + comp(op.args[2], 0, false) + ", " + comp(op.args[3], 0, false) + ", " + comp(op.args[4], 0, false) + ")));");
    } else if (n == "texture") {
        // llvm_gen_texture (llvm_gen.cpp:2715-2830): result, filename, s, t, [dsdx, dtdx, dsdy,
        // dtdy,] then ("name", value) option pairs (llvm_gen_texture_options, :2480-2713).
        if (op.args.size() < 4)
            unsupported("texture with fewer than 4 arguments");
        const Symbol& fn = A(1);
        if (!fn.const_value() || fn.svals.empty())
            unsupported("texture() with a file name that is not constant at compile time");
        size_t slot = 0;
        for (; slot < g.textures.size() && g.textures[slot] != fn.svals[0]; ++slot) {}
        const bool texture_slot_used_before = slot < g.textures.size();
        if (slot == g.textures.size())
            g.textures.push_back(fn.svals[0]);
        int missingcolor = -1, missingalpha = -1, alpha_out = -1;
        auto dpart = [&](int ai, const char* which) {
            return A(ai).has_derivs && !A(ai).is_const() ? "(" + R(op.args[ai]) + ")." + which : std::string("0.0f");
        };
        std::string dd[4];
        size_t i = 4;
        if (op.args.size() >= 8 && A(4).type.base != Base::String && A(5).type.base != Base::String
            && A(6).type.base != Base::String && A(7).type.base != Base::String) {
            for (int k = 0; k < 4; ++k)
                dd[k] = comp(op.args[4 + k], 0, false);
            i = 8;
        } else {
            dd[0] = dpart(2, "dx"), dd[1] = dpart(3, "dx"), dd[2] = dpart(2, "dy"), dd[3] = dpart(3, "dy");
        }
        for (size_t k = i; k + 1 < op.args.size(); k += 2) {   // options that decide whether there is a lookup at all
            const Symbol& key = A((int)k);
            if (!key.const_value() || key.svals.empty())
                continue;
            if (key.svals[0] == "missingcolor")
                missingcolor = (int)k + 1;
            else if (key.svals[0] == "missingalpha")
                missingalpha = (int)k + 1;
            else if (key.svals[0] == "alpha")
                alpha_out = (int)k + 1;
        }
        bool triple = A(0).type.is_triple();
        if (missingcolor >= 0 || missingalpha >= 0) {
            // "missingcolor" / "missingalpha": a file that cannot be had is not an error, the result is the
            // missing colour (zero when only the alpha was given) and alpha the missing alpha
            // (llvm_gen_texture_options, llvm_gen.cpp:2611-2640; osl_texture, optexture.cpp:283-300).
            // Whether the image exists is settled here, when the group is compiled: images are
            // registered or found on texturepath before that (Textures).
            std::string terr;
            if (!oslb200::texture_get(fn.svals[0], g.texturepath, terr)) {
                if (!texture_slot_used_before)
                    g.textures.pop_back();   // no table entry for an image that is not there
                for (int c = 0; c < (triple ? 3 : 1); ++c)
                    w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", "
                      + (missingcolor >= 0 ? comp(op.args[missingcolor], A(missingcolor).type.is_triple() ? c : 0, false)
                                           : std::string("0.0f"))
                      + ");");
                if (alpha_out >= 0)
                    w("assign(" + R(op.args[alpha_out]) + ", "
                      + (missingalpha >= 0 ? comp(op.args[missingalpha], 0, false) : std::string("0.0f")) + ");");
                return;
            }
        }
        w("{");
        w("    TexOpt o_ = tex_default_options();");
        auto code_of = [&](int ai, bool wrap) -> std::string {
            const Symbol& v = A(ai);
            if (!v.const_value() || v.svals.empty())
                unsupported("texture option with a string that is not constant at compile time");
            const std::string& s = v.svals[0];
            if (wrap)
                return s == "clamp" ? "TEX_CLAMP" : s == "periodic" ? "TEX_PERIODIC" : s == "mirror" ? "TEX_MIRROR" : "TEX_BLACK";
            return s == "closest"                       ? "TEX_CLOSEST"
                   : (s == "bilinear" || s == "linear") ? "TEX_BILINEAR"
                   : (s == "bicubic" || s == "cubic")   ? "TEX_BICUBIC"
                                                        : "TEX_SMARTCUBIC";
        };
        for (; i + 1 < op.args.size(); i += 2) {
            const Symbol& key = A((int)i);
            if (!key.const_value() || key.svals.empty())
                unsupported("texture option with a non-constant name");
            const std::string& k = key.svals[0];
            int vi                = (int)i + 1;
            if (k == "wrap")
                w("    o_.swrap = o_.twrap = " + code_of(vi, true) + ";");
            else if (k == "swrap" || k == "twrap")
                w("    o_." + k + " = " + code_of(vi, true) + ";");
            else if (k == "width" || k == "blur")
                w("    o_.s" + k + " = o_.t" + k + " = " + comp(op.args[vi], 0, false) + ";");
            else if (k == "swidth" || k == "twidth" || k == "sblur" || k == "tblur" || k == "fill")
                w("    o_." + k + " = " + comp(op.args[vi], 0, false) + ";");
            else if (k == "interp")
                w("    o_.interp = " + code_of(vi, false) + ";");
            else if (k == "missingcolor" || k == "missingalpha" || k == "alpha")
                ;   // handled above
            else
                unsupported("texture option '" + k + "'");
        }
        if (alpha_out >= 0)
            unsupported("texture option 'alpha' of an existing image");
        w("    V3 r_ = texture_lookup(osl_tex_[" + std::to_string(g.texture_base + (int)slot) + "], o_, "
          + comp(op.args[2], 0, false) + ", " + comp(op.args[3], 0, false) + ", " + dd[0] + ", " + dd[1] + ", " + dd[2]
          + ", " + dd[3] + ", " + (triple ? "3" : "1") + ");");
        w("    assign(" + R(op.args[0]) + ", " + (triple ? "r_" : "r_.x") + ");");
        w("}");
    } else if (n == "luminance" || n == "transformc") {
        // osl_luminance_fv/_dfdv, osl_transformc (opcolor.cpp:464-518): dual form only
        // when both sides carry derivatives
        g.uses_colorsystem = true;
        int ci             = (int)op.args.size() - 1;
        const Symbol& c    = A(ci);
        bool dv            = A(0).has_derivs && c.has_derivs;
        std::string e      = R(op.args[ci]);
        if (c.is_const())
            e = "mkv(" + comp(op.args[ci], 0, false) + ", " + comp(op.args[ci], 1, false) + ", " + comp(op.args[ci], 2, false) + ")";
        else if (c.has_derivs && !dv)
            e = "nd(" + e + ")";
        if (n == "luminance") {
            need(2);
            w("assign(" + R(op.args[0]) + ", color_luminance(osl_cs_, " + e + "));");
        } else {
            need(4);
            w("assign(" + R(op.args[0]) + ", color_transformc(osl_cs_, " + space_code(op.args[1], false) + ", "
              + space_code(op.args[2], false) + ", " + e + "));");
        }
    } else if (n == "blackbody" || n == "wavelength_color") {
        need(2);
        g.uses_colorsystem = true;
        w("assign(" + R(op.args[0]) + ", color_" + std::string(n == "blackbody" ? "blackbody" : "wavelength") + "(osl_cs_, "
          + comp(op.args[1], 0, false) + "));");
    } else if ((n == "point" || n == "vector" || n == "normal") && op.args.size() == 5) {
        // llvm_gen_construct_triple (llvm_gen.cpp:1871-1933): build the triple, then
        // osl_transform_triple(space -> "common") on it, in place
        bool dv = false;
        if (A(0).has_derivs)
            for (int a = 2; a < 5; ++a)
                dv |= A(a).has_derivs;
        for (int c = 0; c < 3; ++c)
            w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", " + comp(op.args[2 + c], 0, dv) + ");");
        if (!(space_is(op.args[1], "common") || space_is(op.args[1], g.commonspace_synonym.c_str()))) {
            auto mf        = space_matrix(op.args[1], false);
            const char* vt = n == "point" ? "0" : (n == "vector" ? "1" : "2");
            std::string d  = R(op.args[0]);
            w("if (" + mf.second + ") assign(" + d + ", m44_transform(" + mf.first + ", " + d + ", " + vt + "));");
        }
    } else if (n == "color" || n == "point" || n == "vector" || n == "normal") {
        if (op.args.size() != 4)
            unsupported("triple constructor with a coordinate-system name");
        bool dv = false;
        if (A(0).has_derivs)
            for (int a = 1; a < 4; ++a)
                dv |= A(a).has_derivs;
        for (int c = 0; c < 3; ++c)
            w("setc(" + R(op.args[0]) + ", " + std::to_string(c) + ", " + comp(op.args[1 + c], 0, dv) + ");");
    } else if (n == "compref") {
        need(3);
        const Symbol& ix = A(2);
        if (ix.is_const()) {
            w("setc(" + R(op.args[0]) + ", 0, " + comp(op.args[1], ix.ivals.empty() ? 0 : ix.ivals[0], A(0).has_derivs) + ");");
        } else {
            // out of range: osl_range_check_err clamps (>= 3 -> 2, < 0 -> 0; shadingsys.cpp:4981-4984)
            w("switch (" + R(op.args[2]) + " < 0 ? 0 : " + R(op.args[2]) + ") {");
            for (int c = 0; c < 3; ++c)
                w(std::string(c == 2 ? "default: case " : "case ") + std::to_string(c) + ": setc(" + R(op.args[0])
                  + ", 0, " + comp(op.args[1], c, A(0).has_derivs) + "); break;");
            w("}");
        }
    } else if (n == "compassign") {
        need(3);
        const Symbol& ix = A(1);
        std::string v    = comp(op.args[2], 0, A(0).has_derivs);
        if (ix.is_const()) {
            w("setc(" + R(op.args[0]) + ", " + std::to_string(ix.ivals.empty() ? 0 : ix.ivals[0]) + ", " + v + ");");
        } else {
            w("switch (" + R(op.args[1]) + " < 0 ? 0 : " + R(op.args[1]) + ") {");
            for (int c = 0; c < 3; ++c)
                w(std::string(c == 2 ? "default: case " : "case ") + std::to_string(c) + ": setc(" + R(op.args[0])
                  + ", " + std::to_string(c) + ", " + v + "); break;");
            w("}");
        }
    } else if (n == "aref") {
        need(3);
        std::string len = std::to_string(A(1).type.arraylen);
        w("{ int ix_ = " + R(op.args[2]) + "; ix_ = ix_ < 0 ? 0 : (ix_ >= " + len + " ? " + len + " - 1 : ix_); ");
        if (A(0).type.base == Base::String || A(0).type.base == Base::Int)
            w("  " + R(op.args[0]) + " = " + R(op.args[1]) + "[ix_]; }");
        else
            w("  assign(" + R(op.args[0]) + ", " + R(op.args[1]) + "[ix_]); }");
    } else if (n == "aassign") {
        need(3);
        std::string len = std::to_string(A(0).type.arraylen);
        w("{ int ix_ = " + R(op.args[1]) + "; ix_ = ix_ < 0 ? 0 : (ix_ >= " + len + " ? " + len + " - 1 : ix_); ");
        if (A(0).type.base == Base::String || A(0).type.base == Base::Int)
            w("  " + R(op.args[0]) + "[ix_] = " + R(op.args[2]) + "; }");
        else
            w("  assign(" + R(op.args[0]) + "[ix_], " + R(op.args[2]) + "); }");
    } else if (n == "arraylength") {
        need(2);
        w(R(op.args[0]) + " = " + std::to_string(A(1).type.arraylen) + ";");
    } else if (n == "sincos") {
        need(3);
        bool dv = (A(1).has_derivs || A(2).has_derivs) && A(0).has_derivs;
        for (int k = 0; k < A(0).type.ncomp(); ++k) {
            std::string ks = std::to_string(k);
            w("{ auto x_ = " + comp(op.args[0], k, dv) + "; setc(" + R(op.args[1]) + ", " + ks + ", o_sin(x_)); setc("
              + R(op.args[2]) + ", " + ks + ", o_cos(x_)); }");
        }
    } else if (n == "dot" || n == "cross" || n == "length" || n == "distance" || n == "normalize") {
        bool dv = false;
        if (A(0).has_derivs)
            for (size_t a = 1; a < op.args.size(); ++a)
                dv |= A((int)a).has_derivs;
        std::string args;
        for (size_t a = 1; a < op.args.size(); ++a) {
            std::string e = R(op.args[a]);
            if (A((int)a).has_derivs && !dv)
                e = "nd(" + e + ")";
            args += (a > 1 ? ", " : "") + e;
        }
        w("assign(" + R(op.args[0]) + ", o_" + n + "(" + args + "));");
    } else if (n == "Dx" || n == "Dy" || n == "filterwidth") {
        need(2);
        w("assign(" + R(op.args[0]) + ", o_" + n + "(" + R(op.args[1]) + "));");
    } else if (n == "Dz") {
        need(2);
        if (A(1).symtype == SymType::Global && A(1).name == "P") {
            g.globals_read.insert(B200_SG_dPdz);
            w("assign(" + R(op.args[0]) + ", sg.dPdz);");
        } else
            w("assign(" + R(op.args[0]) + ", 0.0f);");
    } else if (n == "area") {
        need(2);
        w("assign(" + R(op.args[0]) + ", o_area(" + R(op.args[1]) + "));");
    } else if (n == "calculatenormal") {
        need(2);
        g.globals_read.insert(B200_SG_flipHandedness);
        w("assign(" + R(op.args[0]) + ", o_calculatenormal(" + R(op.args[1]) + ", sg.flipHandedness != 0));");
    } else if (n == "isnan") {
        w(R(op.args[0]) + " = isnan(" + comp(op.args[1], 0, false) + ") ? 1 : 0;");
    } else if (n == "isinf") {
        w(R(op.args[0]) + " = isinf(" + comp(op.args[1], 0, false) + ") ? 1 : 0;");
    } else if (n == "isfinite") {
        w(R(op.args[0]) + " = finitef(" + comp(op.args[1], 0, false) + ") ? 1 : 0;");
    } else if (n == "surfacearea") {
        g.globals_read.insert(B200_SG_surfacearea);
        w("assign(" + R(op.args[0]) + ", sg.surfacearea);");
    } else if (n == "backfacing") {
        g.globals_read.insert(B200_SG_backfacing);
        w(R(op.args[0]) + " = sg.backfacing;");
    } else if (n == "raytype") {
        need(2);
        if (!A(1).const_value())
            unsupported("raytype() of a non-constant name");
        g.globals_read.insert(B200_SG_raytype);
        w(R(op.args[0]) + " = (sg.raytype & " + std::to_string(raytype_bit(A(1).svals.empty() ? "" : A(1).svals[0]))
          + ") != 0;");
    } else if (n == "isconnected") {
        need(2);
        // (up ? 1 : 0) + ((down || renderer output) ? 2 : 0)  (runtimeoptimize.cpp:2495-2499)
        w(R(op.args[0]) + " = "
          + std::to_string((A(1).conn_layer >= 0 ? 1 : 0) + ((A(1).connected_down || A(1).out.placed) ? 2 : 0)) + ";");
    } else if (n == "isconstant") {
        // 1 when the runtime optimizer would have reduced the argument to a constant (constfold.cpp):
        // constants, instance parameter values, and top-level temporaries computed from those by
        // foldable arithmetic
        need(2);
        w(R(op.args[0]) + " = " + std::to_string(folds_to_constant(op.args[1], 0) ? 1 : 0) + ";");
    } else if (n == "hash") {
        size_t nin = op.args.size() - 1;
        std::string e;
        if (nin == 1 && A(1).type.base == Base::Int)
            e = "hash_i(" + R(op.args[1]) + ")";
        else if (nin == 1 && A(1).type.base == Base::Float)
            e = "hash_f(" + comp(op.args[1], 0, false) + ")";
        else if (nin == 2 && A(1).type.base == Base::Float)
            e = "hash_ff(" + comp(op.args[1], 0, false) + ", " + comp(op.args[2], 0, false) + ")";
        else if (nin == 1)
            e = "hash_v(nd(" + R(op.args[1]) + "))";
        else
            e = "hash_vf(nd(" + R(op.args[1]) + "), " + comp(op.args[2], 0, false) + ")";
        w(R(op.args[0]) + " = " + e + ";");
    } else if (n == "noise" || n == "snoise" || n == "cellnoise" || n == "hashnoise") {
        op_noise(op, false);
    } else if (n == "pnoise" || n == "psnoise" || n == "pcellnoise" || n == "phashnoise") {
        op_noise(op, true);
    } else if (n == "spline" || n == "splineinverse") {
        // llvm_gen_spline (llvm_gen.cpp:3673-3740): result, basis name, x, [nknots], knots[]
        if (op.args.size() < 4)
            unsupported("malformed spline op");
        bool has_count = op.args.size() == 5;
        int kn         = has_count ? op.args[4] : op.args[3];
        const Symbol &d = A(0), &basis = A(1), &x = A(2), &knots = S(kn);
        if (!knots.type.arraylen)
            unsupported("spline knots must be an array");
        std::string count = has_count ? R(op.args[3]) : std::to_string(knots.type.arraylen);
        bool dv           = d.has_derivs && (x.has_derivs || knots.has_derivs);
        static const char* names[] = { "catmull-rom", "bezier", "bspline", "hermite", "linear", "constant" };
        std::string bt;
        if (basis.const_value()) {
            std::string bn = basis.svals.empty() ? "" : basis.svals[0];
            int t          = 4;
            for (int k = 0; k < 6; ++k)
                if (bn == names[k])
                    t = k;
            if (bn == "catmullrom")
                t = 0;
            bt = std::to_string(t);
        } else {
            // runtime name: compare the interned string id against the known basis names
            std::string s = R(op.args[1]);
            bt            = "(";
            for (int k = 0; k < 6; ++k)
                if (k != 4)
                    bt += s + " == " + std::to_string(g.intern(names[k])) + " ? " + std::to_string(k) + " : ";
            bt += "4)";
        }
        std::string len = std::to_string(knots.type.arraylen);
        if (n == "splineinverse") {
            w("float k_[" + len + "]; for (int i_ = 0; i_ < " + len + "; ++i_) k_[i_] = nd(" + R(kn) + "[i_]);");
            // derivatives only flow from x (osl_splineinverse_dfdff / _dfdfdf ignore knot derivs)
            std::string xi = (d.has_derivs && x.has_derivs) ? R(op.args[2]) : "nd(" + R(op.args[2]) + ")";
            w("assign(" + R(op.args[0]) + ", spline_inverse(" + xi + ", k_, " + count + ", " + bt + "));");
        } else {
            std::string xe = R(op.args[2]);
            if (x.has_derivs && !dv)
                xe = "nd(" + xe + ")";
            std::string ke = R(kn);
            if (knots.has_derivs && !dv) {
                w(std::string(knots.type.is_triple() ? "V3" : "float") + " k_[" + len + "]; for (int i_ = 0; i_ < " + len
                  + "; ++i_) k_[i_] = nd(" + R(kn) + "[i_]);");
                ke = "k_";
            }
            w("spline_eval(" + R(op.args[0]) + ", " + xe + ", " + ke + ", " + count + ", " + bt + ");");
        }
    } else if (n == "closure") {
        // llvm_gen_closure (llvm_gen.cpp:3786-3903): [weight] name params...
        uses_closures = true;
        size_t i      = 1;
        int weight    = -1;
        if (i < op.args.size() && A((int)i).type.base != Base::String)
            weight = op.args[i++];
        if (i >= op.args.size() || !A((int)i).const_value())
            unsupported("closure with a non-constant name");
        std::string cname = A((int)i).svals.empty() ? "" : A((int)i).svals[0];
        ++i;
        struct Reg {
            const char* name;
            int nparams;
            const char* id;
            const char* keys;  // keyword parameters "name:i,name:f" in slot order after the positional words
        };
        static const Reg regs[] = {
            // MaterialX closures registered from libbsdl lobes (BSDLtoOSL, shading.cpp:156-182)
            { "oren_nayar_diffuse_bsdf", 3, "MX_OREN_NAYAR_DIFFUSE_ID", "energy_compensation:i" },
            { "burley_diffuse_bsdf", 3, "MX_BURLEY_DIFFUSE_ID", nullptr },
            { "sheen_bsdf", 3, "MX_SHEEN_ID", "mode:i" },
            { "layer", 2, "MX_LAYER_ID", nullptr },  // closure-typed params: pool word offsets
            { "uniform_edf", 1, "MX_UNIFORM_EDF_ID", nullptr },
            // libbsdl Data structs: bsdf_conductor_decl.h:43-46, bsdf_dielectric_decl.h:112-120 (c = color: 3 words)
            { "conductor_bsdf", 7, "MX_CONDUCTOR_ID", "thinfilm_thickness:f,thinfilm_ior:f" },
            { "dielectric_bsdf", 8, "MX_DIELECTRIC_ID", "thinfilm_thickness:f,thinfilm_ior:f,absorption:c,dispersion:f" },
            { "generalized_schlick_bsdf", 10, "MX_GENERALIZED_SCHLICK_ID", nullptr },
            { "translucent_bsdf", 2, "MX_TRANSLUCENT_ID", nullptr },
            { "subsurface_bssrdf", 4, "MX_SUBSURFACE_ID", nullptr },
            // spi::ThinLayerLobe (SpiThinLayer, shading.cpp:119-152; Data: SPI/bsdf_thinlayer_decl.h:112-124)
            { "thinlayer", 9, "SPI_THINLAYER", nullptr },
            // participating media (shading.cpp:265-284)
            { "anisotropic_vdf", 3, "MX_ANISOTROPIC_VDF_ID", nullptr },
            { "medium_vdf", 6, "MX_MEDIUM_VDF_ID", nullptr },
            { "emission", 0, "EMISSION_ID" },       { "background", 0, "BACKGROUND_ID" },
            { "diffuse", 1, "DIFFUSE_ID" },         { "oren_nayar", 2, "OREN_NAYAR_ID" },
            { "translucent", 1, "TRANSLUCENT_ID" }, { "phong", 2, "PHONG_ID" },
            { "ward", 4, "WARD_ID" },               { "reflection", 1, "REFLECTION_ID" },
            { "reflection", 2, "FRESNEL_REFLECTION_ID" }, { "refraction", 2, "REFRACTION_ID" },
            { "transparent", 0, "TRANSPARENT_ID" }, { "transparent_bsdf", 0, "MX_TRANSPARENT_ID" },
            { "microfacet", 7, "MICROFACET_ID" },
        };
        auto find_reg = [&](size_t npos) -> const Reg* {
            for (const Reg& r : regs)
                if (cname == r.name && (int)npos == r.nparams)
                    return &r;
            return nullptr;
        };
        // positional parameters end where a constant string followed by a value starts the
        // "keyword", value pairs (llvm_gen_closure / llvm_gen_keyword_fill)
        std::vector<int> pos, kw;
        for (; i < op.args.size(); ++i) {
            const Symbol& a = S(op.args[i]);
            if (a.type.base == Base::String && a.const_value() && i + 1 < op.args.size() && find_reg(pos.size()))
                break;
            pos.push_back(op.args[i]);
        }
        for (; i < op.args.size(); ++i)
            kw.push_back(op.args[i]);
        const Reg* reg     = find_reg(pos.size());
        const char* idname = reg ? reg->id : nullptr;
        if (!idname)
            unsupported("closure '" + cname + "' with " + std::to_string(pos.size()) + " parameters is not registered");
        int nwords = 0;
        for (int a : pos)
            nwords += S(a).type.ncomp();
        // keyword slots: zero unless given (the reference memsets the parameter block)
        std::vector<std::pair<std::string, char>> keys;
        if (reg->keys) {
            std::string ks = reg->keys;
            size_t p = 0;
            while (p < ks.size()) {
                size_t c = ks.find(',', p);
                std::string item = ks.substr(p, c == std::string::npos ? std::string::npos : c - p);
                keys.push_back({ item.substr(0, item.find(':')), item.back() });
                if (c == std::string::npos)
                    break;
                p = c + 1;
            }
        }
        const int key_base = nwords;
        for (auto& kt : keys)
            nwords += kt.second == 'c' ? 3 : 1;
        const std::string cn = cname;
        g.closure_in_loop |= loop_depth > 0;
        g.pool_words_bound += 4 + nwords;
        g.closure_names.insert(cname);
        if (cn == "layer")
            g.closure_adds += 1;
        else if (cn == "anisotropic_vdf" || cn == "medium_vdf")
            g.uses_media = true;
        if (cn == "sheen_bsdf" && !kw.empty())   // a "mode" keyword may select the Zeltner-Burley LTC sheen
            g.uses_sheen_ltc = true;
        else if (cn != "emission" && cn != "background" && cn != "uniform_edf")
            g.lobe_bound += 1;
        if (cn == "phong" || cn == "ward" || cn == "microfacet" || cn == "oren_nayar" || cn == "oren_nayar_diffuse_bsdf"
            || cn == "burley_diffuse_bsdf" || cn == "sheen_bsdf" || cn == "layer")
            g.uses_glossy_lobes = true;
        if (cn == "conductor_bsdf" || cn == "dielectric_bsdf" || cn == "generalized_schlick_bsdf"
            || cn == "translucent_bsdf" || cn == "subsurface_bssrdf")
            g.uses_glossy_lobes = g.uses_mx_lobes = true;
        if (cn == "thinlayer")   // spi::ThinLayerLobe: GGXDist, energy curve and frame of the MX lobes + its own state
            g.uses_glossy_lobes = g.uses_mx_lobes = g.uses_thinlayer = true;
        std::string wexpr = "mkv(1.0f)";
        if (weight >= 0) {
            w("V3 w_; assign(w_, " + R(weight) + ");");
            wexpr = "w_";
        }
        w(std::string("int c_ = clos_component(*sg.pool, ") + idname + ", " + std::to_string(nwords) + ", "
          + (weight >= 0 ? "true" : "false") + ", " + wexpr + ");");
        w("if (c_) {");
        int off = 0;
        for (int a : pos) {
            std::string e = R(a);
            if (S(a).has_derivs)
                e = "nd(" + e + ")";
            if (S(a).type.base == Base::String) {
                // the only string parameter on this path is microfacet's distribution
                // name; store the code process_closure switches on (shading.cpp:1518-1534)
                auto id = [&](const char* s) { return std::to_string(g.intern(s)); };
                e = "((" + e + ") == " + id("ggx") + " ? 1 : (" + e + ") == " + id("beckmann") + " ? 2 : (" + e
                    + ") == " + id("default") + " ? 3 : 0)";
            }
            w("    putp(sg.pool->w + c_ + " + std::to_string(4 + off) + ", " + e + ");");
            off += S(a).type.ncomp();
        }
        int koff = key_base;
        for (size_t k = 0; k < keys.size(); ++k) {
            const char kt = keys[k].second;
            std::string e = kt == 'i' ? "0" : (kt == 'c' ? "mkv(0.0f)" : "0.0f");
            for (size_t j = 0; j + 1 < kw.size(); j += 2) {
                const Symbol& ks = S(kw[j]);
                const Symbol& kv = S(kw[j + 1]);
                bool type_ok = kt == 'i' ? kv.type.base == Base::Int
                                         : (kt == 'c' ? kv.type.ncomp() == 3 : kv.type.base == Base::Float);
                if (ks.type.base == Base::String && ks.const_value() && !ks.svals.empty() && ks.svals[0] == keys[k].first
                    && type_ok && kv.type.arraylen == 0) {
                    e = R(kw[j + 1]);
                    if (kv.has_derivs)
                        e = "nd(" + e + ")";
                }
            }
            w("    putp(sg.pool->w + c_ + " + std::to_string(4 + koff) + ", " + e + ");");
            koff += kt == 'c' ? 3 : 1;
        }
        w("}");
        w(R(op.args[0]) + " = c_;");
    } else if (n == "setmessage") {
        // osl_setmessage (opmessage.cpp): the first set of a name wins, a second one is an error
        // in the reference (reported, value ignored)
        const Symbol& nm = S(op.args[0]);
        auto it = (nm.type.base == Base::String && nm.const_value() && !nm.svals.empty()) ? messages.find(nm.svals[0])
                                                                                            : messages.end();
        const Symbol& v = S(op.args[1]);
        if (it != messages.end() && v.type.base == it->second.base && v.type.ncomp() == it->second.ncomp && !v.type.arraylen) {
            std::string bit = std::to_string(1u << it->second.id) + "u";
            std::string e   = R(op.args[1]);
            if (v.has_derivs)
                e = "nd(" + e + ")";
            w("if (!(gd.msgset & " + bit + ")) { gd.M" + std::to_string(it->second.id) + " = " + e + "; gd.msgset |= " + bit + "; }");
        } else
            g.warnings.push_back("setmessage in layer '" + L->layername + "' is not supported on the device (name not constant, "
                                 "or a closure / string / matrix / array value): ignored");
    } else if (n == "getmessage") {
        // osl_getmessage: 1 and the value when a message of that name AND type was set before
        const bool has_source = op.args.size() == 4;
        const Symbol& nm      = S(op.args[has_source ? 2 : 1]);
        const Symbol& dst     = S(op.args.back());
        auto it = (!has_source && nm.type.base == Base::String && nm.const_value() && !nm.svals.empty())
                      ? messages.find(nm.svals[0])
                      : messages.end();
        if (it != messages.end() && dst.type.base == it->second.base && dst.type.ncomp() == it->second.ncomp
            && !dst.type.arraylen) {
            std::string bit = std::to_string(1u << it->second.id) + "u";
            std::string val = "gd.M" + std::to_string(it->second.id);
            if (dst.has_derivs)
                val = dst.type.ncomp() == 3 ? "mkdv(" + val + ")" : "mkd(" + val + ")";
            w("if (gd.msgset & " + bit + ") { " + R(op.args.back()) + " = " + val + "; " + R(op.args[0]) + " = 1; } else "
              + R(op.args[0]) + " = 0;");
        } else
            w(R(op.args[0]) + " = 0;");
    } else if (n == "getattribute") {
        // llvm_gen_getattribute (llvm_gen.cpp:3303-3420) -> RendererServices::get_attribute.  On this
        // back end a renderer attribute is either one of the few the harness renderers answer for
        // every object ("osl:version", "shading:index") or per-point userdata (testshade's
        // SimpleRenderer falls back to get_userdata, simplerend.cpp:480-503): the b200_userdata
        // entry of that name and type.  Object / array-index lookups and names only known at run
        // time find nothing.
        const Symbol& dst = S(op.args.back());
        const Symbol& nm  = S(op.args[1]);
        const std::string res = R(op.args[0]);
        bool done = false;
        // uniform renderer attributes (b200_attribute): the statements that hand attribute `at` to dst,
        // or "" when the requested type does not match (get_attribute is asked for an exact TypeDesc)
        auto attr_stores = [&](const Attribute& at) -> std::string {
            const int n = (dst.type.arraylen ? dst.type.arraylen : 1) * dst.type.ncomp();
            const std::string d = R((int)op.args.back());
            std::string out;
            if (dst.type.base == Base::Int) {
                if (at.type != 0 || (int)at.ivals.size() != n)
                    return "";
                for (int k = 0; k < n; ++k)
                    out += d + (dst.type.arraylen ? "[" + std::to_string(k) + "]" : "") + " = " + std::to_string(at.ivals[k]) + "; ";
            } else if (dst.type.base == Base::String) {
                if (at.type != 2 || (int)at.svals.size() != n)
                    return "";
                for (int k = 0; k < n; ++k)
                    out += d + (dst.type.arraylen ? "[" + std::to_string(k) + "]" : "") + " = "
                           + std::to_string(g.intern(at.svals[k])) + "; ";
            } else if (dst.type.base == Base::Float || dst.type.is_triple()) {
                if (at.type != 1 || (int)at.fvals.size() != n)
                    return "";
                const int per = dst.type.ncomp(), cnt = dst.type.arraylen ? dst.type.arraylen : 1;
                for (int e = 0; e < cnt; ++e) {
                    const std::string el = d + (dst.type.arraylen ? "[" + std::to_string(e) + "]" : "");
                    if (per == 1)
                        out += "assign(" + el + ", " + cfloat(at.fvals[e]) + "); ";
                    else
                        out += "assign(" + el + ", mkv(" + cfloat(at.fvals[3 * e]) + ", " + cfloat(at.fvals[3 * e + 1]) + ", "
                               + cfloat(at.fvals[3 * e + 2]) + ")); ";
                }
            } else
                return "";
            return out;
        };
        if (op.args.size() == 3 && nm.type.base == Base::String && !g.attributes.empty() && !material_mode) {
            if (nm.const_value() && !nm.svals.empty()) {
                for (const Attribute& at : g.attributes)
                    if (!done && at.name == nm.svals[0]) {
                        const std::string st = attr_stores(at);
                        if (!st.empty()) {
                            w(st);
                            w(res + " = 1;");
                            done = true;
                        }
                    }
            } else if (!nm.type.arraylen) {
                // a name the shader computes: compared with the table at run time
                w("{ const int nm_ = " + R(op.args[1]) + "; " + res + " = 0;");
                for (const Attribute& at : g.attributes) {
                    const std::string st = attr_stores(at);
                    if (!st.empty())
                        w("  if (nm_ == " + std::to_string(g.intern(at.name)) + ") { " + st + res + " = 1; }");
                }
                w("}");
                done = true;
            }
        }
        if (!done && op.args.size() == 3 && nm.type.base == Base::String && nm.const_value() && !nm.svals.empty()
            && !dst.type.arraylen && !material_mode) {
            const std::string& name = nm.svals[0];
            if (name == "osl:version" && dst.type.base == Base::Int) {
                w(R(op.args[2]) + " = 11600;");
                w(res + " = 1;");
                done = true;
            } else if (name == "shading:index" && dst.type.base == Base::Int) {
                w(R(op.args[2]) + " = sg.shadeindex;");
                w(res + " = 1;");
                done = true;
            } else if ((name == "shader:shadername" || name == "shader:layername" || name == "shader:groupname")
                       && dst.type.base == Base::String) {
                // folded at optimisation time when the name is a constant (constfold_getattribute, constfold.cpp)
                const std::string& val = name == "shader:shadername" ? L->m.shadername
                                         : name == "shader:layername" ? L->layername : g.name;
                w(R(op.args[2]) + " = " + std::to_string(g.intern(val)) + ";");
                w(res + " = 1;");
                done = true;
            } else
                for (const UserData& u : g.userdata)
                    if (!done && u.name == name && u.ncomp == dst.type.ncomp() && u.is_int == (dst.type.base == Base::Int)
                        && (dst.type.base == Base::Int || dst.type.base == Base::Float || dst.type.is_triple())) {
                        emit_userdata_load(u, dst, R(op.args[2]), "ga_");
                        w(res + " = ga_ ? 1 : 0;");
                        done = true;
                    }
        }
        if (!done)
            w(res + " = 0;");
    } else if ((n == "printf" || n == "error" || n == "warning") && journal_ok) {
        emit_printf(op);
    } else if (n == "printf" || n == "error" || n == "warning" || n == "fprintf") {
        // error / warning / fprintf, and printf inside renderer materials, have no effect on
        // shading results: dropped with a recorded warning.
        std::string msg = "op '" + n + "' ignored on device (no journal yet) in layer '" + L->layername + "'";
        bool dup = false;
        for (auto& s : g.warnings)
            dup |= (s == msg);
        if (!dup)
            g.warnings.push_back(msg);
    } else {
        unsupported("op '" + n + "' is not implemented");
    }
}

void
Gen::gen_layer(int layer)
{
    L                    = &g.layers[layer];
    li                   = layer;
    Master& m            = L->m;
    std::string k        = std::to_string(layer);
    int nlayers          = (int)g.layers.size();
    ind                  = 0;
    w("static __device__ __forceinline__ void layer_" + k + "(SG& sg, GD& gd, const B200Launch& L)");
    w("{");
    ind = 1;
    w("gd.ran |= " + std::to_string(1u << layer) + "u;");
    for (Symbol& s : m.syms) {
        if (s.symtype == SymType::Local || s.symtype == SymType::Temp) {
            std::string arr = s.type.arraylen ? "[" + std::to_string(s.type.arraylen) + "]" : "";
            w(ctype(s) + " " + ident(s.name) + arr + (s.type.arraylen ? " = {};" : " = {};"));
        } else if (s.is_const() && s.type.arraylen) {
            std::string init;
            for (auto& v : initvals(s))
                init += (init.empty() ? "" : ", ") + v;
            w("const " + ctype(s) + " K_" + ident(s.name) + "[" + std::to_string(s.type.arraylen) + "] = {" + init + "};");
        }
    }
    // written globals that carry derivatives: layer-local copies (see ref())
    std::vector<std::string> written_globals;
    for (Symbol& s : m.syms)
        if (s.symtype == SymType::Global && s.written && s.has_derivs && s.name != "Ci") {
            auto it = global_table().find(s.name);
            if (it != global_table().end() && it->second.dx >= 0) {
                g.globals_read.insert(it->second.field);
                g.globals_read.insert(it->second.dx);
                g.globals_read.insert(it->second.dy);
                w(std::string(s.type.is_triple() ? "Dv" : "Df") + " gw_" + s.name + " = " + it->second.expr + "_d();");
                written_globals.push_back(s.name);
            }
        }
    // the group entry runs earlier non-lazy layers unconditionally
    // (llvm_instance.cpp:1693-1720)
    if (layer == nlayers - 1)
        for (int e = 0; e < nlayers - 1; ++e)
            if (!g.layers[e].unused && !g.layers[e].lazy)
                w("if (!(gd.ran & " + std::to_string(1u << e) + "u)) layer_" + std::to_string(e) + "(sg, gd, L);");
    // parameter initialisation (llvm_instance.cpp:703-1000, 1666-1690)
    for (size_t si = 0; si < m.syms.size(); ++si) {
        Symbol& s = m.syms[si];
        if (!s.is_param() || s.conn_layer >= 0)
            continue;
        std::vector<std::string> vals = initvals(s);
        std::string r                 = ref(layer, s);
        // interpolated parameter: the renderer's userdata wins (osl_bind_interpolated_param,
        // llvm_instance.cpp:805-970), points without it run the default / init ops
        const UserData* ud = nullptr;
        if (s.interpolated && !material_mode && !s.type.arraylen)
            for (const UserData& u : g.userdata)
                if (u.name == s.name && u.ncomp == s.type.ncomp() && u.is_int == (s.type.base == Base::Int)
                    && s.type.base != Base::String && s.type.base != Base::Matrix && s.type.base != Base::Closure)
                    ud = &u;
        if (ud) {
            emit_userdata_load(*ud, s, r, "got_" + std::to_string(si));
            w("if (!got_" + std::to_string(si) + ") {");
            ++ind;
        }

        if (s.type.arraylen)
            for (size_t i = 0; i < vals.size(); ++i)
                w(r + "[" + std::to_string(i) + "] = " + vals[i] + ";");
        else
            w(r + " = " + vals[0] + ";");
        auto mi = m.methods.find(s.name);
        if (s.initexpr && mi != m.methods.end()) {
            ensured.clear();
            emit_block(mi->second.first, mi->second.second, nullptr);
        }
        if (ud) {
            --ind;
            w("}");
        }
    }
    ensured.clear();
    auto mi = m.methods.find("___main___");
    if (mi != m.methods.end())
        emit_block(mi->second.first, mi->second.second, nullptr);
    w("layer_end:;");
    for (const std::string& gn : written_globals) {
        const std::string e = global_table().find(gn)->second.expr;
        w(e + " = gw_" + gn + ".val; " + e + "_dx = gw_" + gn + ".dx; " + e + "_dy = gw_" + gn + ".dy;");
    }
    // hand results to downstream layers (llvm_instance.cpp:1738-1802)
    for (const Connection& c : g.connections)
        if (c.srclayer == layer && !g.layers[c.dstlayer].unused)
            emit_copy(c.dstlayer, g.layers[c.dstlayer].m.syms[c.dstsym], layer, m.syms[c.srcsym]);
    // renderer outputs are written by the kernel epilogue from the group data
    // (reference: llvm_instance.cpp:1807-1848 does it at the end of the layer)
    ind = 0;
    w("}");
}

// Kernel prologue/epilogue shared by every group.  Loads touch only the
// ShaderGlobals planes the group reads; uniform fields come from the launch
// block (constant bank).
const char* PRELUDE = R"CUDA(
using namespace osld;

struct B200Launch {
    const float* varying[%NFIELDS%];
    float uniform[%NFIELDS%][4];
    long long plane_stride;
    const int* shadeindex;
    void* output_base;
    const void* userdata_base;
    long long npoints;
    long long shadeindex_base;   // added to the point index when shadeindex == NULL
    long long out_adjust[%MAXOUT%];   // per-output byte rebase (host staging path)
    int stage_outputs;                // 1: stage dense output records in shared memory
    int pad_;
    // named coordinate systems the group references (slot order fixed at code generation):
    // [k][0] = space -> common, [k][1] = its inverse; xf_ok[k] = 0 when the renderer does
    // not know the name (identity is stored)
    float xf[%MAXSPACES%][2][16];
    int xf_ok[%MAXSPACES%];
    // printf journal: [0] = words used (bumped atomically), [1] = overflow flag, records from [2]
    unsigned* journal;
    unsigned journal_words;
    unsigned pad2_;
};

__device__ __forceinline__ float ldf(const B200Launch& L, int f, int c, long long i)
{
    const float* p = L.varying[f];
    return p ? __ldg(p + c * L.plane_stride + i) : L.uniform[f][c];
}
__device__ __forceinline__ V3 ldv(const B200Launch& L, int f, long long i)
{
    return mkv(ldf(L, f, 0, i), ldf(L, f, 1, i), ldf(L, f, 2, i));
}
__device__ __forceinline__ int ldi(const B200Launch& L, int f, long long i)
{
    const int* p = (const int*)L.varying[f];
    return p ? __ldg(p + i) : __float_as_int(L.uniform[f][0]);
}
)CUDA";

const char* OUTPUT_HELPERS = R"CUDA(
__device__ __forceinline__ float* outp(const B200Launch& L, const SG& sg, int k, long long offset, long long stride)
{
    return (float*)((char*)L.output_base + offset + L.out_adjust[k] + stride * (long long)sg.shadeindex);
}
// ---- printf journal -------------------------------------------------------------
// Reserve one record and fill its header; returns where the argument words go, or
// nullptr when there is no journal or it is full (the overflow flag is raised).
__device__ __forceinline__ unsigned* jr_reserve(const B200Launch& L, SG& sg, unsigned fmt, unsigned nargs)
{
    if (!L.journal)
        return nullptr;
    const unsigned n  = nargs + 4u;
    const unsigned at = atomicAdd(L.journal, n);
    const unsigned seq = sg.jseq++;
    if (at + n + 2u > L.journal_words) {
        L.journal[1] = 1u;
        return nullptr;
    }
    unsigned* r = L.journal + 2 + at;
    r[0] = n;
    r[1] = (unsigned)sg.shadeindex;
    r[2] = seq;
    r[3] = fmt;
    return r + 4;
}
// ---- shared-memory staging of dense output records + TMA bulk store ----------
// Each thread drops its record (W words) at stage[t*W ..]; 16-byte stores when W%4==0
// are bank-conflict free for the 48 B record of the layered group, scalar stores
// are conflict free for odd W (12 B colour records).
template<int W> __device__ __forceinline__ void stage_record(float* s, int t, const float (&rec)[W])
{
    if (W % 4 == 0) {
        float4* p = reinterpret_cast<float4*>(s + t * W);
#pragma unroll
        for (int j = 0; j < W / 4; ++j)
            p[j] = make_float4(rec[4 * j], rec[4 * j + 1], rec[4 * j + 2], rec[4 * j + 3]);
    } else if (W % 2 == 0) {
        float2* p = reinterpret_cast<float2*>(s + t * W);
#pragma unroll
        for (int j = 0; j < W / 2; ++j)
            p[j] = make_float2(rec[2 * j], rec[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < W; ++j)
            s[t * W + j] = rec[j];
    }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, unsigned bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_nocommit(void* gdst, const void* ssrc, unsigned bytes)
{
    unsigned sa = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recent group have finished READING shared memory (double buffering)
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void wr(float* p, float v) { p[0] = v; }
__device__ __forceinline__ void wr(float* p, int v) { ((int*)p)[0] = v; }
__device__ __forceinline__ void wr(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ void wr(float* p, Df v) { p[0] = v.val; }
__device__ __forceinline__ void wr(float* p, const Dv& v) { wr(p, v.val); }
__device__ __forceinline__ void wrd(float* p, float v) { p[0] = v; p[1] = 0.0f; p[2] = 0.0f; }
__device__ __forceinline__ void wrd(float* p, Df v) { p[0] = v.val; p[1] = v.dx; p[2] = v.dy; }
__device__ __forceinline__ void wrd(float* p, V3 v) { wr(p, v); for (int i = 3; i < 9; ++i) p[i] = 0.0f; }
__device__ __forceinline__ void wrd(float* p, const Dv& v) { wr(p, v.val); wr(p + 3, v.dx); wr(p + 6, v.dy); }
)CUDA";

std::string
Gen::run()
{
    int nlayers = (int)g.layers.size();
    if (nlayers > 32)
        throw std::runtime_error("B200 back end: more than 32 layers in a group is not supported yet");
    // layer bodies first (they record which globals are read)
    journal_ok = g.journal_enabled;
    g.textures.clear();
    g.texture_base     = 0;
    g.pool_words_bound = 1;
    g.lobe_bound = g.closure_adds = 0;
    g.closure_in_loop             = false;
    std::ostringstream bodies;
    scan_messages();
    std::string gd = "struct GD {\n    unsigned ran;\n";
    for (int l = 0; l < nlayers; ++l) {
        if (g.layers[l].unused)
            continue;
        L  = &g.layers[l];
        li = l;
        for (Symbol& s : g.layers[l].m.syms)
            if (s.is_param()) {
                std::string arr = s.type.arraylen ? "[" + std::to_string(s.type.arraylen) + "]" : "";
                gd += "    " + ctype(s) + " L" + std::to_string(l) + "_" + ident(s.name) + arr + ";\n";
            }
    }
    gd += message_fields();
    gd += "};\n";
    for (int l = 0; l < nlayers; ++l)
        if (!g.layers[l].unused)
            o << "static __device__ __forceinline__ void layer_" << l << "(SG& sg, GD& gd, const B200Launch& L);\n";
    for (int l = 0; l < nlayers; ++l)
        if (!g.layers[l].unused)
            gen_layer(l);
    std::string layers_src = o.str();

    // SG holds only what is read
    auto rd = [&](int f) { return g.globals_read.count(f) != 0; };
    std::ostringstream sg, ld;
    sg << "struct SG {\n    int shadeindex;\n    unsigned jseq;   // printf records written so far by this point\n";
    struct TD {
        const char* name;
        int f, dx, dy;
    };
    const TD triples_d[] = { { "P", B200_SG_P, B200_SG_dPdx, B200_SG_dPdy },
                             { "I", B200_SG_I, B200_SG_dIdx, B200_SG_dIdy },
                             { "Ps", B200_SG_Ps, B200_SG_dPsdx, B200_SG_dPsdy } };
    for (const TD& t : triples_d) {
        if (!rd(t.f))
            continue;
        sg << "    V3 " << t.name << ";\n";
        ld << "        sg." << t.name << " = ldv(L, " << t.f << ", i);\n";
        if (rd(t.dx)) {
            sg << "    V3 " << t.name << "_dx, " << t.name << "_dy;\n";
            sg << "    __device__ __forceinline__ Dv " << t.name << "_d() const { return mkdv(" << t.name << ", "
               << t.name << "_dx, " << t.name << "_dy); }\n";
            ld << "        sg." << t.name << "_dx = ldv(L, " << t.dx << ", i);\n";
            ld << "        sg." << t.name << "_dy = ldv(L, " << t.dy << ", i);\n";
        }
    }
    const TD scal_d[] = { { "u", B200_SG_u, B200_SG_dudx, B200_SG_dudy }, { "v", B200_SG_v, B200_SG_dvdx, B200_SG_dvdy } };
    for (const TD& t : scal_d) {
        if (!rd(t.f))
            continue;
        sg << "    float " << t.name << ";\n";
        ld << "        sg." << t.name << " = ldf(L, " << t.f << ", 0, i);\n";
        if (rd(t.dx)) {
            sg << "    float " << t.name << "_dx, " << t.name << "_dy;\n";
            sg << "    __device__ __forceinline__ Df " << t.name << "_d() const { return mkd(" << t.name << ", "
               << t.name << "_dx, " << t.name << "_dy); }\n";
            ld << "        sg." << t.name << "_dx = ldf(L, " << t.dx << ", 0, i);\n";
            ld << "        sg." << t.name << "_dy = ldf(L, " << t.dy << ", 0, i);\n";
        }
    }
    const TD plain3[] = { { "N", B200_SG_N, -1, -1 },       { "Ng", B200_SG_Ng, -1, -1 },
                          { "dPdu", B200_SG_dPdu, -1, -1 }, { "dPdv", B200_SG_dPdv, -1, -1 },
                          { "dPdtime", B200_SG_dPdtime, -1, -1 }, { "dPdz", B200_SG_dPdz, -1, -1 } };
    for (const TD& t : plain3)
        if (rd(t.f)) {
            sg << "    V3 " << t.name << ";\n";
            ld << "        sg." << t.name << " = ldv(L, " << t.f << ", i);\n";
        }
    const TD plain1[] = { { "time", B200_SG_time, -1, -1 }, { "dtime", B200_SG_dtime, -1, -1 },
                          { "surfacearea", B200_SG_surfacearea, -1, -1 } };
    for (const TD& t : plain1)
        if (rd(t.f)) {
            sg << "    float " << t.name << ";\n";
            ld << "        sg." << t.name << " = ldf(L, " << t.f << ", 0, i);\n";
        }
    const TD ints[] = { { "raytype", B200_SG_raytype, -1, -1 }, { "flipHandedness", B200_SG_flipHandedness, -1, -1 },
                        { "backfacing", B200_SG_backfacing, -1, -1 } };
    for (const TD& t : ints)
        if (rd(t.f)) {
            sg << "    int " << t.name << ";\n";
            ld << "        sg." << t.name << " = ldi(L, " << t.f << ", i);\n";
        }
    // Closures in a grid execution (testshade): built in a per-point pool like the reference's
    // StackClosurePool (render_state.h:27-54) and dropped - only the integrator consumes them.
    if (uses_closures)
        sg << "    int Ci;\n    ClosurePool* pool;\n";
    sg << "};\n";

    std::string prelude = PRELUDE;
    for (size_t p; (p = prelude.find("%NFIELDS%")) != std::string::npos;)
        prelude.replace(p, 9, std::to_string((int)B200_SG_NFIELDS));
    for (size_t p; (p = prelude.find("%MAXOUT%")) != std::string::npos;)
        prelude.replace(p, 8, std::to_string(B200_MAX_OUTPUTS));
    for (size_t p; (p = prelude.find("%MAXSPACES%")) != std::string::npos;)
        prelude.replace(p, 11, std::to_string(B200_MAX_SPACES));
    if ((int)g.outputs.size() > B200_MAX_OUTPUTS)
        throw std::runtime_error("B200 back end: too many renderer outputs in one group");

    std::ostringstream out;
    out << "// generated by libosl_b200 for shader group '" << g.name << "'\n";
    out << "#include \"osl_b200_device.cuh\"\n";
    if (uses_closures)
        out << "#define OSLD_POOL_WORDS " << (g.closure_in_loop ? 256 : std::min(256, std::max(2, g.pool_words_bound)))
            << "\n#include \"osl_b200_closure.cuh\"\n";
    if (!g.textures.empty())
        out << "#include \"osl_b200_texture.cuh\"\nextern \"C\" __device__ osld::TexDesc osl_tex_[" << g.textures.size()
            << "];\n";
    if (g.uses_colorsystem)
        out << "#include \"osl_b200_color.cuh\"\n" << colorsystem_cuda_definition(g.colorspace);
    out << prelude << sg.str() << OUTPUT_HELPERS << gd << layers_src;
    // ---- kernel: one CTA walks tiles of BLOCK consecutive points --------------
    auto out_sym  = [&](int k) -> Symbol& { return g.layers[g.outputs[k].first].m.syms[g.outputs[k].second]; };
    auto out_expr = [&](int k) {
        L  = &g.layers[g.outputs[k].first];
        li = g.outputs[k].first;
        const Symbol& s = out_sym(k);
        if (s.symtype == SymType::Global) {   // a ShaderGlobals field handed back: read it from sg, not from a layer local
            const GlobalInfo& gi = global_table().at(s.name);
            return std::string(gi.expr) + (s.out.derivs && gi.dx >= 0 && s.has_derivs ? "_d()" : "");
        }
        return ref(li, s);
    };
    std::string B = "%BLOCK%";
    long long stage_words = 0;
    for (const OutCluster& c : g.clusters)
        stage_words += c.stride / 4;
    out << "extern \"C\" __global__ void __launch_bounds__(" << B << "%MINBLOCKS%"
        << ") osl_b200_group_kernel(const __grid_constant__ B200Launch L)\n{\n";
    // Staging is per WARP and double buffered: a warp's 32 records are contiguous in the
    // output arena, so lane 0 issues the warp's own TMA bulk store and only __syncwarp is
    // needed - no CTA barrier on the path (the CTA-wide version spent 43 % of its stall
    // cycles in barriers, profiles/ncu_layers4096_r01_details.txt).  The globals of the next
    // tile are loaded before the current tile is shaded, so their latency hides behind it.
    if (g.stage_ok)
        out << "    __shared__ __align__(128) float stage_[2 * " << stage_words << " * " << B << "];\n";
    out << "    const long long ntiles_ = (L.npoints + " << B << " - 1) / " << B << ";\n";
    out << "    const bool staged_ = " << (g.stage_ok ? "(L.shadeindex == nullptr) && (L.stage_outputs != 0)" : "false") << ";\n";
    std::string ldn = ld.str();
    for (size_t p = 0; (p = ldn.find("sg.", p)) != std::string::npos; p += 5)
        ldn.replace(p, 3, "sgn_.");
    for (size_t p = 0; (p = ldn.find(", i);", p)) != std::string::npos; p += 7)
        ldn.replace(p, 5, ", in_);");
    auto emit_fetch = [&](const char* tile_expr) {
        out << "        {\n            const long long tn_ = " << tile_expr << ";\n";
        out << "            const long long in_ = tn_ * " << B << " + threadIdx.x;\n";
        out << "            actn_ = tn_ < ntiles_ && in_ < L.npoints;\n";
        out << "            if (actn_) {\n";
        out << "                sgn_.shadeindex = L.shadeindex ? __ldg(L.shadeindex + in_) : (int)(in_ + L.shadeindex_base);\n";
        out << ldn;
        out << "            }\n        }\n";
    };
    out << "    SG sgn_ = SG();\n    bool actn_ = false;\n";
    emit_fetch("(long long)blockIdx.x");
    out << "    int it_ = 0;\n";
    out << "    for (long long tile_ = blockIdx.x; tile_ < ntiles_; tile_ += gridDim.x, ++it_) {\n";
    out << "        const long long i = tile_ * " << B << " + threadIdx.x;\n";
    out << "        const bool active_ = actn_;\n";
    out << "        SG sg = sgn_;\n        sg.jseq = 0u;\n";
    emit_fetch("tile_ + gridDim.x");
    out << "        GD gd;\n        gd.ran = 0u;\n";
    if (!messages.empty())
        out << "        gd.msgset = 0u;\n";
    if (uses_closures)
        out << "        float pool_store_[OSLD_POOL_STORE];\n        ClosurePool pool_;\n        pool_.bind(pool_store_, 1);\n"
               "        pool_.reset();\n        sg.pool = &pool_;\n        sg.Ci = 0;\n";
    out << "        if (active_) {\n";
    out << "            layer_" << (nlayers - 1) << "(sg, gd, L);\n";
    out << "        }\n";
    if (g.stage_ok) {
        out << "        if (staged_) {\n";
        out << "            float* const stg_ = stage_ + (it_ & 1) * (" << stage_words << " * " << B << ");\n";
        out << "            const int lane_ = threadIdx.x & 31, w0_ = threadIdx.x & ~31;\n";
        out << "            if (lane_ == 0) bulk_wait_read1();   // this warp's store from two tiles ago has drained its buffer\n";
        out << "            __syncwarp();\n";
        out << "            if (active_) {\n";
        long long woff = 0;
        for (const OutCluster& c : g.clusters) {
            long long W = c.stride / 4;
            out << "                {\n                    float rec_[" << W << "];\n";
            for (int k : c.outs) {
                Symbol& s = out_sym(k);
                out << "                    " << (s.out.derivs ? "wrd" : "wr") << "(rec_ + " << (s.out.offset - c.lo) / 4
                    << ", " << out_expr(k) << ");\n";
            }
            out << "                    stage_record<" << W << ">(stg_ + " << woff << " * " << B << ", threadIdx.x, rec_);\n";
            out << "                }\n";
            woff += W;
        }
        out << "            }\n";
        out << "            fence_async_smem();\n            __syncwarp();\n";
        out << "            const long long i0_ = tile_ * " << B << " + w0_;   // first point of this warp\n";
        out << "            long long cnt_ = L.npoints - i0_;\n";
        out << "            cnt_ = cnt_ < 0 ? 0 : (cnt_ < 32 ? cnt_ : 32);\n";
        woff = 0;
        for (const OutCluster& c : g.clusters) {
            long long W = c.stride / 4;
            out << "            if (cnt_ > 0) {\n";
            out << "                char* gp_ = (char*)L.output_base + " << c.lo << "LL + L.out_adjust[" << c.outs[0] << "] + "
                << c.stride << "LL * (i0_ + L.shadeindex_base);\n";
            out << "                const unsigned bytes_ = (unsigned)(cnt_ * " << c.stride << "LL);\n";
            out << "                const float* sp_ = stg_ + " << woff << " * " << B << " + w0_ * " << W << ";\n";
            out << "                if ((((unsigned long long)gp_ | bytes_) & 15ull) == 0) {\n";
            out << "                    if (lane_ == 0) bulk_store_nocommit(gp_, sp_, bytes_);\n";
            out << "                } else {\n";
            out << "                    for (unsigned w_ = lane_; w_ < bytes_ / 4; w_ += 32) ((float*)gp_)[w_] = sp_[w_];\n";
            out << "                }\n            }\n";
            woff += W;
        }
        out << "            if (lane_ == 0) bulk_commit();   // one bulk group per warp and tile\n";
        out << "        } else\n";
    }
    out << "        if (active_) {\n";
    for (size_t k = 0; k < g.outputs.size(); ++k) {
        Symbol& s = out_sym((int)k);
        out << "            " << (s.out.derivs ? "wrd" : "wr") << "(outp(L, sg, " << k << ", " << s.out.offset << "LL, "
            << s.out.stride << "LL), " << out_expr((int)k) << ");\n";
    }
    out << "        }\n";
    out << "    }\n";
    if (g.stage_ok)
        out << "    if (staged_ && (threadIdx.x & 31) == 0) bulk_wait_all();\n";
    out << "}\n";
    return out.str();
}

std::string
Gen::run_material(const std::string& ns)
{
    material_mode = true;
    int nlayers = (int)g.layers.size();
    if (nlayers > 32)
        throw std::runtime_error("B200 back end: more than 32 layers in a group is not supported yet");
    scan_messages();
    std::string gd = "struct GD {\n    unsigned ran;\n";
    for (int l = 0; l < nlayers; ++l) {
        if (g.layers[l].unused)
            continue;
        L  = &g.layers[l];
        li = l;
        for (Symbol& s : g.layers[l].m.syms)
            if (s.is_param()) {
                std::string arr = s.type.arraylen ? "[" + std::to_string(s.type.arraylen) + "]" : "";
                gd += "    " + ctype(s) + " L" + std::to_string(l) + "_" + ident(s.name) + arr + ";\n";
            }
    }
    gd += message_fields();
    gd += "};\n";
    for (int l = 0; l < nlayers; ++l)
        if (!g.layers[l].unused)
            o << "static __device__ __forceinline__ void layer_" << l << "(SG& sg, GD& gd, const B200Launch& L);\n";
    for (int l = 0; l < nlayers; ++l)
        if (!g.layers[l].unused)
            gen_layer(l);
    std::ostringstream out;
    out << "namespace " << ns << " {\n" << gd << o.str();
    out << "static __device__ OSLD_ENTRY_INLINE void entry(SG& sg)\n{\n    GD gd;\n    gd.ran = 0u;\n    B200Launch L;\n";
    if (!messages.empty())
        out << "    gd.msgset = 0u;\n";
    out << "    layer_" << (nlayers - 1) << "(sg, gd, L);\n}\n}  // namespace " << ns << "\n";
    return out.str();
}

}  // namespace

std::string
generate_cuda(Group& g)
{
    return Gen(g).run();
}

// All material groups of a scene + the wavefront integrator in one module.
std::string
generate_cuda_render(std::vector<Group*>& groups, bool has_background, RenderModuleInfo* info)
{
    std::string mats;
    bool color = false, glossy = false, mx = false, in_loop = false, media = false, sheen_ltc = false, thin = false;
    int ntex = 0;  // one texture table per module: each group's slots follow the previous group's
    int pool_words = 2, lobes = 1, adds = 0;
    for (size_t k = 0; k < groups.size(); ++k) {
        Group& g       = *groups[k];
        g.texture_base = ntex;
        g.textures.clear();
        g.pool_words_bound = 1;
        g.lobe_bound = g.closure_adds = 0;
        g.closure_in_loop             = false;
        g.closure_names.clear();
        mats += Gen(g).run_material("mat" + std::to_string(k));
        color |= g.uses_colorsystem;
        glossy |= g.uses_glossy_lobes;
        mx |= g.uses_mx_lobes;
        media |= g.uses_media;
        sheen_ltc |= g.uses_sheen_ltc;
        thin |= g.uses_thinlayer;
        in_loop |= g.closure_in_loop;
        pool_words = std::max(pool_words, g.pool_words_bound);
        lobes      = std::max(lobes, g.lobe_bound);
        adds       = std::max(adds, g.closure_adds);
        ntex += (int)g.textures.size();
    }
    // the integrator's per-thread closure arena, lobe array and tree-walk stack are sized from
    // what the scene's materials can build, capped at the reference's own limits (1 KB pool,
    // 8 lobes, 16-deep stack: render_state.h:27, shading.h:319, shading.cpp:1456)
    RenderModuleInfo mi;
    mi.pool_words    = in_loop ? 256 : std::min(256, pool_words);
    mi.max_lobes     = in_loop ? 8 : std::min(8, lobes);
    mi.closure_stack = in_loop ? 16 : std::min(16, adds + 1);
    mi.pool_in_smem  = mi.pool_words <= 64;
    if (info)
        *info = mi;
    std::ostringstream out;
    out << "// generated by libosl_b200: render module with " << groups.size() << " material group(s)\n";
    out << "#define OSLD_POOL_WORDS " << mi.pool_words << "\n#define OSLD_MAX_LOBES " << mi.max_lobes
        << "\n#define OSLD_CLOSURE_STACK " << mi.closure_stack << "\n";
    if (mi.pool_in_smem)
        out << "#define OSLD_POOL_SMEM 1\n";
    out << "#include \"osl_b200_device.cuh\"\n#include \"osl_b200_closure.cuh\"\n#include \"osl_b200_sg.cuh\"\n";
    if (ntex)
        out << "#include \"osl_b200_texture.cuh\"\nextern \"C\" __device__ osld::TexDesc osl_tex_[" << ntex << "];\n";
    if (color)  // one colour system per module: the shading system's, i.e. the first group's
        out << "#include \"osl_b200_color.cuh\"\n" << colorsystem_cuda_definition(groups[0]->colorspace);
    out << "using namespace osld;\nstruct B200Launch { int unused_; };\n" << mats;
    out << "static __device__ __forceinline__ void osl_execute_shader(int shaderID, SG& sg)\n{\n    switch (shaderID) {\n";
    for (size_t k = 0; k < groups.size(); ++k)
        out << "    case " << k << ": mat" << k << "::entry(sg); break;\n";
    out << "    default: break;\n    }\n}\n";
    // the integrator is specialised to the lobes the scene's materials can create
    if (glossy)
        out << "#define OSLD_GLOSSY_LOBES 1\n";
    if (mx)
        out << "#define OSLD_MX_LOBES 1\n";
    if (media)
        out << "#define OSLD_HAS_MEDIA 1\n";
    if (sheen_ltc)
        out << "#define OSLD_SHEEN_LTC 1\n";
    if (thin)
        out << "#define OSLD_THINLAYER 1\n";
    mi.uses_mx_lobes = mx;
    mi.uses_media    = media;
    mi.uses_luts     = mx || sheen_ltc;
    if (info) {
        info->uses_luts     = mx || sheen_ltc;
        info->uses_mx_lobes = mx;
        info->uses_media    = media;
    }
    if (has_background)
        out << "#define OSLD_HAS_BACKGROUND 1\n";
    out << "#include \"osl_b200_render.cuh\"\n";
    return out.str();
}

}  // namespace oslb200
