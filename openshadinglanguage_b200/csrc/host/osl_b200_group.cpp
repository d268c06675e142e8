// osl_b200_group.cpp — .oso reader, group assembly and the three optimizer
// facts the B200 code generator consumes (product code).
//
// Reference behaviour restated here:
//   .oso grammar                  src/liboslexec/osogram.y:88-318, osolex.l
//   Shader/Parameter/Connect      src/liboslexec/shadingsys.cpp:3039-3300
//   unused layers                 src/liboslexec/llvm_instance.cpp:2301-2318
//   run_lazily()                  src/liboslexec/oslexec_pvt.h:1344-1372
//   derivative propagation        src/liboslexec/runtimeoptimize.cpp:2542-2720, 3223-3231
#include "osl_b200_group.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace oslb200 {

namespace {

std::vector<std::string> tokenize(const std::string& line)
{
    // tokens: "quoted strings", %hint{...}, %hint, or runs of non-space
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        if (isspace((unsigned char)line[i])) {
            ++i;
            continue;
        }
        size_t j = i;
        if (line[i] == '"') {
            ++j;
            while (j < n && line[j] != '"')
                j += (line[j] == '\\') ? 2 : 1;
            ++j;
        } else if (line[i] == '%') {
            while (j < n && !isspace((unsigned char)line[j]) && line[j] != '{')
                ++j;
            if (j < n && line[j] == '{') {
                // hint payload may contain quoted strings with braces/spaces
                bool inq = false;
                while (j < n) {
                    if (line[j] == '"' && line[j - 1] != '\\')
                        inq = !inq;
                    if (line[j] == '}' && !inq) {
                        ++j;
                        break;
                    }
                    ++j;
                }
            }
        } else {
            while (j < n && !isspace((unsigned char)line[j]))
                ++j;
        }
        out.push_back(line.substr(i, j - i));
        i = j;
    }
    return out;
}

std::string unescape(const std::string& s)
{
    std::string r;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '\\' && i + 1 < s.size()) {
            char c = s[++i];
            r += c == 'n' ? '\n' : c == 't' ? '\t' : c;
        } else
            r += s[i];
    }
    return r;
}

bool parse_base(const std::string& s, Base& b)
{
    static const std::map<std::string, Base> tbl
        = { { "int", Base::Int },       { "float", Base::Float },   { "string", Base::String },
            { "color", Base::Color },   { "point", Base::Point },   { "vector", Base::Vector },
            { "normal", Base::Normal }, { "matrix", Base::Matrix }, { "void", Base::Void } };
    auto it = tbl.find(s);
    if (it == tbl.end())
        return false;
    b = it->second;
    return true;
}

bool is_integer(const std::string& t)
{
    size_t i = (t[0] == '-') ? 1 : 0;
    if (i >= t.size())
        return false;
    for (; i < t.size(); ++i)
        if (!isdigit((unsigned char)t[i]))
            return false;
    return true;
}

}  // namespace

Master
parse_oso(const std::string& text)
{
    Master m;
    std::istringstream in(text);
    std::string line;
    int lineno = 0;
    auto fail  = [&](const std::string& msg) {
        throw std::runtime_error("oso parse error, line " + std::to_string(lineno) + ": " + msg);
    };
    if (!std::getline(in, line) || line.compare(0, 19, "OpenShadingLanguage") != 0)
        throw std::runtime_error("not an OSO file (missing 'OpenShadingLanguage' header)");
    ++lineno;
    bool have_shader = false, in_code = false;
    std::string method;
    int begin = 0;
    while (std::getline(in, line)) {
        ++lineno;
        size_t p = line.find_first_not_of(" \t\r");
        if (p == std::string::npos || line[p] == '#')
            continue;
        if (line.compare(0, 20, "%preprocessed_source") == 0)
            break;
        std::vector<std::string> t = tokenize(line);
        if (t.empty())
            continue;
        if (!have_shader) {
            if (t.size() < 2)
                fail("expected '<shadertype> <name>'");
            m.shadertype = t[0];
            m.shadername = t[1];
            have_shader  = true;
            continue;
        }
        if (t[0] == "code") {
            if (t.size() < 2)
                fail("code section without a name");
            if (in_code)
                m.methods[method] = { begin, (int)m.ops.size() };
            in_code = true;
            method  = t[1];
            begin   = (int)m.ops.size();
            continue;
        }
        if (!in_code) {
            // symbol:  symtype type name [defaults] [hints]
            Symbol s;
            static const std::map<std::string, SymType> st
                = { { "param", SymType::Param },   { "oparam", SymType::OutputParam },
                    { "local", SymType::Local },   { "temp", SymType::Temp },
                    { "global", SymType::Global }, { "const", SymType::Const } };
            auto it = st.find(t[0]);
            if (it == st.end())
                fail("unknown symbol type '" + t[0] + "'");
            s.symtype = it->second;
            size_t i  = 1;
            if (i >= t.size())
                fail("truncated symbol line");
            std::string tn = t[i++];
            if (tn == "closure") {
                s.type.base = Base::Closure;
                ++i;  // "color"
            } else if (tn == "struct") {
                fail("struct symbols are not supported by the B200 back end");
            } else {
                size_t br = tn.find('[');
                if (br != std::string::npos) {
                    s.type.arraylen = atoi(tn.c_str() + br + 1);
                    s.unsized       = tn.compare(br, 2, "[]") == 0;   // "type[]": sized below / by the instance
                    tn              = tn.substr(0, br);
                }
                if (!parse_base(tn, s.type.base))
                    fail("unknown type '" + tn + "'");
            }
            if (i >= t.size())
                fail("symbol without a name");
            s.name    = t[i++];
            size_t br = s.name.find('[');
            if (br != std::string::npos) {
                s.type.arraylen = atoi(s.name.c_str() + br + 1);
                s.name          = s.name.substr(0, br);
            }
            for (; i < t.size(); ++i) {
                const std::string& v = t[i];
                if (v[0] == '%') {
                    if (v == "%initexpr")
                        s.initexpr = true;
                    else if (v == "%meta{int,lockgeom,0}")
                        s.interpolated = true;   // [[ int lockgeom = 0 ]]: bound to userdata per point
                    continue;
                }
                if (v[0] == '"')
                    s.svals.push_back(unescape(v.substr(1, v.size() - 2)));
                else if (s.type.base == Base::Int)
                    s.ivals.push_back(atoi(v.c_str()));
                else if (s.type.base == Base::String)
                    s.svals.push_back(v);
                else
                    s.fvals.push_back(strtof(v.c_str(), nullptr));
            }
            if (s.unsized) {
                // an unsized array parameter is as long as its default list until an instance value or a
                // connection gives it another length (instance.cpp:250-330)
                size_t nvals    = s.type.base == Base::Int ? s.ivals.size()
                                  : (s.type.base == Base::String ? s.svals.size() : s.fvals.size());
                TypeSpec elem   = s.type;
                elem.arraylen   = 0;
                s.type.arraylen = std::max(1, (int)(nvals / (size_t)std::max(1, elem.ncomp())));
            }
            m.byname[s.name] = (int)m.syms.size();
            m.syms.push_back(std::move(s));
            continue;
        }
        // op line
        if (t[0] == "end")
            continue;
        Opcode op;
        op.name = t[0];
        bool have_rw = false;
        for (size_t i = 1; i < t.size(); ++i) {
            const std::string& v = t[i];
            if (v.compare(0, 6, "%argrw") == 0) {
                size_t a = v.find('"'), b = v.rfind('"');
                op.rw   = v.substr(a + 1, b - a - 1);
                have_rw = true;
            } else if (v.compare(0, 10, "%argderivs") == 0) {
                size_t a = v.find('{');
                std::string body = v.substr(a + 1, v.size() - a - 2);
                std::istringstream bs(body);
                std::string num;
                while (std::getline(bs, num, ','))
                    if (!num.empty())
                        op.derivs.push_back(atoi(num.c_str()));
            } else if (v[0] == '%') {
                // %filename %line etc: not needed
            } else if (is_integer(v)) {
                op.jumps.push_back(atoi(v.c_str()));
            } else {
                int si = m.find(v);
                if (si < 0)
                    fail("op '" + op.name + "' references unknown symbol '" + v + "'");
                op.args.push_back(si);
            }
        }
        if (!have_rw) {
            op.rw.assign(op.args.size(), 'r');
            if (!op.args.empty())
                op.rw[0] = 'w';
        }
        if (op.rw.size() != op.args.size())
            fail("argrw length does not match argument count for op '" + op.name + "'");
        m.ops.push_back(std::move(op));
    }
    if (!have_shader)
        throw std::runtime_error("oso parse error: no shader declaration");
    if (in_code)
        m.methods[method] = { begin, (int)m.ops.size() };
    return m;
}

int
Group::layer_index(const std::string& n) const
{
    for (size_t i = 0; i < layers.size(); ++i)
        if (layers[i].layername == n)
            return (int)i;
    return -1;
}

int
Group::intern(const std::string& s)
{
    for (size_t i = 0; i < strings.size(); ++i)
        if (strings[i] == s)
            return (int)i;
    strings.push_back(s);
    return (int)strings.size() - 1;
}

void
Group::add_layer(const std::string& oso_text, const std::string& layername,
                 const std::vector<ParamValue>& params)
{
    Layer l;
    l.m         = parse_oso(oso_text);
    l.layername = layername;
    for (const ParamValue& pv : params) {
        int si = l.m.find(pv.name);
        if (si < 0 || !l.m.syms[si].is_param())
            throw std::runtime_error("Parameter: shader '" + l.m.shadername + "' has no parameter '"
                                     + pv.name + "'");
        Symbol& s = l.m.syms[si];
        // int values may initialise float params and vice versa (Parameter()
        // accepts a TypeDesc; shadingsys.cpp:2880-3035 checks assignability)
        if (s.type.base == Base::Int) {
            s.ivals = pv.ivals;
            if (s.ivals.empty())
                for (float f : pv.fvals)
                    s.ivals.push_back((int)f);
        } else if (s.type.base == Base::String) {
            s.svals = pv.svals;
        } else {
            s.fvals = pv.fvals;
            if (s.fvals.empty())
                for (int i : pv.ivals)
                    s.fvals.push_back((float)i);
            if (s.type.is_triple() && s.fvals.size() == 1)
                s.fvals.assign(3, s.fvals[0]);
        }
        if (s.unsized) {
            size_t nvals    = s.type.base == Base::Int ? s.ivals.size()
                              : (s.type.base == Base::String ? s.svals.size() : s.fvals.size());
            TypeSpec elem   = s.type;
            elem.arraylen   = 0;
            s.type.arraylen = std::max(1, (int)(nvals / (size_t)std::max(1, elem.ncomp())));
        }
        s.initexpr = false;
    }
    layers.push_back(std::move(l));
}

void
Group::connect(const std::string& sl, const std::string& sp, const std::string& dl,
               const std::string& dp)
{
    int si = layer_index(sl), di = layer_index(dl);
    if (si < 0 || di < 0)
        throw std::runtime_error("ConnectShaders: unknown layer '" + (si < 0 ? sl : dl) + "'");
    if (si >= di)
        throw std::runtime_error("ConnectShaders: source layer must precede destination layer");
    int ss = layers[si].m.find(sp), ds = layers[di].m.find(dp);
    if (ss < 0 || !layers[si].m.syms[ss].is_param())
        throw std::runtime_error("ConnectShaders: layer '" + sl + "' has no parameter '" + sp + "'");
    if (ds < 0 || !layers[di].m.syms[ds].is_param())
        throw std::runtime_error("ConnectShaders: layer '" + dl + "' has no parameter '" + dp + "'");
    if (layers[di].m.syms[ds].unsized && layers[si].m.syms[ss].type.arraylen)
        layers[di].m.syms[ds].type.arraylen = layers[si].m.syms[ss].type.arraylen;
    layers[si].m.syms[ss].connected_down = true;
    layers[di].m.syms[ds].conn_layer     = si;
    layers[di].m.syms[ds].conn_sym       = ss;
    connections.push_back({ si, ss, di, ds });
}

void
Group::add_output(const std::string& name, long long offset, long long stride, bool derivs)
{
    int li = -1, si = -1;
    size_t dot = name.find('.');
    if (dot != std::string::npos) {
        li = layer_index(name.substr(0, dot));
        if (li >= 0)
            si = layers[li].m.find(name.substr(dot + 1));
    } else {
        for (int l = (int)layers.size() - 1; l >= 0 && si < 0; --l) {
            int s = layers[l].m.find(name);
            if (s >= 0 && layers[l].m.syms[s].is_param()) {
                li = l;
                si = s;
            }
        }
    }
    if (si < 0 && dot == std::string::npos && !layers.empty()) {
        // A ShaderGlobals field the entry layer uses: execute() hands the globals back to the renderer
        // as the shaders left them (ShaderGlobals is in / out, oslexec.h:833; a displacement shader's
        // "P += ..." is read back this way, simpleraytracer.cpp:1365-1384).  Here the renderer asks for
        // the field by name like for any other output.
        int l = (int)layers.size() - 1;
        int s = layers[l].m.find(name);
        if (s >= 0 && layers[l].m.syms[s].symtype == SymType::Global) {
            li = l;
            si = s;
        }
    }
    if (li < 0 || si < 0)
        throw std::runtime_error("renderer output '" + name + "' not found in group");
    Symbol& s    = layers[li].m.syms[si];
    s.out.placed = true;
    s.out.offset = offset;
    s.out.stride = stride;
    s.out.derivs = derivs;
    outputs.push_back({ li, si });
}

static bool
deriv_global(const std::string& n)
{
    return n == "P" || n == "I" || n == "u" || n == "v" || n == "Ps";
}

static void
track_derivs(Layer& l)
{
    Master& m = l.m;
    size_t ns = m.syms.size();
    std::vector<std::set<int>> deps(ns);
    std::vector<int> need;
    for (const Opcode& op : m.ops) {
        for (size_t w = 0; w < op.args.size(); ++w) {
            if (!op.writes((int)w))
                continue;
            for (size_t r = 0; r < op.args.size(); ++r)
                if (op.reads((int)r) && !m.syms[op.args[r]].is_const())
                    deps[op.args[w]].insert(op.args[r]);
        }
        for (int ai : op.derivs) {
            if (ai < 0 || ai >= (int)op.args.size())
                continue;
            const Symbol& s = m.syms[op.args[ai]];
            if (s.is_const() || !s.type.is_float_based() || s.type.base == Base::Matrix)
                continue;
            if (s.symtype == SymType::Global && !deriv_global(s.name))
                continue;
            need.push_back(op.args[ai]);
        }
    }
    for (size_t i = 0; i < ns; ++i) {
        Symbol& s = m.syms[i];
        if (s.symtype == SymType::Global && s.written && s.type.is_float_based() && s.name != "N")
            s.has_derivs = true;
        if (s.out.placed && s.out.derivs && s.written && s.type.is_float_based())
            s.has_derivs = true;
        if (s.has_derivs)
            need.push_back((int)i);
    }
    std::vector<char> seen(ns, 0);
    while (!need.empty()) {
        int i = need.back();
        need.pop_back();
        if (seen[i])
            continue;
        seen[i]   = 1;
        Symbol& s = m.syms[i];
        if (s.type.is_float_based() && s.type.base != Base::Matrix && !s.is_const())
            s.has_derivs = true;
        for (int r : deps[i])
            need.push_back(r);
    }
    for (Symbol& s : m.syms) {
        if (s.symtype == SymType::Global && !deriv_global(s.name))
            s.has_derivs = false;
        if (!s.type.is_float_based() || s.type.base == Base::Matrix)
            s.has_derivs = false;
    }
}

void
Group::finalize()
{
    int n = (int)layers.size();
    if (n == 0)
        throw std::runtime_error("ShaderGroupEnd: group has no layers");
    for (Layer& l : layers)
        for (const Opcode& op : l.m.ops)
            for (size_t a = 0; a < op.args.size(); ++a)
                if (op.writes((int)a))
                    l.m.syms[op.args[a]].written = true;
    for (int i = 0; i < n; ++i) {
        Layer& l    = layers[i];
        bool has_out = false, has_down = false;
        for (const Symbol& s : l.m.syms) {
            has_out |= s.out.placed;
            has_down |= s.connected_down;
        }
        l.unused = (i != n - 1) && !has_out && !has_down;
        l.lazy   = (i != n - 1) && !has_out;
    }
    // output clusters (records that interleave) and whether they can be staged
    clusters.clear();
    for (size_t k = 0; k < outputs.size(); ++k) {
        const Symbol& s = layers[outputs[k].first].m.syms[outputs[k].second];
        long long size  = 4LL * s.type.ncomp() * (s.out.derivs ? 3 : 1);
        if (s.type.arraylen || s.out.stride < size || (s.out.offset & 3) || (s.out.stride & 3))
            throw std::runtime_error("renderer output '" + s.name
                                     + "': arrays, strides smaller than the value, or unaligned "
                                       "offsets/strides are not supported");
        bool placed = false;
        for (OutCluster& c : clusters) {
            long long lo = std::min(c.lo, s.out.offset), hi = std::max(c.hi, s.out.offset + size);
            if (c.stride == s.out.stride && hi - lo <= c.stride) {
                c.lo = lo;
                c.hi = hi;
                c.outs.push_back((int)k);
                placed = true;
                break;
            }
        }
        if (!placed) {
            OutCluster c;
            c.stride = s.out.stride;
            c.lo     = s.out.offset;
            c.hi     = s.out.offset + size;
            c.outs.push_back((int)k);
            clusters.push_back(c);
        }
    }
    long long stage_bytes = 0;
    stage_ok              = !clusters.empty();
    stage_record_bytes    = 0;
    for (OutCluster& c : clusters) {
        // dense: the fields cover [lo, lo+stride) with no gap and no overlap
        std::vector<std::pair<long long, long long>> spans;
        for (int k : c.outs) {
            const Symbol& s = layers[outputs[k].first].m.syms[outputs[k].second];
            spans.push_back({ s.out.offset, s.out.offset + 4LL * s.type.ncomp() * (s.out.derivs ? 3 : 1) });
        }
        std::sort(spans.begin(), spans.end());
        long long at = c.lo;
        bool ok      = true;
        for (auto& sp : spans) {
            ok &= (sp.first == at);
            at = sp.second;
        }
        c.dense = ok && at == c.lo + c.stride;
        stage_ok &= c.dense;
        stage_bytes += c.stride * block;
        stage_record_bytes = std::max(stage_record_bytes, c.stride);
    }
    if (2 * stage_bytes > 48 * 1024)  // double-buffered staging must fit static shared memory
        stage_ok = false;
    for (int i = n - 1; i >= 0; --i) {
        track_derivs(layers[i]);
        for (const Symbol& s : layers[i].m.syms)
            if (s.conn_layer >= 0 && s.has_derivs)
                layers[s.conn_layer].m.syms[s.conn_sym].has_derivs = true;
    }
}

}  // namespace oslb200
