// oslexec_b200.h — C++ host-side mirror of the reference's ShadingSystem API
// for the one path this back end accelerates (product code, header only).
//
// Same names, argument meaning and error behaviour (bool returns, messages via
// geterror()) as src/include/OSL/oslexec.h:170-1165 for: attribute(),
// LoadMemoryCompiledShader(), ShaderGroupBegin/Parameter/Shader/
// ConnectShaders/ShaderGroupEnd, add_symlocs(), optimize_group() and a
// batched executor whose execute() takes a whole SoA batch.  Everything is
// forwarded to the C ABI in osl_b200.h; no OIIO / Imath / LLVM needed.
#pragma once
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../osl_b200.h"

namespace OSL_B200 {

// TypeDesc-lite: enough of OIIO::TypeDesc for Parameter() and SymLocationDesc
struct TypeDesc {
    enum BASETYPE { UNKNOWN, INT, FLOAT, STRING };
    enum AGGREGATE { SCALAR = 1, VEC3 = 3, MATRIX44 = 16 };
    BASETYPE basetype = UNKNOWN;
    int aggregate     = SCALAR;
    int arraylen      = 0;
    constexpr TypeDesc() {}
    constexpr TypeDesc(BASETYPE b, int agg = SCALAR, int arr = 0) : basetype(b), aggregate(agg), arraylen(arr) {}
    size_t numelements() const { return arraylen > 0 ? (size_t)arraylen : 1; }
    size_t size() const { return numelements() * aggregate * (basetype == STRING ? sizeof(char*) : 4); }
};
static constexpr TypeDesc TypeInt(TypeDesc::INT), TypeFloat(TypeDesc::FLOAT), TypeString(TypeDesc::STRING),
    TypeColor(TypeDesc::FLOAT, TypeDesc::VEC3), TypePoint(TypeDesc::FLOAT, TypeDesc::VEC3),
    TypeVector(TypeDesc::FLOAT, TypeDesc::VEC3), TypeNormal(TypeDesc::FLOAT, TypeDesc::VEC3);

enum class SymArena { Unknown, Absolute, Heap, Outputs, UserData };

// oslexec.h:69-105
struct SymLocationDesc {
    std::string name;
    TypeDesc type;
    long long offset = -1, stride = 0;
    SymArena arena = SymArena::Outputs;
    bool derivs    = false;
    SymLocationDesc() {}
    SymLocationDesc(const std::string& name, TypeDesc type, bool derivs = false, SymArena arena = SymArena::Outputs,
                    long long offset = -1, long long stride = 0)
        : name(name), type(type), offset(offset), stride(stride), arena(arena), derivs(derivs)
    {
    }
};

class ShaderGroup {
public:
    ~ShaderGroup()
    {
        if (handle)
            b200_group_destroy(handle);
    }
    std::string name;
    struct Layer {
        std::string oso, layername;
        struct P {
            std::string name;
            int type;
            std::vector<int> i;
            std::vector<float> f;
            std::vector<std::string> s;
        };
        std::vector<P> params;
    };
    std::vector<Layer> layers;
    std::vector<Layer::P> pending;  // Parameter() calls apply to the next Shader()
    struct Conn {
        std::string sl, sp, dl, dp;
    };
    std::vector<Conn> conns;
    std::vector<SymLocationDesc> symlocs;
    b200_group* handle = nullptr;
    bool ended         = false;
};
typedef std::shared_ptr<ShaderGroup> ShaderGroupRef;

class ShadingSystem {
public:
    // attribute("searchpath:shader", dir), attribute("llvm_jit_fma", 0|1),
    // attribute("b200_block", N)   (oslexec.h:177-319 naming)
    bool attribute(const std::string& name, const std::string& val)
    {
        m_sattr[name] = val;
        return true;
    }
    bool attribute(const std::string& name, int val)
    {
        m_iattr[name] = val;
        return true;
    }
    const std::string& geterror() const { return m_err; }

    // oslexec.h: LoadMemoryCompiledShader(shadername, buffer)
    bool LoadMemoryCompiledShader(const std::string& shadername, const std::string& oso)
    {
        m_mem[shadername] = oso;
        return true;
    }
    ShaderGroupRef ShaderGroupBegin(const std::string& groupname = "")
    {
        auto g  = std::make_shared<ShaderGroup>();
        g->name = groupname;
        return g;
    }
    bool Parameter(ShaderGroup& g, const std::string& name, TypeDesc t, const void* val)
    {
        ShaderGroup::Layer::P p;
        p.name   = name;
        size_t n = t.numelements() * t.aggregate;
        if (t.basetype == TypeDesc::INT) {
            p.type = 0;
            p.i.assign((const int*)val, (const int*)val + n);
        } else if (t.basetype == TypeDesc::FLOAT) {
            p.type = 1;
            p.f.assign((const float*)val, (const float*)val + n);
        } else if (t.basetype == TypeDesc::STRING) {
            p.type = 2;
            for (size_t k = 0; k < n; ++k)
                p.s.push_back(((const char* const*)val)[k]);
        } else
            return error("Parameter: unknown type for '" + name + "'");
        g.pending.push_back(p);
        return true;
    }
    // oslexec.h:723 — usage is accepted and ignored like the reference does for "surface"/"shader"
    bool Shader(ShaderGroup& g, const std::string& /*shaderusage*/, const std::string& shadername,
                const std::string& layername)
    {
        ShaderGroup::Layer l;
        auto it = m_mem.find(shadername);
        if (it != m_mem.end())
            l.oso = it->second;
        else {
            std::string dir = m_sattr.count("searchpath:shader") ? m_sattr["searchpath:shader"] : ".";
            std::istringstream dirs(dir);
            std::string d;
            while (l.oso.empty() && std::getline(dirs, d, ':')) {
                std::ifstream f(d + "/" + shadername + ".oso");
                if (f) {
                    std::stringstream ss;
                    ss << f.rdbuf();
                    l.oso = ss.str();
                }
            }
            if (l.oso.empty())
                return error("Could not find shader \"" + shadername + "\"");
        }
        l.layername = layername.empty() ? shadername : layername;
        l.params.swap(g.pending);
        g.layers.push_back(std::move(l));
        return true;
    }
    bool ConnectShaders(ShaderGroup& g, const std::string& srclayer, const std::string& srcparam,
                        const std::string& dstlayer, const std::string& dstparam)
    {
        g.conns.push_back({ srclayer, srcparam, dstlayer, dstparam });
        return true;
    }
    // oslexec.h:1075 — only SymArena::Outputs is meaningful for this back end
    void add_symlocs(ShaderGroup* g, const SymLocationDesc* locs, size_t n)
    {
        for (size_t i = 0; i < n; ++i)
            g->symlocs.push_back(locs[i]);
    }
    bool ShaderGroupEnd(ShaderGroup& g)
    {
        g.ended = true;
        return true;
    }
    // oslexec.h:1093 — JIT the group (generation + NVRTC; no GPU needed)
    bool optimize_group(ShaderGroup* g)
    {
        if (g->handle)
            return true;
        std::vector<b200_layer> L(g->layers.size());
        std::vector<std::vector<b200_param>> P(g->layers.size());
        std::vector<std::vector<const char*>> S;
        for (size_t i = 0; i < g->layers.size(); ++i) {
            auto& l = g->layers[i];
            for (auto& p : l.params) {
                b200_param bp { p.name.c_str(), p.type, 0, nullptr };
                if (p.type == 0) { bp.nvalues = (int)p.i.size(); bp.values = p.i.data(); }
                else if (p.type == 1) { bp.nvalues = (int)p.f.size(); bp.values = p.f.data(); }
                else {
                    S.emplace_back();
                    for (auto& s : p.s) S.back().push_back(s.c_str());
                    bp.nvalues = (int)p.s.size();
                    bp.values  = S.back().data();
                }
                P[i].push_back(bp);
            }
            L[i] = b200_layer { l.oso.c_str(), l.layername.c_str(), (int)P[i].size(), P[i].data() };
        }
        std::vector<b200_connection> C;
        for (auto& c : g->conns)
            C.push_back({ c.sl.c_str(), c.sp.c_str(), c.dl.c_str(), c.dp.c_str() });
        std::vector<b200_symloc> O;
        for (auto& s : g->symlocs)
            if (s.arena == SymArena::Outputs)
                O.push_back({ s.name.c_str(), s.offset, s.stride, s.derivs ? 1 : 0 });
        std::string opts = std::string("fma=") + (m_iattr.count("llvm_jit_fma") && !m_iattr["llvm_jit_fma"] ? "0" : "1");
        if (m_iattr.count("b200_block"))
            opts += ",block=" + std::to_string(m_iattr["b200_block"]);
        b200_group_desc d { g->name.c_str(), (int)L.size(), L.data(), (int)C.size(), C.data(),
                            (int)O.size(), O.data(), opts.c_str() };
        if (b200_group_compile(&d, &g->handle) != B200_OK)
            return error(b200_last_error());
        return true;
    }

    // The batched executor of oslexec.h:982-1033 with batch = the whole range.
    class BatchedExecutor {
    public:
        explicit BatchedExecutor(ShadingSystem& ss) : m_ss(ss) {}
        bool jit_group(ShaderGroup* g) { return m_ss.optimize_group(g); }
        // device pointers (sg planes, shadeindex, output arena); asynchronous on `stream`
        bool execute(ShaderGroup& g, long long batch_size, const int* wide_shadeindex, const b200_globals& bsg,
                     void* userdata_base, void* output_base, int device = 0, void* stream = nullptr)
        {
            if (!g.handle && !m_ss.optimize_group(&g))
                return false;
            if (b200_group_execute(g.handle, device, stream, batch_size, &bsg, wide_shadeindex, userdata_base,
                                   output_base) != B200_OK)
                return m_ss.error(b200_last_error());
            return true;
        }
        // host pointers: the drop-in for ShadingSystem::execute over host ShaderGlobals
        bool execute_host(ShaderGroup& g, long long batch_size, const b200_globals& bsg, void* output_base,
                          int device = 0)
        {
            if (!g.handle && !m_ss.optimize_group(&g))
                return false;
            if (b200_group_execute_host(g.handle, device, batch_size, &bsg, output_base) != B200_OK)
                return m_ss.error(b200_last_error());
            return true;
        }

    private:
        ShadingSystem& m_ss;
    };
    BatchedExecutor batched() { return BatchedExecutor(*this); }

private:
    bool error(const std::string& m)
    {
        m_err = m;
        return false;
    }
    std::map<std::string, std::string> m_sattr, m_mem;
    std::map<std::string, int> m_iattr;
    std::string m_err;
};

}  // namespace OSL_B200
