// OSL/oslexec.h — the reference's execution API (namespace OSL) over the B200 back end.
//
// Drop-in mirror of the part of src/include/OSL/oslexec.h (+ shaderglobals.h,
// batched_shaderglobals.h, wide.h, rendererservices.h) that a renderer uses to build a shader
// group and execute it: same names, same argument order and meaning, same error behaviour
// (bool returns, messages through geterror()).  A renderer written against the reference
// compiles against this header and links libosl_b200.so; every call below ends in the C ABI
// of include/osl_b200.h.  No OIIO / Imath / LLVM is needed: the few value types the
// signatures mention (ustring, TypeDesc, Vec3, Matrix44) are small stand-ins with the
// reference's layout.
//
//   ShadingSystem(RendererServices*, TextureSystem*, ErrorHandler*)      oslexec.h:172
//   attribute / getattribute                                             :324-596
//   ShaderGroupBegin (incl. the serialized "param ...; shader ...; connect ...;" form),
//   Parameter, Shader, ConnectShaders, ShaderGroupEnd, ReParameter       :634-760
//   create_thread_info / get_context / release_context                   :806-824
//   execute / execute_init / execute_layer / execute_cleanup             :833-918
//   batched<W>().jit_group / jit_all_groups / execute                    :982-1033
//   find_symbol / symbol_typedesc / symbol_address                       :956-971
//   register_closure / query_closure                                     :1041-1048
//   add_symlocs / find_symloc, SymLocationDesc, SymArena                 :69-105, 1075
//   optimize_group, raytype_bit                                          :1093, 1061
//   RendererServices::get_matrix / get_inverse_matrix / get_userdata / supports
//                                                                        rendererservices.h:91-602
//   ShaderGlobals, BatchedShaderGlobals<W>, Block<T,W>, Wide<T,W>        shaderglobals.h:55-146,
//                                                     batched_shaderglobals.h:21-195, wide.h:210-444
// On top of the reference API: BatchedExecutor<W>::execute(ctx, group, npoints, const b200_globals&,
// ...) runs a batch of any size in one launch (the reference's W-lane call works, but a GPU
// wants the whole tile), and ShadingSystem::b200_device() picks the GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <string_view>
#include <vector>

#include "../osl_b200.h"

#define OSL_B200_BACKEND 1
#define OSL_USE_BATCHED 1

namespace OSL {

using string_view = std::string_view;

// interned string with pointer identity (OIIO::ustring); ustringhash is the same thing here
class ustring {
public:
    ustring() : m_s(nullptr) {}
    ustring(const char* s) : m_s(s ? intern(s) : nullptr) {}
    ustring(const std::string& s) : m_s(intern(s)) {}
    ustring(string_view s) : m_s(intern(std::string(s))) {}
    const char* c_str() const { return m_s ? m_s->c_str() : ""; }
    const std::string& string() const
    {
        static const std::string empty;
        return m_s ? *m_s : empty;
    }
    bool empty() const { return !m_s || m_s->empty(); }
    size_t hash() const { return std::hash<std::string>()(string()); }
    bool operator==(const ustring& o) const { return m_s == o.m_s || string() == o.string(); }
    bool operator!=(const ustring& o) const { return !(*this == o); }
    bool operator<(const ustring& o) const { return string() < o.string(); }
    operator string_view() const { return string(); }

private:
    static const std::string* intern(const std::string& s)
    {
        static std::mutex mu;
        static std::set<std::string> table;
        std::lock_guard<std::mutex> lk(mu);
        return &*table.insert(s).first;
    }
    const std::string* m_s;
};
typedef ustring ustringhash;

// OIIO::TypeDesc, as far as Parameter() / SymLocationDesc / get_userdata need it
struct TypeDesc {
    enum BASETYPE { UNKNOWN, INT, FLOAT, STRING, PTR };
    enum AGGREGATE { SCALAR = 1, VEC3 = 3, MATRIX44 = 16 };
    enum VECSEMANTICS { NOSEMANTICS, COLOR, POINT, VECTOR, NORMAL };
    unsigned char basetype = UNKNOWN, aggregate = SCALAR, vecsemantics = NOSEMANTICS;
    int arraylen = 0;
    constexpr TypeDesc() {}
    constexpr TypeDesc(BASETYPE b, AGGREGATE agg = SCALAR, VECSEMANTICS sem = NOSEMANTICS, int arr = 0)
        : basetype(b), aggregate(agg), vecsemantics(sem), arraylen(arr)
    {
    }
    constexpr TypeDesc(BASETYPE b, int arr) : basetype(b), arraylen(arr) {}
    size_t numelements() const { return arraylen > 0 ? (size_t)arraylen : 1; }
    size_t basesize() const { return basetype == STRING || basetype == PTR ? sizeof(void*) : 4; }
    size_t size() const { return numelements() * aggregate * basesize(); }
    bool operator==(const TypeDesc& o) const
    {
        return basetype == o.basetype && aggregate == o.aggregate && arraylen == o.arraylen;
    }
    bool operator!=(const TypeDesc& o) const { return !(*this == o); }
};
static constexpr TypeDesc TypeUnknown, TypeInt(TypeDesc::INT), TypeFloat(TypeDesc::FLOAT), TypeString(TypeDesc::STRING),
    TypeColor(TypeDesc::FLOAT, TypeDesc::VEC3, TypeDesc::COLOR), TypePoint(TypeDesc::FLOAT, TypeDesc::VEC3, TypeDesc::POINT),
    TypeVector(TypeDesc::FLOAT, TypeDesc::VEC3, TypeDesc::VECTOR),
    TypeNormal(TypeDesc::FLOAT, TypeDesc::VEC3, TypeDesc::NORMAL), TypeMatrix(TypeDesc::FLOAT, TypeDesc::MATRIX44);

struct Vec3 {   // Imath::V3f layout
    float x = 0, y = 0, z = 0;
    Vec3() {}
    Vec3(float a) : x(a), y(a), z(a) {}
    Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
typedef Vec3 Color3;
struct Matrix44 {   // Imath::M44f layout: row-major x[4][4]
    float x[4][4];
    Matrix44()
    {
        std::memset(x, 0, sizeof x);
        x[0][0] = x[1][1] = x[2][2] = x[3][3] = 1.0f;
    }
    float* operator[](int i) { return x[i]; }
    const float* operator[](int i) const { return x[i]; }
};
typedef const void* TransformationPtr;

struct ClosureColor;
class ShadingContext;
class RendererServices;
class ShadingSystem;
class TextureSystem;
struct PerThreadInfo {};
class ErrorHandler {
public:
    virtual ~ErrorHandler() {}
    virtual void operator()(int /*errcode*/, const std::string& msg) { std::fprintf(stderr, "%s\n", msg.c_str()); }
    enum { EH_ERROR = 3 << 16, EH_WARNING = 2 << 16, EH_INFO = 1 << 16 };
};

// shaderglobals.h:55-146
struct ShaderGlobals {
    Vec3 P, dPdx, dPdy;
    Vec3 dPdz;
    Vec3 I, dIdx, dIdy;
    Vec3 N;
    Vec3 Ng;
    float u, dudx, dudy;
    float v, dvdx, dvdy;
    Vec3 dPdu, dPdv;
    float time;
    float dtime;
    Vec3 dPdtime;
    Vec3 Ps, dPsdx, dPsdy;
    void* renderstate;
    void* tracedata;
    void* objdata;
    ShadingContext* context;
    void* shadingStateUniform;
    int thread_index;
    int shade_index;
    RendererServices* renderer;
    TransformationPtr object2common;
    TransformationPtr shader2common;
    ClosureColor* Ci;
    float surfacearea;
    int raytype;
    int flipHandedness;
    int backfacing;
};

// wide.h:210-444: Block<T,W> is W lanes of T, SoA: a Block<Vec3> is x[W], y[W], z[W]
template<class T, int W> struct alignas(64) Block {
    T data[W];
    T& operator[](int lane) { return data[lane]; }
    const T& operator[](int lane) const { return data[lane]; }
    void assign_all(const T& v)
    {
        for (int i = 0; i < W; ++i)
            data[i] = v;
    }
};
template<int W> struct alignas(64) Block<Vec3, W> {
    float x[W], y[W], z[W];
    struct Ref {
        Block& b;
        int lane;
        Ref& operator=(const Vec3& v)
        {
            b.x[lane] = v.x; b.y[lane] = v.y; b.z[lane] = v.z;
            return *this;
        }
        operator Vec3() const { return Vec3(b.x[lane], b.y[lane], b.z[lane]); }
    };
    Ref operator[](int lane) { return Ref { *this, lane }; }
    Vec3 operator[](int lane) const { return Vec3(x[lane], y[lane], z[lane]); }
    void assign_all(const Vec3& v)
    {
        for (int i = 0; i < W; ++i) {
            x[i] = v.x; y[i] = v.y; z[i] = v.z;
        }
    }
};
// Wide<T,W>: an accessor onto a Block (wide.h:1335-)
template<class T, int W> struct Wide {
    typedef typename std::remove_const<T>::type value_type;
    Wide(const Block<value_type, W>& b) : m_b(&b) {}
    const value_type& operator[](int lane) const { return (*m_b)[lane]; }
    const Block<value_type, W>* m_b;
};

// batched_shaderglobals.h:21-195
struct UniformShaderGlobals {
    void* renderstate = nullptr;
    void* tracedata   = nullptr;
    void* objdata     = nullptr;
    ShadingContext* context    = nullptr;
    RendererServices* renderer = nullptr;
    int raytype = 0;
    int pad0 = 0, pad1 = 0, pad2 = 0, pad3 = 0, pad4 = 0;
};
template<int W> struct alignas(64) VaryingShaderGlobals {
    template<class T> using Blk = OSL::Block<T, W>;
    Blk<Vec3> P, dPdx, dPdy;
    Blk<Vec3> dPdz;
    Blk<Vec3> I, dIdx, dIdy;
    Blk<Vec3> N;
    Blk<Vec3> Ng;
    Blk<float> u, dudx, dudy;
    Blk<float> v, dvdx, dvdy;
    Blk<Vec3> dPdu, dPdv;
    Blk<float> time;
    Blk<float> dtime;
    Blk<Vec3> dPdtime;
    Blk<Vec3> Ps, dPsdx, dPsdy;
    Blk<TransformationPtr> object2common;
    Blk<TransformationPtr> shader2common;
    Blk<ClosureColor*> Ci;
    Blk<float> surfacearea;
    Blk<int> flipHandedness;
    Blk<int> backfacing;
};
template<int W> struct alignas(64) BatchedShaderGlobals {
    BatchedShaderGlobals() {}
    BatchedShaderGlobals(const BatchedShaderGlobals&) = delete;
    UniformShaderGlobals uniform;
    VaryingShaderGlobals<W> varying;
};

enum class SymArena { Unknown, Absolute, Heap, Outputs, UserData };

// oslexec.h:69-105
struct SymLocationDesc {
    ustring name;
    TypeDesc type;
    ptrdiff_t offset = -1;
    ptrdiff_t stride = 0;
    SymArena arena   = SymArena::Unknown;
    bool derivs      = false;
    SymLocationDesc() {}
    SymLocationDesc(string_view name, TypeDesc type, bool derivs = false, SymArena arena = SymArena::Heap,
                    ptrdiff_t offset = -1, ptrdiff_t stride = 0)
        : name(name), type(type), offset(offset), stride(stride), arena(arena), derivs(derivs)
    {
    }
};

// oslclosure.h / genclosure.h: the registration record of one closure parameter
struct ClosureParam {
    TypeDesc type;
    int offset;
    const char* key;
    int field_size;
};
typedef void (*PrepareClosureFunc)(RendererServices*, int id, void* data);
typedef void (*SetupClosureFunc)(RendererServices*, int id, void* data);
#define CLOSURE_FINISH_PARAM(st) { OSL::TypeDesc(), (int)sizeof(st), nullptr, 0 }

struct ParamHints {
    enum { none = 0, interpolated = 1, interactive = 2 };
};
struct ShaderSymbol;

// rendererservices.h:91-602: the callbacks a batch launch needs before it starts.  They are
// evaluated on the host: get_matrix once per batch and named space (coordinate systems are
// uniform over a launch), get_userdata once per point and interpolated parameter.
class RendererServices {
public:
    virtual ~RendererServices() {}
    virtual int supports(string_view /*feature*/) const { return false; }
    virtual bool get_matrix(ShaderGlobals* /*sg*/, Matrix44& /*result*/, TransformationPtr /*xform*/, float /*time*/)
    {
        return false;
    }
    virtual bool get_matrix(ShaderGlobals* sg, Matrix44& result, TransformationPtr xform)
    {
        return get_matrix(sg, result, xform, sg ? sg->time : 0.0f);
    }
    virtual bool get_matrix(ShaderGlobals* /*sg*/, Matrix44& /*result*/, ustringhash /*from*/, float /*time*/)
    {
        return false;
    }
    virtual bool get_matrix(ShaderGlobals* sg, Matrix44& result, ustringhash from)
    {
        return get_matrix(sg, result, from, sg ? sg->time : 0.0f);
    }
    virtual bool get_userdata(bool /*derivatives*/, ustringhash /*name*/, TypeDesc /*type*/, ShaderGlobals* /*sg*/,
                              void* /*val*/)
    {
        return false;
    }
    virtual bool get_attribute(ShaderGlobals*, bool, ustringhash, TypeDesc, ustringhash, void*) { return false; }
    virtual TextureSystem* texturesys() const { return nullptr; }
};

class ShaderGroup {
public:
    ~ShaderGroup()
    {
        if (handle)
            b200_group_destroy(handle);
    }
    struct P {
        std::string name;
        int type = 1;   // 0 int, 1 float-based, 2 string
        std::vector<int> i;
        std::vector<float> f;
        std::vector<std::string> s;
        bool interpolated = false;
    };
    struct Layer {
        std::string oso, shadername, layername;
        std::vector<P> params;
    };
    struct Conn {
        std::string sl, sp, dl, dp;
    };
    std::string name;
    std::vector<Layer> layers;
    std::vector<P> pending;   // Parameter() calls apply to the next Shader()
    std::vector<Conn> conns;
    std::vector<SymLocationDesc> symlocs;
    b200_group* handle = nullptr;
    bool ended         = false;
    std::mutex mu;            // serialises JIT of this group (reference: llvm_instance.cpp:2091)
};
typedef std::shared_ptr<ShaderGroup> ShaderGroupRef;

// One per host thread, never shared (oslexec.h:800-817).  Holds the staging of the
// per-batch renderer callbacks; closures and messages live on the device.
class ShadingContext {
public:
    explicit ShadingContext(ShadingSystem& ss, PerThreadInfo* ti) : m_ss(ss), m_thread(ti) {}
    ShadingSystem& shadingsys() const { return m_ss; }
    PerThreadInfo* thread_info() const { return m_thread; }
    ShaderGroup* group() const { return m_group; }

private:
    friend class ShadingSystem;
    ShadingSystem& m_ss;
    PerThreadInfo* m_thread;
    ShaderGroup* m_group = nullptr;
    std::vector<char> m_userdata;
    std::vector<b200_transform> m_xf;
    std::vector<std::string> m_xfnames;
};

class ShadingSystem {
public:
    ShadingSystem(RendererServices* renderer = nullptr, TextureSystem* texsys = nullptr, ErrorHandler* err = nullptr)
        : m_renderer(renderer), m_err(err)
    {
        (void)texsys;   // texture() images are registered with b200_texture_add or found on texturepath
    }
    ~ShadingSystem() {}

    // ---- attributes (oslexec.h:177-319 names) ---------------------------------------------
    bool attribute(string_view name, TypeDesc type, const void* val)
    {
        if (type == TypeInt)
            return attribute(name, *(const int*)val);
        if (type == TypeFloat)
            return attribute(name, *(const float*)val);
        if (type == TypeString)
            return attribute(name, string_view(*(const char* const*)val));
        return false;
    }
    bool attribute(string_view name, int val)
    {
        m_iattr[std::string(name)] = val;
        return true;
    }
    bool attribute(string_view name, float val)
    {
        m_fattr[std::string(name)] = val;
        return true;
    }
    bool attribute(string_view name, double val) { return attribute(name, (float)val); }
    bool attribute(string_view name, string_view val)
    {
        if (name == "options") {   // "k=v,k=v" list
            std::istringstream in { std::string(val) };
            std::string kv;
            while (std::getline(in, kv, ',')) {
                size_t e = kv.find('=');
                if (e == std::string::npos)
                    continue;
                std::string k = kv.substr(0, e), v = kv.substr(e + 1);
                char* end = nullptr;
                long iv   = std::strtol(v.c_str(), &end, 10);
                if (end && *end == 0)
                    attribute(k, (int)iv);
                else
                    attribute(k, string_view(v));
            }
            return true;
        }
        m_sattr[std::string(name)] = std::string(val);
        return true;
    }
    bool attribute(ShaderGroup* group, string_view name, TypeDesc type, const void* val)
    {
        // per-group attributes ("renderer_outputs", "entry_layers", "groupname"): outputs are
        // placed with add_symlocs on this back end; the names are accepted for compatibility
        (void)group; (void)name; (void)type; (void)val;
        return true;
    }
    bool getattribute(string_view name, int& val) const
    {
        auto it = m_iattr.find(std::string(name));
        if (it == m_iattr.end())
            return false;
        val = it->second;
        return true;
    }
    bool getattribute(string_view name, std::string& val) const
    {
        auto it = m_sattr.find(std::string(name));
        if (it == m_sattr.end())
            return false;
        val = it->second;
        return true;
    }
    // getattribute(group, "b200_cuda_source" | "num_renderer_outputs" ...) (oslexec.h:520-596)
    bool getattribute(ShaderGroup* group, string_view name, std::string& val)
    {
        if (!group || !optimize_group(group))
            return false;
        if (name == "b200_cuda_source") {
            val = b200_group_cuda_source(group->handle);
            return true;
        }
        if (name == "groupname") {
            val = group->name;
            return true;
        }
        return false;
    }
    std::string geterror(bool clear = true)
    {
        std::string e = m_errmsg;
        if (clear)
            m_errmsg.clear();
        return e;
    }
    bool has_error() const { return !m_errmsg.empty(); }
    int b200_device() const
    {
        auto it = m_iattr.find("b200_device");
        return it == m_iattr.end() ? 0 : it->second;
    }

    // ---- shader sources -----------------------------------------------------------------
    bool LoadMemoryCompiledShader(string_view shadername, string_view buffer)
    {
        m_mem[std::string(shadername)] = std::string(buffer);
        return true;
    }

    // ---- group construction (oslexec.h:634-760) -------------------------------------------
    ShaderGroupRef ShaderGroupBegin(string_view groupname = string_view())
    {
        auto g  = std::make_shared<ShaderGroup>();
        g->name = std::string(groupname);
        return g;
    }
    // the serialized form: "param float Kd 0.5; shader matte layer1; connect a.out b.in;"
    // with ',' accepted for ';' (shadingsys.cpp:3232-3300)
    ShaderGroupRef ShaderGroupBegin(string_view groupname, string_view usage, string_view groupspec)
    {
        ShaderGroupRef g = ShaderGroupBegin(groupname);
        std::string spec(groupspec);
        for (char& c : spec)
            if (c == ',')
                c = ';';
        std::istringstream in(spec);
        std::string stmt;
        while (std::getline(in, stmt, ';')) {
            std::istringstream ts(stmt);
            std::vector<std::string> t;
            std::string w;
            while (ts >> w)
                t.push_back(w);
            if (t.empty())
                continue;
            size_t k = 0;
            if (t[0] == "param")
                k = 1;
            if (t[k] == "shader") {
                if (t.size() < k + 3 || !Shader(*g, usage, t[k + 1], t[k + 2]))
                    return nullptr;
            } else if (t[k] == "connect") {
                if (t.size() < k + 3)
                    return nullptr;
                auto split = [](const std::string& s, std::string& l, std::string& p) {
                    size_t d = s.find('.');
                    l = s.substr(0, d);
                    p = d == std::string::npos ? "" : s.substr(d + 1);
                };
                std::string sl, sp, dl, dp;
                split(t[k + 1], sl, sp);
                split(t[k + 2], dl, dp);
                ConnectShaders(*g, sl, sp, dl, dp);
            } else {   // <type> <name> <values...>
                if (t.size() < k + 3) {
                    error("ShaderGroupBegin: cannot parse \"" + stmt + "\"");
                    return nullptr;
                }
                const std::string &type = t[k], &pname = t[k + 1];
                if (type == "int") {
                    std::vector<int> v;
                    for (size_t j = k + 2; j < t.size(); ++j)
                        v.push_back(std::atoi(t[j].c_str()));
                    Parameter(*g, pname, TypeDesc(TypeDesc::INT, v.size() > 1 ? (int)v.size() : 0), v.data());
                } else if (type == "string") {
                    std::string s = t[k + 2];
                    if (s.size() >= 2 && s.front() == '"')
                        s = s.substr(1, s.size() - 2);
                    const char* cs = s.c_str();
                    Parameter(*g, pname, TypeString, &cs);
                } else {
                    std::vector<float> v;
                    for (size_t j = k + 2; j < t.size(); ++j)
                        v.push_back(std::strtof(t[j].c_str(), nullptr));
                    TypeDesc td = type == "float" ? TypeDesc(TypeDesc::FLOAT, v.size() > 1 ? (int)v.size() : 0)
                                                  : (type == "matrix" ? TypeMatrix : TypeColor);
                    Parameter(*g, pname, td, v.data());
                }
            }
        }
        return g;
    }
    bool Parameter(ShaderGroup& g, string_view name, TypeDesc t, const void* val, int hints = ParamHints::none)
    {
        ShaderGroup::P p;
        p.name         = std::string(name);
        p.interpolated = (hints & ParamHints::interpolated) != 0;
        size_t n       = t.numelements() * t.aggregate;
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
            return error("Parameter: unknown type for '" + p.name + "'");
        g.pending.push_back(std::move(p));
        return true;
    }
    bool Parameter(ShaderGroup& g, string_view name, TypeDesc t, const void* val, bool lockgeom)
    {
        return Parameter(g, name, t, val, lockgeom ? ParamHints::none : ParamHints::interpolated);
    }
    bool Shader(ShaderGroup& g, string_view /*shaderusage*/, string_view shadername, string_view layername)
    {
        ShaderGroup::Layer l;
        std::string sn(shadername);
        auto it = m_mem.find(sn);
        if (it != m_mem.end())
            l.oso = it->second;
        else {
            std::string dirs = m_sattr.count("searchpath:shader") ? m_sattr["searchpath:shader"] : ".";
            std::istringstream ds(dirs);
            std::string d;
            while (l.oso.empty() && std::getline(ds, d, ':')) {
                std::ifstream f(d + "/" + sn + ".oso");
                if (f) {
                    std::stringstream ss;
                    ss << f.rdbuf();
                    l.oso = ss.str();
                }
            }
            if (l.oso.empty())
                return error("Could not find shader \"" + sn + "\"");
        }
        l.shadername = sn;
        l.layername  = layername.empty() ? sn : std::string(layername);
        l.params.swap(g.pending);
        g.layers.push_back(std::move(l));
        return true;
    }
    bool ConnectShaders(ShaderGroup& g, string_view srclayer, string_view srcparam, string_view dstlayer,
                        string_view dstparam)
    {
        g.conns.push_back({ std::string(srclayer), std::string(srcparam), std::string(dstlayer), std::string(dstparam) });
        return true;
    }
    bool ShaderGroupEnd(ShaderGroup& g)
    {
        g.ended = true;
        return true;
    }
    // oslexec.h:749: change an instance value after the group was declared.  Instance values
    // are compile-time constants of the generated kernel, so the group is re-JITed on next use.
    bool ReParameter(ShaderGroup& g, string_view layername, string_view paramname, TypeDesc t, const void* val)
    {
        for (auto& l : g.layers) {
            if (l.layername != layername)
                continue;
            ShaderGroup::P* slot = nullptr;
            for (auto& p : l.params)
                if (p.name == paramname)
                    slot = &p;
            if (!slot) {
                l.params.emplace_back();
                slot       = &l.params.back();
                slot->name = std::string(paramname);
            }
            ShaderGroup tmp;
            if (!Parameter(tmp, paramname, t, val))
                return false;
            bool interp = slot->interpolated;
            *slot       = tmp.pending[0];
            slot->interpolated = interp;
            std::lock_guard<std::mutex> lk(g.mu);
            if (g.handle) {
                b200_group_destroy(g.handle);
                g.handle = nullptr;
            }
            return true;
        }
        return error("ReParameter: no layer \"" + std::string(layername) + "\"");
    }

    // ---- output / userdata placement (oslexec.h:1075) -------------------------------------
    void add_symlocs(ShaderGroup* g, const SymLocationDesc* locs, size_t n)
    {
        for (size_t i = 0; i < n; ++i)
            g->symlocs.push_back(locs[i]);
    }
    template<class C> void add_symlocs(ShaderGroup* g, const C& locs) { add_symlocs(g, locs.data(), locs.size()); }
    const SymLocationDesc* find_symloc(const ShaderGroup* g, ustring name, SymArena arena) const
    {
        for (auto& s : g->symlocs)
            if (s.name == name && s.arena == arena)
                return &s;
        return nullptr;
    }

    // ---- closures (oslexec.h:1041-1048) ---------------------------------------------------
    // The device integrator knows testrender's closure set by id (src/testrender/shading.h:25-60);
    // a registration is recorded and must agree with it.
    void register_closure(string_view name, int id, const ClosureParam* params, PrepareClosureFunc, SetupClosureFunc)
    {
        Closure c;
        c.id = id;
        for (const ClosureParam* p = params; p && p->type.basetype != TypeDesc::UNKNOWN; ++p)
            c.params.push_back(*p);
        ClosureParam fin = { TypeDesc(), 0, nullptr, 0 };
        c.params.push_back(fin);
        m_closures[std::string(name)] = c;
    }
    bool query_closure(const char** name, int* id, const ClosureParam** params)
    {
        for (auto& kv : m_closures)
            if ((name && *name && kv.first == *name) || ((!name || !*name) && id && kv.second.id == *id)) {
                if (name)
                    *name = kv.first.c_str();
                if (id)
                    *id = kv.second.id;
                if (params)
                    *params = kv.second.params.data();
                return true;
            }
        return false;
    }
    int raytype_bit(ustring name)
    {
        static const char* names[] = { "camera", "shadow", "reflection", "refraction", "diffuse", "glossy",
                                       "subsurface", "displacement" };
        for (int i = 0; i < 8; ++i)
            if (name == ustring(names[i]))
                return 1 << i;
        return 0;
    }

    // ---- contexts (oslexec.h:800-824) -----------------------------------------------------
    PerThreadInfo* create_thread_info() { return new PerThreadInfo; }
    void destroy_thread_info(PerThreadInfo* ti) { delete ti; }
    ShadingContext* get_context(PerThreadInfo* ti = nullptr, void* /*texture_threadinfo*/ = nullptr)
    {
        return new ShadingContext(*this, ti);
    }
    void release_context(ShadingContext* ctx) { delete ctx; }

    // ---- JIT (oslexec.h:1093; BackendLLVM::run -> generator + NVRTC, no GPU needed) --------
    bool optimize_group(ShaderGroup* g, ShadingContext* = nullptr, bool /*do_jit*/ = true)
    {
        std::lock_guard<std::mutex> lk(g->mu);
        if (g->handle)
            return true;
        std::vector<b200_layer> L(g->layers.size());
        std::vector<std::vector<b200_param>> P(g->layers.size());
        std::vector<std::vector<const char*>> S;
        S.reserve(64);
        for (size_t i = 0; i < g->layers.size(); ++i) {
            auto& l = g->layers[i];
            for (auto& p : l.params) {
                if (p.interpolated)
                    continue;   // the value comes from userdata; the .oso default stands otherwise
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
        std::vector<b200_userdata> U;
        for (auto& s : g->symlocs) {
            if (s.arena == SymArena::Outputs)
                O.push_back({ s.name.c_str(), (long long)s.offset, (long long)s.stride, s.derivs ? 1 : 0 });
            else if (s.arena == SymArena::UserData)
                U.push_back({ s.name.c_str(), (int)s.type.aggregate, s.type.basetype == TypeDesc::INT, (long long)s.offset,
                              (long long)s.stride, s.derivs ? 1 : 0, -1, 0 });
        }
        // FMA: the reference's scalar default is strict, batched turns it on (testshade.cpp:294-298)
        int fma = 0;
        getattribute("llvm_jit_fma", fma);
        std::string opts = std::string("fma=") + (fma ? "1" : "0");
        int v;
        if (getattribute("b200_block", v))
            opts += ",block=" + std::to_string(v);
        if (getattribute("b200_journal", v) && v)
            opts += ",journal=" + std::to_string(v);
        std::string sv;
        if (getattribute("searchpath:texture", sv))
            opts += ",texturepath=" + sv;
        if (getattribute("colorspace", sv))
            opts += ",colorspace=" + sv;
        if (U.empty() && m_renderer)
            opts += ",userdata=record";   // fed from RendererServices::get_userdata at execute
        b200_group_desc d { g->name.c_str(), (int)L.size(), L.data(), (int)C.size(), C.data(),
                            (int)O.size(), O.data(), opts.c_str(), (int)U.size(), U.data() };
        if (b200_group_compile(&d, &g->handle) != B200_OK)
            return error(b200_last_error());
        return true;
    }

    // ---- scalar execution: one point per call (oslexec.h:833-918) ----------------------------
    // The reference's per-point entry.  It works (a one-point launch through the host path);
    // renderers that care about throughput call batched<W>() with whole tiles.
    bool execute(ShadingContext& ctx, ShaderGroup& group, int thread_index, int shadeindex, ShaderGlobals& sg,
                 void* userdata_base_ptr, void* output_base_ptr, bool run = true)
    {
        sg.thread_index = thread_index;
        sg.shade_index  = shadeindex;
        if (!run)
            return optimize_group(&group);
        b200_globals bg;
        std::memset(&bg, 0, sizeof bg);
        const float* f = &sg.P.x;
        static const int triple[B200_SG_NFIELDS] = { 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 1, 1, 0, 0, 1, 1, 1, 1, 0, 0, 0, 0 };
        for (int k = 0; k < B200_SG_surfacearea; ++k)   // the float fields up to dPsdy are contiguous
            for (int c = 0; c < (triple[k] ? 3 : 1); ++c)
                bg.uniform[k][c] = *f++;
        bg.uniform[B200_SG_surfacearea][0] = sg.surfacearea;
        std::memcpy(&bg.uniform[B200_SG_raytype][0], &sg.raytype, 4);
        std::memcpy(&bg.uniform[B200_SG_flipHandedness][0], &sg.flipHandedness, 4);
        std::memcpy(&bg.uniform[B200_SG_backfacing][0], &sg.backfacing, 4);
        sg.renderer = m_renderer;
        sg.context  = &ctx;
        return run_batch(ctx, group, 1, shadeindex, bg, &sg, 1, userdata_base_ptr, output_base_ptr);
    }
    bool execute(ShadingContext* ctx, ShaderGroup& group, ShaderGlobals& sg, bool run = true)
    {
        ShadingContext* c = ctx ? ctx : get_context();
        bool ok           = execute(*c, group, 0, 0, sg, nullptr, nullptr, run);
        if (!ctx)
            release_context(c);
        return ok;
    }
    bool execute_init(ShadingContext& ctx, ShaderGroup& group, int thread_index, int shadeindex, ShaderGlobals& sg,
                      void* userdata_base_ptr, void* output_base_ptr, bool run = true)
    {
        // the whole group runs in one launch; init only binds and JITs (context.cpp:92-163)
        (void)thread_index; (void)shadeindex; (void)sg; (void)userdata_base_ptr; (void)output_base_ptr; (void)run;
        ctx.m_group = &group;
        return optimize_group(&group);
    }
    bool execute_layer(ShadingContext& ctx, int thread_index, int shadeindex, ShaderGlobals& sg, void* userdata_base_ptr,
                       void* output_base_ptr, int layernumber)
    {
        // entry layers other than the last are not exposed by this back end
        if (!ctx.m_group || layernumber != (int)ctx.m_group->layers.size() - 1)
            return error("execute_layer: only the group's last layer can be an entry point on this back end");
        return execute(ctx, *ctx.m_group, thread_index, shadeindex, sg, userdata_base_ptr, output_base_ptr, true);
    }
    bool execute_cleanup(ShadingContext& ctx)
    {
        ctx.m_group = nullptr;
        return true;
    }

    // ---- symbols (oslexec.h:956-971): renderer outputs placed with add_symlocs ----------------
    const void* get_symbol(ShadingContext& ctx, ustring layername, ustring symbolname, TypeDesc& type) const
    {
        (void)ctx; (void)layername; (void)symbolname; (void)type;
        return nullptr;   // outputs are delivered through the Outputs arena, not the context heap
    }

    // ---- batched execution (oslexec.h:982-1033) ----------------------------------------------
    bool configure_batch_execution_at(int width) { return width == 16 || width == 8 || width == 4; }

    template<int WidthT> class BatchedExecutor {
        ShadingSystem& m_shading_system;

    public:
        explicit BatchedExecutor(ShadingSystem& ss) : m_shading_system(ss) {}
        BatchedExecutor(const BatchedExecutor&) = default;
        void jit_group(ShaderGroup* group, ShadingContext* ctx) { m_shading_system.optimize_group(group, ctx); }
        void jit_all_groups(int /*nthreads*/ = 0) {}
        // The reference's call: WidthT lanes, AoSoA blocks, per-lane shade indices.
        bool execute(ShadingContext& ctx, ShaderGroup& group, int batch_size, Wide<const int, WidthT> wide_shadeindex,
                     BatchedShaderGlobals<WidthT>& bsg, void* userdata_base_ptr, void* output_base_ptr, bool run = true)
        {
            if (!run)
                return m_shading_system.optimize_group(&group);
            if (batch_size <= 0 || batch_size > WidthT)
                return m_shading_system.error("BatchedExecutor::execute: batch_size out of range");
            // a Block<Vec3,W> is x[W] y[W] z[W]: exactly the SoA planes of b200_globals with
            // plane_stride = W, so the blocks are passed as they are
            b200_globals bg;
            std::memset(&bg, 0, sizeof bg);
            VaryingShaderGlobals<WidthT>& v = bsg.varying;
            bg.plane_stride = WidthT;
            const float* planes[B200_SG_NFIELDS]
                = { v.P.x, v.dPdx.x, v.dPdy.x, v.dPdz.x, v.I.x, v.dIdx.x, v.dIdy.x, v.N.x, v.Ng.x, v.u.data, v.dudx.data,
                    v.dudy.data, v.v.data, v.dvdx.data, v.dvdy.data, v.dPdu.x, v.dPdv.x, v.time.data, v.dtime.data,
                    v.dPdtime.x, v.Ps.x, v.dPsdx.x, v.dPsdy.x, v.surfacearea.data, nullptr,
                    (const float*)v.flipHandedness.data, (const float*)v.backfacing.data };
            for (int k = 0; k < B200_SG_NFIELDS; ++k)
                bg.varying[k] = planes[k];
            std::memcpy(&bg.uniform[B200_SG_raytype][0], &bsg.uniform.raytype, 4);
            bsg.uniform.renderer = m_shading_system.m_renderer;
            bsg.uniform.context  = &ctx;
            // the renderer callbacks see one ShaderGlobals per lane
            std::vector<ShaderGlobals> lanes;
            if (m_shading_system.m_renderer) {
                lanes.resize(batch_size);
                for (int l = 0; l < batch_size; ++l)
                    m_shading_system.lane_globals(bsg, l, lanes[l]);
            }
            // consecutive shade indices (a row of pixels: testshade.cpp:1855-1881) run as one
            // launch; anything else lane by lane
            bool consecutive = true;
            for (int l = 1; l < batch_size; ++l)
                consecutive &= wide_shadeindex[l] == wide_shadeindex[0] + l;
            if (consecutive)
                return m_shading_system.run_batch(ctx, group, batch_size, wide_shadeindex[0], bg,
                                                  lanes.empty() ? nullptr : lanes.data(), batch_size, userdata_base_ptr,
                                                  output_base_ptr);
            for (int l = 0; l < batch_size; ++l) {
                b200_globals one = bg;
                for (int k = 0; k < B200_SG_NFIELDS; ++k)
                    if (one.varying[k])
                        one.varying[k] += l;
                if (!m_shading_system.run_batch(ctx, group, 1, wide_shadeindex[l], one, lanes.empty() ? nullptr : &lanes[l],
                                                1, userdata_base_ptr, output_base_ptr))
                    return false;
            }
            return true;
        }
        // Whole-tile call (this back end's reason to exist): npoints shading points with the
        // shade indices first_shadeindex.., HOST SoA planes in `globals`.
        bool execute(ShadingContext& ctx, ShaderGroup& group, long long npoints, long long first_shadeindex,
                     const b200_globals& globals, void* userdata_base_ptr, void* output_base_ptr)
        {
            return m_shading_system.run_batch(ctx, group, npoints, first_shadeindex, globals, nullptr, 0,
                                              userdata_base_ptr, output_base_ptr);
        }
        // ... and with DEVICE pointers, asynchronous on `stream` (no renderer callbacks: transforms
        // and userdata are the caller's, in `globals` and `userdata_base_ptr`)
        bool execute_device(ShaderGroup& group, long long npoints, const int* wide_shadeindex, const b200_globals& globals,
                            const void* userdata_base_ptr, void* output_base_ptr, void* stream = nullptr)
        {
            if (!m_shading_system.optimize_group(&group))
                return false;
            if (b200_group_execute(group.handle, m_shading_system.b200_device(), stream, npoints, &globals, wide_shadeindex,
                                   userdata_base_ptr, output_base_ptr) != B200_OK)
                return m_shading_system.error(b200_last_error());
            return true;
        }
        bool execute_init(ShadingContext& ctx, ShaderGroup& group, int, Wide<const int, WidthT>, BatchedShaderGlobals<WidthT>&,
                          void*, void*, bool = true)
        {
            ctx.m_group = &group;
            return m_shading_system.optimize_group(&group);
        }
    };
    template<int WidthT> BatchedExecutor<WidthT> batched() { return BatchedExecutor<WidthT>(*this); }

    bool error(const std::string& m)
    {
        m_errmsg = m;
        if (m_err)
            (*m_err)(ErrorHandler::EH_ERROR, m);
        return false;
    }
    RendererServices* renderer() const { return m_renderer; }

private:
    template<int W> friend class BatchedExecutor;
    struct Closure {
        int id = 0;
        std::vector<ClosureParam> params;
    };
    template<int W> void lane_globals(const BatchedShaderGlobals<W>& b, int l, ShaderGlobals& sg)
    {
        std::memset((void*)&sg, 0, sizeof sg);
        const VaryingShaderGlobals<W>& v = b.varying;
        sg.P = v.P[l]; sg.dPdx = v.dPdx[l]; sg.dPdy = v.dPdy[l]; sg.dPdz = v.dPdz[l];
        sg.I = v.I[l]; sg.dIdx = v.dIdx[l]; sg.dIdy = v.dIdy[l]; sg.N = v.N[l]; sg.Ng = v.Ng[l];
        sg.u = v.u[l]; sg.dudx = v.dudx[l]; sg.dudy = v.dudy[l]; sg.v = v.v[l]; sg.dvdx = v.dvdx[l]; sg.dvdy = v.dvdy[l];
        sg.dPdu = v.dPdu[l]; sg.dPdv = v.dPdv[l]; sg.time = v.time[l]; sg.dtime = v.dtime[l]; sg.dPdtime = v.dPdtime[l];
        sg.Ps = v.Ps[l]; sg.dPsdx = v.dPsdx[l]; sg.dPsdy = v.dPsdy[l];
        sg.object2common = v.object2common[l]; sg.shader2common = v.shader2common[l];
        sg.surfacearea = v.surfacearea[l]; sg.flipHandedness = v.flipHandedness[l]; sg.backfacing = v.backfacing[l];
        sg.raytype = b.uniform.raytype; sg.renderstate = b.uniform.renderstate; sg.tracedata = b.uniform.tracedata;
        sg.objdata = b.uniform.objdata; sg.renderer = m_renderer; sg.context = b.uniform.context;
    }
    // Everything a launch needs from the renderer, then the launch: named coordinate systems
    // through RendererServices::get_matrix (uniform per batch), interpolated parameters through
    // get_userdata (one record per point, layout chosen by the library: option userdata=record).
    bool run_batch(ShadingContext& ctx, ShaderGroup& group, long long npoints, long long first, const b200_globals& globals,
                   ShaderGlobals* lanes, int nlanes, void* userdata_base_ptr, void* output_base_ptr)
    {
        if (!optimize_group(&group))
            return false;
        b200_globals bg = globals;
        ShaderGlobals* sg0 = lanes;
        if (m_renderer && bg.ntransforms == 0) {
            const int ns = b200_group_num_spaces(group.handle);
            ctx.m_xf.clear();
            ctx.m_xfnames.clear();
            ctx.m_xfnames.reserve(ns);
            for (int k = 0; k < ns; ++k) {
                Matrix44 M;
                std::string name = b200_group_space_name(group.handle, k);
                bool ok = false;
                if (name == "shader" && sg0)
                    ok = m_renderer->get_matrix(sg0, M, sg0->shader2common);
                else if (name == "object" && sg0)
                    ok = m_renderer->get_matrix(sg0, M, sg0->object2common);
                else
                    ok = m_renderer->get_matrix(sg0, M, ustring(name));
                if (!ok)
                    continue;
                ctx.m_xfnames.push_back(name);
                b200_transform t;
                t.name = ctx.m_xfnames.back().c_str();
                std::memcpy(t.m, M.x, sizeof t.m);
                ctx.m_xf.push_back(t);
            }
            bg.ntransforms = (int)ctx.m_xf.size();
            bg.transforms  = ctx.m_xf.data();
        }
        const void* ud   = userdata_base_ptr;
        long long udsize = 0;
        long long rec    = 0;
        const int nf     = b200_group_userdata_fields(group.handle, &rec);
        if (rec > 0 && nf > 0) {
            // library-defined records, filled from get_userdata for the points of this batch
            if (!lanes || nlanes < npoints)
                return error("this group has interpolated parameters: the whole-tile execute needs the caller's "
                             "userdata arena (SymArena::UserData symlocs) instead of RendererServices::get_userdata");
            ctx.m_userdata.assign((size_t)(rec * npoints), 0);
            for (int f = 0; f < nf; ++f) {
                b200_userdata u;
                b200_group_userdata_field(group.handle, f, &u);
                TypeDesc td = u.is_int ? TypeInt : (u.ncomp == 3 ? TypeColor : TypeFloat);
                for (long long i = 0; i < npoints; ++i) {
                    char* base = ctx.m_userdata.data() + rec * i;
                    lanes[i].shade_index = (int)(first + i);
                    int ok = m_renderer && m_renderer->get_userdata(u.derivs != 0, ustring(u.name), td, &lanes[i], base + (u.offset % rec));
                    std::memcpy(base + (u.valid_offset % rec), &ok, 4);
                }
            }
            ud     = ctx.m_userdata.data();
            udsize = (long long)ctx.m_userdata.size();
        } else if (ud && nf > 0) {
            // caller-described arrays (SymArena::UserData): the arena must cover every field
            for (int f = 0; f < nf; ++f) {
                b200_userdata u;
                b200_group_userdata_field(group.handle, f, &u);
                long long end = u.offset + u.stride * (first + npoints - 1) + 4LL * u.ncomp * (u.derivs ? 3 : 1);
                udsize        = end > udsize ? end : udsize;
            }
        }
        if (b200_group_execute_host_at(group.handle, b200_device(), npoints, &bg, first, ud, udsize, output_base_ptr)
            != B200_OK)
            return error(b200_last_error());
        return true;
    }

    RendererServices* m_renderer;
    ErrorHandler* m_err;
    std::map<std::string, std::string> m_sattr, m_mem;
    std::map<std::string, int> m_iattr;
    std::map<std::string, float> m_fattr;
    std::map<std::string, Closure> m_closures;
    std::string m_errmsg;
};

}  // namespace OSL
