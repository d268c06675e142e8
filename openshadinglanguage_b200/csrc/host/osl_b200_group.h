// osl_b200_group.h — host-side group IR for the B200 back end (product code).
//
// Plays the role of the reference's ShaderMaster / ShaderInstance / ShaderGroup
// (src/liboslexec/oslexec_pvt.h:409,1231,1777) for exactly what the back end
// consumes: parsed .oso symbols and ops, instance parameter values,
// connections, renderer outputs (SymLocationDesc, include/OSL/oslexec.h:69-105)
// and the three optimizer facts the code generator needs — unused(),
// run_lazily(), has_derivs().
#pragma once
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace oslb200 {

enum class Base { Int, Float, String, Color, Point, Vector, Normal, Matrix, Closure, Void };

struct TypeSpec {
    Base base   = Base::Float;
    int arraylen = 0;  // 0: not an array
    bool is_triple() const
    {
        return base == Base::Color || base == Base::Point || base == Base::Vector
               || base == Base::Normal;
    }
    bool is_float_based() const { return base == Base::Float || base == Base::Matrix || is_triple(); }
    int ncomp() const { return is_triple() ? 3 : (base == Base::Matrix ? 16 : 1); }
};

enum class SymType { Param, OutputParam, Local, Temp, Global, Const };

struct OutputLoc {
    bool placed      = false;
    long long offset = 0, stride = 0;
    bool derivs      = false;
};

struct Symbol {
    std::string name;
    SymType symtype = SymType::Local;
    TypeSpec type;
    std::vector<float> fvals;
    std::vector<int> ivals;
    std::vector<std::string> svals;
    bool initexpr       = false;
    bool unsized        = false;  // declared "type[]": length from the default list, the instance value or a connection
    bool interpolated   = false;  // lockgeom=0: the renderer may supply the value per point (userdata)
    bool has_derivs     = false;
    bool written        = false;
    bool connected_down = false;
    int conn_layer      = -1;  // upstream layer feeding this param, or -1
    int conn_sym        = -1;  // symbol index in that layer
    OutputLoc out;
    bool is_const() const { return symtype == SymType::Const; }
    bool is_param() const { return symtype == SymType::Param || symtype == SymType::OutputParam; }
    // would the runtime optimizer have folded this to a constant?
    bool const_value() const
    {
        return is_const() || (symtype == SymType::Param && conn_layer < 0 && !initexpr && !written && !interpolated);
    }
};

struct Opcode {
    std::string name;
    std::vector<int> args;  // symbol indices
    std::vector<int> jumps;
    std::string rw;
    std::vector<int> derivs;  // arg indices whose derivatives the op takes
    bool reads(int i) const { return rw[i] == 'r' || rw[i] == 'W'; }
    bool writes(int i) const { return rw[i] == 'w' || rw[i] == 'W'; }
};

struct Master {
    std::string shadertype, shadername;
    std::vector<Symbol> syms;
    std::map<std::string, int> byname;
    std::vector<Opcode> ops;
    std::map<std::string, std::pair<int, int>> methods;  // code sections
    int find(const std::string& n) const
    {
        auto it = byname.find(n);
        return it == byname.end() ? -1 : it->second;
    }
};

// Parse OSO 1.00 text (grammar: src/liboslexec/osogram.y:88-318).  Throws
// std::runtime_error with a line-numbered message on malformed input.
Master parse_oso(const std::string& text);

struct ParamValue {
    std::string name;
    std::vector<float> fvals;
    std::vector<int> ivals;
    std::vector<std::string> svals;
};

struct Layer {
    Master m;
    std::string layername;
    bool lazy   = false;
    bool unused = false;
};

struct Connection {
    int srclayer, srcsym, dstlayer, dstsym;
};

// Renderer outputs whose records interleave (same stride, offsets inside one
// stride) form one cluster: for consecutive shade indices the cluster is one
// contiguous byte range, which is what the kernel stages in shared memory and
// the host path moves with one copy.
struct OutCluster {
    long long stride = 0, lo = 0, hi = 0;  // one record covers bytes [lo, hi)
    std::vector<int> outs;                 // indices into Group::outputs
    bool dense = false;                    // fields tile [lo, lo+stride) exactly
};

// One printf() site of the group: the format string and the layout of the argument words
// its device record carries (the journal is decoded and formatted on the host).
struct JournalArg {
    Base base    = Base::Float;
    int ncomp    = 1;
    int arraylen = 0;
};
struct JournalFormat {
    std::string fmt;
    std::vector<JournalArg> args;
    int kind = 0;            // 0 printf, 1 error(), 2 warning(), 3 error of the shading system itself
    std::string shadername;  // for the "Shader error [name]: " prefix
};

// one b200_userdata entry (include/osl_b200.h)
struct UserData {
    std::string name;
    int ncomp = 1;
    bool is_int = false, derivs = false;
    long long offset = 0, stride = 0, valid_offset = -1, valid_stride = 0;
};

// a uniform renderer attribute (b200_attribute): what RendererServices::get_attribute answers for every point
struct Attribute {
    std::string name;
    int type = 1;   // 0 int, 1 float, 2 string
    std::vector<int> ivals;
    std::vector<float> fvals;
    std::vector<std::string> svals;
};

struct Group {
    std::string name;
    std::vector<Attribute> attributes;         // uniform renderer attributes getattribute() can return
    std::vector<UserData> userdata;            // what the renderer supplies for interpolated params
    std::vector<OutCluster> clusters;
    bool stage_ok = false;  // every cluster dense and small enough to stage
    long long stage_record_bytes = 0;  // largest output record (cluster stride)
    int block     = 256;    // CTA size the kernel is generated for
    std::vector<Layer> layers;
    std::vector<Connection> connections;
    std::vector<std::pair<int, int>> outputs;  // (layer, sym) with out.placed
    std::vector<std::string> strings;          // interned string table (id = index)
    std::vector<std::string> warnings;
    std::set<int> globals_read;                // b200_sg_field ids the kernel loads
    bool fma = true;                           // allow FMA contraction in generated code
    bool uses_glossy_lobes = false;            // set by codegen: phong / ward / microfacet closures
    bool uses_thinlayer    = false;            // set by codegen: the thinlayer closure (spi::ThinLayerLobe)
    bool uses_sheen_ltc    = false;            // set by codegen: sheen_bsdf with a "mode" keyword
    bool uses_media        = false;            // set by codegen: medium_vdf / anisotropic_vdf closures
    bool uses_mx_lobes     = false;            // set by codegen: conductor / dielectric / generalized schlick ...
    bool uses_colorsystem  = false;            // set by codegen: luminance / blackbody / transformc ...
    std::string colorspace = "Rec709";         // ShadingSystem attribute "colorspace"
    std::vector<JournalFormat> jformats;       // printf sites (id = index), grid kernels only
    bool journal_enabled = false;              // group option journal=WORDS (> 0): record printf output
    std::vector<std::string> spaces;           // named coordinate systems referenced (launch-block slots)
    std::string commonspace_synonym = "world"; // ShadingSystem attribute "commonspace"
    std::vector<std::string> textures;         // constant texture() file names (module table slots)
    int texture_base = 0;                      // first slot of this group in a multi-group module
    // Static bounds of one execution's closure arena, set by codegen from the group's closure ops
    // (a layer runs at most once per execution).  closure_in_loop: a closure op sits inside a loop,
    // the bounds do not hold and the integrator keeps the reference's 1 KB pool / 8 lobes.
    int pool_words_bound = 1;                  // word 0 is reserved
    int lobe_bound       = 0;                  // BSDF components that can reach the CompositeBSDF
    int closure_adds     = 0;                  // add / layer nodes: bounds the tree-walk stack
    bool closure_in_loop = false;
    std::set<std::string> closure_names;       // closures the group can emit: its closure-type signature
    std::string texturepath;                   // ':'-separated directories searched for texture files

    int layer_index(const std::string& n) const;
    void add_layer(const std::string& oso_text, const std::string& layername,
                   const std::vector<ParamValue>& params);
    void connect(const std::string& sl, const std::string& sp, const std::string& dl,
                 const std::string& dp);
    void add_output(const std::string& name, long long offset, long long stride, bool derivs);
    void finalize();  // unused/lazy/derivs analysis
    int intern(const std::string& s);
};

// Colour system of a working space as floats (osl_b200_color.cpp; false: unknown name)
// and as the CUDA constant array `osl_cs_` that generated modules embed.
bool colorsystem_table(const std::string& colorspace, std::vector<float>& out);
std::string colorsystem_cuda_definition(const std::string& colorspace);

// Emit CUDA C++ for the whole group (kernel name: osl_b200_group_kernel).
std::string generate_cuda(Group& g);

// Emit one CUDA module holding every material group of a scene (namespaces
// mat0, mat1, ...), the shader dispatch switch and the wavefront integrator
// kernels of csrc/device/osl_b200_render.cuh.
// what the render module was specialised to (the host sizes shared memory from it)
struct RenderModuleInfo {
    int pool_words    = 256;    // OSLD_POOL_WORDS: per-thread closure arena
    int max_lobes     = 8;      // OSLD_MAX_LOBES
    int closure_stack = 16;     // OSLD_CLOSURE_STACK
    bool pool_in_smem = false;  // OSLD_POOL_SMEM: arena staged in shared memory
    bool uses_mx_lobes = false; // OSLD_MX_LOBES: the module reads the libbsdl energy tables
    bool uses_luts     = false; // the module reads the LUT block (energy tables and / or sheen LTC coefficients)
    bool uses_media    = false; // OSLD_HAS_MEDIA: every path slot carries a medium stack
};
std::string generate_cuda_render(std::vector<Group*>& groups, bool has_background = false,
                                 RenderModuleInfo* info = nullptr);

}  // namespace oslb200
