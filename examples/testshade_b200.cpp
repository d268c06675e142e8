// testshade_b200 — the reference's testshade grid harness (src/testshade/testshade.cpp) against
// include/OSL/oslexec.h, i.e. written with the reference's own API calls:
//
//   testshade_b200 [-g W H] [--center] [--searchpath DIR] [--fma 0|1] [--iters N] [--jitonly]
//                  [--batched16] [--param name value]... [--layer NAME] shader
//                  [--connect L1 P1 L2 P2]... [-o OUTPUT file.f32]... [-od3 OUTPUT file.f32]...
//
// The group is built with the call sequence of testshade.cpp:2028-2080 (ShaderGroupBegin /
// Parameter / Shader / ConnectShaders / ShaderGroupEnd), outputs are placed with add_symlocs
// like setup_output_images (testshade.cpp:1138-1158), the renderer is a RendererServices
// subclass with SimpleRenderer's named transforms and userdata (simplerend.cpp:354-590), and
// shading runs either
//   --batched16 : the reference's batched loop, 16 lanes per BatchedExecutor<16>::execute call
//                 with BatchedShaderGlobals<16> filled lane by lane (testshade.cpp:1829-1922), or
//   (default)   : one whole-grid call, which is what a GPU back end wants.
// Both give the same bytes.  -o writes a colour output as raw float32, -od3 one with
// derivatives (val, dx, dy).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/OSL/oslexec.h"

using namespace OSL;

// SimpleRenderer (src/testshade/simplerend.cpp): named coordinate systems and the hard-wired
// userdata of the test suite
class SimpleRenderer : public RendererServices {
public:
    Matrix44 shader, object, myspace;
    SimpleRenderer()
    {
        // setup_transformations (testshade.cpp:925-950)
        const float c45 = std::cos(float(M_PI / 4)), s45 = std::sin(float(M_PI / 4));
        shader[0][0] = c45; shader[0][1] = s45; shader[1][0] = -s45; shader[1][1] = c45;
        shader[3][0] = 1.0f * c45; shader[3][1] = 1.0f * s45;
        object[0][0] = 0; object[0][1] = 1; object[1][0] = -1; object[1][1] = 0;
        object[3][0] = -1.0f; object[3][1] = 0.0f;
        myspace[1][1] = 2.0f;
    }
    int supports(string_view feature) const override { return feature == "B200"; }
    bool get_matrix(ShaderGlobals*, Matrix44& result, TransformationPtr xform, float) override
    {
        if (!xform)
            return false;
        result = *(const Matrix44*)xform;
        return true;
    }
    bool get_matrix(ShaderGlobals*, Matrix44& result, ustringhash from, float) override
    {
        if (from == ustring("myspace")) {
            result = myspace;
            return true;
        }
        return false;
    }
    bool get_userdata(bool derivatives, ustringhash name, TypeDesc type, ShaderGlobals* sg, void* val) override
    {
        float* f = (float*)val;
        auto put = [&](float v, float dx, float dy) {
            f[0] = v;
            if (derivatives) {
                f[1] = dx;
                f[2] = dy;
            }
            return true;
        };
        if (name == ustring("s") && type == TypeFloat)
            return put(sg->u, sg->dudx, sg->dudy);
        if (name == ustring("t") && type == TypeFloat)
            return put(sg->v, sg->dvdx, sg->dvdy);
        if (name == ustring("red") && type == TypeFloat && sg->P.x > 0.5f)
            return put(sg->u, sg->dudx, sg->dudy);
        if (name == ustring("green") && type == TypeFloat && sg->P.x < 0.5f)
            return put(sg->v, sg->dvdx, sg->dvdy);
        if (name == ustring("blue") && type == TypeFloat && ((static_cast<int>(sg->P.y * 12) % 2) == 0))
            return put(1.0f - sg->u, -sg->dudx, -sg->dudy);
        return false;
    }
};

struct Grid {
    int xres, yres;
    bool center;
    // setup_shaderglobals (testshade.cpp:957-1046)
    void point(int x, int y, ShaderGlobals& sg, const SimpleRenderer& rend) const
    {
        std::memset((void*)&sg, 0, sizeof sg);
        if (center) {
            sg.u = (float)(x + 0.5f) / xres;
            sg.v = (float)(y + 0.5f) / yres;
            sg.dudx = 1.0f / xres;
            sg.dvdy = 1.0f / yres;
        } else {
            sg.u = (xres == 1) ? 0.5f : (float)x / (xres - 1);
            sg.v = (yres == 1) ? 0.5f : (float)y / (yres - 1);
            sg.dudx = 1.0f / std::max(1, xres - 1);
            sg.dvdy = 1.0f / std::max(1, yres - 1);
        }
        sg.P = Vec3(sg.u, sg.v, 1.0f);
        sg.dPdx = Vec3(1.0f / std::max(1, xres - 1), 0, 0);
        sg.dPdy = Vec3(0, 1.0f / std::max(1, yres - 1), 0);
        sg.N = sg.Ng = Vec3(0, 0, 1);
        sg.dPdu = Vec3(1, 0, 0);
        sg.dPdv = Vec3(0, 1, 0);
        sg.surfacearea = 1;
        sg.raytype = 1;   // camera
        sg.shader2common = &rend.shader;
        sg.object2common = &rend.object;
    }
};

int
main(int argc, char** argv)
{
    int xres = 1, yres = 1, iters = 1, fma = 1;
    bool center = false, jitonly = false, batched16 = false;
    SimpleRenderer rend;
    ShadingSystem* shadingsys = new ShadingSystem(&rend, nullptr, nullptr);
    ShaderGroupRef group = shadingsys->ShaderGroupBegin("testshade_b200");
    std::string layername;
    struct Out {
        std::string name, file;
        bool derivs;
    };
    std::vector<Out> outs;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need     = [&](int k) {
            if (i + k >= argc) {
                fprintf(stderr, "missing argument after %s\n", a.c_str());
                exit(2);
            }
        };
        if (a == "-g") { need(2); xres = atoi(argv[++i]); yres = atoi(argv[++i]); }
        else if (a == "--center" || a == "-center") center = true;
        else if (a == "--jitonly") jitonly = true;
        else if (a == "--batched16" || a == "--batched") batched16 = true;
        else if (a == "--iters") { need(1); iters = atoi(argv[++i]); }
        else if (a == "--fma") { need(1); fma = atoi(argv[++i]); }
        else if (a == "--searchpath") { need(1); shadingsys->attribute("searchpath:shader", argv[++i]); }
        else if (a == "--layer" || a == "-layer") { need(1); layername = argv[++i]; }
        else if (a == "--param" || a == "-param") {
            need(2);
            std::string name = argv[++i], val = argv[++i];
            char* end = nullptr;
            float f   = strtof(val.c_str(), &end);
            if (end && *end == 0)
                shadingsys->Parameter(*group, name, TypeFloat, &f);
            else {
                const char* s = val.c_str();
                shadingsys->Parameter(*group, name, TypeString, &s);
            }
        } else if (a == "--connect" || a == "-connect") {
            need(4);
            shadingsys->ConnectShaders(*group, argv[i + 1], argv[i + 2], argv[i + 3], argv[i + 4]);
            i += 4;
        } else if (a == "-o" || a == "-od3") { need(2); outs.push_back({ argv[i + 1], argv[i + 2], a == "-od3" }); i += 2; }
        else {
            if (!shadingsys->Shader(*group, "surface", a, layername)) {
                fprintf(stderr, "ERROR: %s\n", shadingsys->geterror().c_str());
                return 1;
            }
            layername.clear();
        }
    }
    shadingsys->attribute("llvm_jit_fma", fma);
    shadingsys->ShaderGroupEnd(*group);
    // setup_output_images (testshade.cpp:1138-1158): one dense arena per output, offset = running
    // total, stride = element size
    const long long npoints = (long long)xres * yres;
    long long offset        = 0;
    std::vector<SymLocationDesc> locs;
    for (auto& o : outs) {
        const long long stride = o.derivs ? 36 : 12;
        locs.emplace_back(o.name, TypeColor, o.derivs, SymArena::Outputs, offset, stride);
        offset += stride * npoints;
    }
    shadingsys->add_symlocs(group.get(), locs);
    auto t0 = std::chrono::steady_clock::now();
    if (!shadingsys->optimize_group(group.get())) {
        fprintf(stderr, "ERROR: %s\n", shadingsys->geterror().c_str());
        return 1;
    }
    double setup = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Setup (generate + NVRTC): %.3f s, cubin %lld bytes\n", setup, [&] {
        long long n = 0;
        b200_group_cubin(group->handle, &n);
        return n;
    }());
    if (jitonly)
        return 0;
    Grid grid { xres, yres, center };
    PerThreadInfo* thread_info = shadingsys->create_thread_info();
    ShadingContext* ctx        = shadingsys->get_context(thread_info);
    std::vector<float> arena((size_t)offset / 4 + 1);
    constexpr int W = 16;
    auto executor   = shadingsys->batched<W>();
    t0              = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        if (batched16) {
            // batched_shade_region<16> (testshade.cpp:1829-1922)
            BatchedShaderGlobals<W> sgBatch;
            Block<int, W> wide_shadeindex_block;
            sgBatch.uniform.raytype = 1;
            for (long long first = 0; first < npoints; first += W) {
                int batchSize = (int)std::min<long long>(W, npoints - first);
                for (int bi = 0; bi < batchSize; ++bi) {
                    ShaderGlobals sg;
                    long long idx = first + bi;
                    grid.point((int)(idx % xres), (int)(idx / xres), sg, rend);
                    auto& v = sgBatch.varying;
                    v.P[bi] = sg.P; v.dPdx[bi] = sg.dPdx; v.dPdy[bi] = sg.dPdy; v.dPdz[bi] = sg.dPdz;
                    v.I[bi] = sg.I; v.dIdx[bi] = sg.dIdx; v.dIdy[bi] = sg.dIdy; v.N[bi] = sg.N; v.Ng[bi] = sg.Ng;
                    v.u[bi] = sg.u; v.dudx[bi] = sg.dudx; v.dudy[bi] = sg.dudy;
                    v.v[bi] = sg.v; v.dvdx[bi] = sg.dvdx; v.dvdy[bi] = sg.dvdy;
                    v.dPdu[bi] = sg.dPdu; v.dPdv[bi] = sg.dPdv; v.time[bi] = 0; v.dtime[bi] = 0; v.dPdtime[bi] = Vec3(0);
                    v.Ps[bi] = Vec3(0); v.dPsdx[bi] = Vec3(0); v.dPsdy[bi] = Vec3(0);
                    v.object2common[bi] = sg.object2common; v.shader2common[bi] = sg.shader2common;
                    v.surfacearea[bi] = sg.surfacearea; v.flipHandedness[bi] = 0; v.backfacing[bi] = 0;
                    wide_shadeindex_block[bi] = (int)idx;
                }
                if (!executor.execute(*ctx, *group, batchSize, Wide<const int, W>(wide_shadeindex_block), sgBatch, nullptr,
                                      arena.data())) {
                    fprintf(stderr, "ERROR: %s\n", shadingsys->geterror().c_str());
                    return 1;
                }
            }
        } else {
            // one call for the whole grid: SoA planes of the fields that vary
            std::vector<ShaderGlobals> pts(npoints);
            std::vector<float> u(npoints), v(npoints), P(3 * npoints);
            for (long long i = 0; i < npoints; ++i) {
                grid.point((int)(i % xres), (int)(i / xres), pts[i], rend);
                u[i] = pts[i].u; v[i] = pts[i].v;
                P[i] = pts[i].P.x; P[npoints + i] = pts[i].P.y; P[2 * npoints + i] = pts[i].P.z;
            }
            b200_globals sg;
            memset(&sg, 0, sizeof sg);
            sg.plane_stride       = npoints;
            sg.varying[B200_SG_u] = u.data();
            sg.varying[B200_SG_v] = v.data();
            sg.varying[B200_SG_P] = P.data();
            const ShaderGlobals& p0 = pts[0];
            sg.uniform[B200_SG_dudx][0] = p0.dudx;
            sg.uniform[B200_SG_dvdy][0] = p0.dvdy;
            sg.uniform[B200_SG_dPdx][0] = p0.dPdx.x;
            sg.uniform[B200_SG_dPdy][1] = p0.dPdy.y;
            sg.uniform[B200_SG_N][2] = sg.uniform[B200_SG_Ng][2] = 1.0f;
            sg.uniform[B200_SG_dPdu][0] = sg.uniform[B200_SG_dPdv][1] = 1.0f;
            sg.uniform[B200_SG_surfacearea][0] = 1.0f;
            int camera = 1;
            memcpy(&sg.uniform[B200_SG_raytype][0], &camera, 4);
            // named spaces of the renderer, as SimpleRenderer would answer get_matrix
            b200_transform xf[3] = { { "shader", {} }, { "object", {} }, { "myspace", {} } };
            memcpy(xf[0].m, rend.shader.x, 64);
            memcpy(xf[1].m, rend.object.x, 64);
            memcpy(xf[2].m, rend.myspace.x, 64);
            sg.ntransforms = 3;
            sg.transforms  = xf;
            long long rec = 0;
            bool ok;
            if (b200_group_userdata_fields(group->handle, &rec) > 0) {
                // interpolated parameters: the renderer's get_userdata is asked per point, so
                // the batch goes through the lane interface 16 points at a time... unless the
                // renderer describes its arrays with SymArena::UserData symlocs.  Here: lanes.
                fprintf(stderr, "group has interpolated parameters: use --batched16\n");
                return 1;
            }
            ok = executor.execute(*ctx, *group, npoints, 0, sg, nullptr, arena.data());
            if (!ok) {
                fprintf(stderr, "ERROR: %s\n", shadingsys->geterror().c_str());
                return 1;
            }
        }
    }
    double run = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Run  : %.4f s  (%d iters, %.3f Mpoints/s end to end)\n", run, iters, 1e-6 * npoints * iters / run);
    long long off = 0;
    for (size_t k = 0; k < outs.size(); ++k) {
        const long long words = (outs[k].derivs ? 9 : 3) * npoints;
        FILE* f = fopen(outs[k].file.c_str(), "wb");
        if (f) {
            fwrite(arena.data() + off, 4, words, f);
            fclose(f);
            printf("Output %s to %s (%dx%dx%d float32)\n", outs[k].name.c_str(), outs[k].file.c_str(), xres, yres,
                   outs[k].derivs ? 9 : 3);
        }
        off += words;
    }
    shadingsys->release_context(ctx);
    shadingsys->destroy_thread_info(thread_info);
    group.reset();
    delete shadingsys;
    return 0;
}
