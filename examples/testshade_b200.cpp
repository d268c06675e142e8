// testshade_b200 — minimal C++ mirror of the reference's testshade grid harness
// (src/testshade/testshade.cpp) on top of include/OSL/oslexec_b200.h.
//
//   testshade_b200 [-g W H] [--center] [--searchpath DIR] [--fma 0|1] [--iters N]
//                  [--jitonly] [--param name value]... [--layer NAME] shader
//                  [--connect L1 P1 L2 P2]... [-o OUTPUT file.f32]...
//
// Builds the group with the same call sequence as the reference
// (testshade.cpp:2028-2080), fills SoA ShaderGlobals the way
// setup_shaderglobals does (testshade.cpp:957-1046), executes through the
// host-pointer C-ABI path and writes each output as raw float32.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../include/OSL/oslexec_b200.h"

using namespace OSL_B200;

int
main(int argc, char** argv)
{
    int xres = 1, yres = 1, iters = 1, fma = 1;
    bool center = false, jitonly = false;
    ShadingSystem ss;
    ShaderGroupRef group = ss.ShaderGroupBegin("testshade_b200");
    std::string layername;
    struct Out {
        std::string name, file;
    };
    std::vector<Out> outs;
    std::vector<std::string> shaders;
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
        else if (a == "--iters") { need(1); iters = atoi(argv[++i]); }
        else if (a == "--fma") { need(1); fma = atoi(argv[++i]); }
        else if (a == "--searchpath") { need(1); ss.attribute("searchpath:shader", argv[++i]); }
        else if (a == "--layer" || a == "-layer") { need(1); layername = argv[++i]; }
        else if (a == "--param" || a == "-param") {
            need(2);
            std::string name = argv[++i], val = argv[++i];
            char* end = nullptr;
            float f   = strtof(val.c_str(), &end);
            if (end && *end == 0)
                ss.Parameter(*group, name, TypeFloat, &f);
            else {
                const char* s = val.c_str();
                ss.Parameter(*group, name, TypeString, &s);
            }
        } else if (a == "--connect" || a == "-connect") {
            need(4);
            ss.ConnectShaders(*group, argv[i + 1], argv[i + 2], argv[i + 3], argv[i + 4]);
            i += 4;
        } else if (a == "-o") { need(2); outs.push_back({ argv[i + 1], argv[i + 2] }); i += 2; }
        else {
            if (!ss.Shader(*group, "surface", a, layername)) {
                fprintf(stderr, "ERROR: %s\n", ss.geterror().c_str());
                return 1;
            }
            layername.clear();
        }
    }
    ss.attribute("llvm_jit_fma", fma);
    ss.ShaderGroupEnd(*group);
    // setup_output_images (testshade.cpp:1138-1158): one dense arena per output,
    // offset = running total, stride = element size.  Colours assumed for outputs.
    long long npoints = (long long)xres * yres, offset = 0;
    std::vector<SymLocationDesc> locs;
    for (auto& o : outs) {
        locs.emplace_back(o.name, TypeColor, false, SymArena::Outputs, offset, 12);
        offset += 12 * npoints;
    }
    ss.add_symlocs(group.get(), locs.data(), locs.size());
    auto t0 = std::chrono::steady_clock::now();
    if (!ss.optimize_group(group.get())) {
        fprintf(stderr, "ERROR: %s\n", ss.geterror().c_str());
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
    // ShaderGlobals, SoA
    std::vector<float> u(npoints), v(npoints);
    for (int y = 0; y < yres; ++y)
        for (int x = 0; x < xres; ++x) {
            size_t i = (size_t)y * xres + x;
            if (center) {
                u[i] = (float)(x + 0.5f) / xres;
                v[i] = (float)(y + 0.5f) / yres;
            } else {
                u[i] = (xres == 1) ? 0.5f : (float)x / (xres - 1);
                v[i] = (yres == 1) ? 0.5f : (float)y / (yres - 1);
            }
        }
    std::vector<float> P(3 * npoints, 1.0f);
    std::copy(u.begin(), u.end(), P.begin());
    std::copy(v.begin(), v.end(), P.begin() + npoints);
    b200_globals sg;
    memset(&sg, 0, sizeof sg);
    sg.plane_stride       = npoints;
    sg.varying[B200_SG_u] = u.data();
    sg.varying[B200_SG_v] = v.data();
    sg.varying[B200_SG_P] = P.data();
    float du = center ? 1.0f / xres : 1.0f / std::max(1, xres - 1);
    float dv = center ? 1.0f / yres : 1.0f / std::max(1, yres - 1);
    sg.uniform[B200_SG_dudx][0] = du;
    sg.uniform[B200_SG_dvdy][0] = dv;
    sg.uniform[B200_SG_dPdx][0] = 1.0f / std::max(1, xres - 1);
    sg.uniform[B200_SG_dPdy][1] = 1.0f / std::max(1, yres - 1);
    sg.uniform[B200_SG_N][2] = sg.uniform[B200_SG_Ng][2] = 1.0f;
    sg.uniform[B200_SG_dPdu][0] = sg.uniform[B200_SG_dPdv][1] = 1.0f;
    sg.uniform[B200_SG_surfacearea][0] = 1.0f;
    int camera = 1;
    memcpy(&sg.uniform[B200_SG_raytype][0], &camera, 4);
    std::vector<float> arena((size_t)offset / 4 + 1);
    auto exec = ss.batched();
    t0        = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it)
        if (!exec.execute_host(*group, npoints, sg, arena.data())) {
            fprintf(stderr, "ERROR: %s\n", ss.geterror().c_str());
            return 1;
        }
    double run = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("Run  : %.4f s  (%d iters, %.3f Mpoints/s end to end)\n", run, iters, 1e-6 * npoints * iters / run);
    for (size_t k = 0; k < outs.size(); ++k) {
        FILE* f = fopen(outs[k].file.c_str(), "wb");
        if (!f)
            continue;
        fwrite(arena.data() + k * 3 * npoints, 4, 3 * npoints, f);
        fclose(f);
        printf("Output %s to %s (%dx%dx3 float32)\n", outs[k].name.c_str(), outs[k].file.c_str(), xres, yres);
    }
    return 0;
}
