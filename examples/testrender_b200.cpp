// testrender_b200 — C++ driver of the wavefront path tracer through the C ABI (b200_render_*),
// the hand-over a renderer like the reference's testrender makes after it has parsed its scene
// and built its BVH (src/testrender/simpleraytracer.cpp:281-512, 1221-1272; bvh.cpp:43-219).
//
//   testrender_b200 scene.b200scene out.f32 [options]
//
// scene.b200scene is the prepared scene as a flat binary (written by
// openshadinglanguage_b200.render.scene / tests/test_cpp_api.py: the arrays of
// b200_render_scene plus one group description per material); out.f32 receives yres*xres*3
// float32.  The frame is rendered as interleaved 64x64 tiles in two work sets, the way two GPUs
// would share it, and reassembled here.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/osl_b200.h"

namespace {
struct Reader {
    std::vector<char> buf;
    size_t pos = 0;
    template<class T> T get()
    {
        T v;
        memcpy(&v, buf.data() + pos, sizeof(T));
        pos += sizeof(T);
        return v;
    }
    std::string str()
    {
        int n = get<int>();
        std::string s(buf.data() + pos, buf.data() + pos + n);
        pos += n;
        return s;
    }
    template<class T> std::vector<T> arr()
    {
        long long n = get<long long>();
        std::vector<T> v((size_t)n);
        memcpy(v.data(), buf.data() + pos, (size_t)n * sizeof(T));
        pos += (size_t)n * sizeof(T);
        return v;
    }
};
}  // namespace

int
main(int argc, char** argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s scene.b200scene out.f32 [options]\n", argv[0]);
        return 2;
    }
    Reader R;
    {
        FILE* f = fopen(argv[1], "rb");
        if (!f) {
            fprintf(stderr, "cannot open %s\n", argv[1]);
            return 1;
        }
        fseek(f, 0, SEEK_END);
        R.buf.resize((size_t)ftell(f));
        fseek(f, 0, SEEK_SET);
        if (fread(R.buf.data(), 1, R.buf.size(), f) != R.buf.size())
            return 1;
        fclose(f);
    }
    if (R.get<int>() != 0x42323030) {
        fprintf(stderr, "not a b200scene file\n");
        return 1;
    }
    b200_render_scene S;
    memset(&S, 0, sizeof S);
    auto verts = R.arr<float>(), normals = R.arr<float>(), uvs = R.arr<float>();
    auto tris = R.arr<int>(), ntris = R.arr<int>(), uvtris = R.arr<int>(), shaderids = R.arr<int>(), meshids = R.arr<int>();
    auto area = R.arr<float>(), nodes = R.arr<float>();
    auto indices = R.arr<unsigned>(), lightprims = R.arr<unsigned>();
    auto is_light = R.arr<int>();
    S.nverts = (int)verts.size() / 3; S.ntris = (int)tris.size() / 3; S.nnodes = (int)nodes.size() / 8;
    S.nlightprims = (int)lightprims.size(); S.nshaders = (int)is_light.size(); S.nmeshes = (int)area.size();
    S.verts = verts.data(); S.normals = normals.data(); S.uvs = uvs.data(); S.triangles = tris.data();
    S.n_triangles = ntris.data(); S.uv_triangles = uvtris.data(); S.shaderids = shaderids.data(); S.meshids = meshids.data();
    S.mesh_surfacearea = area.data(); S.bvh_nodes = nodes.data(); S.bvh_indices = indices.data();
    S.lightprims = lightprims.data(); S.shader_is_light = is_light.data();
    for (int k = 0; k < 3; ++k) S.eye[k] = R.get<float>();
    for (int k = 0; k < 3; ++k) S.dir[k] = R.get<float>();
    for (int k = 0; k < 3; ++k) S.up[k] = R.get<float>();
    S.fov = R.get<float>();
    S.xres = R.get<int>(); S.yres = R.get<int>(); S.aa = R.get<int>(); S.max_bounces = R.get<int>();
    S.rr_depth = R.get<int>(); S.no_jitter = R.get<int>(); S.show_globals = R.get<int>();
    S.background_shader = R.get<int>(); S.background_resolution = R.get<int>();
    // materials: ShaderGroupBegin ... ShaderGroupEnd per material, as data
    const int nmat = R.get<int>();
    struct Mat {
        std::vector<std::string> oso, lname, sval;
        std::vector<std::vector<b200_param>> params;
        std::vector<std::vector<std::string>> pnames;
        std::vector<std::vector<std::vector<float>>> fvals;
        std::vector<std::vector<std::vector<int>>> ivals;
        std::vector<std::vector<std::vector<const char*>>> svals;
        std::vector<std::vector<std::vector<std::string>>> sstore;
        std::vector<b200_layer> layers;
        std::vector<std::string> cs;
        std::vector<b200_connection> conns;
    };
    std::vector<Mat> mats(nmat);
    std::vector<b200_group_desc> descs(nmat);
    for (int m = 0; m < nmat; ++m) {
        Mat& M = mats[m];
        const int nl = R.get<int>();
        M.params.resize(nl); M.pnames.resize(nl); M.fvals.resize(nl); M.ivals.resize(nl); M.svals.resize(nl); M.sstore.resize(nl);
        for (int l = 0; l < nl; ++l) {
            M.oso.push_back(R.str());
            M.lname.push_back(R.str());
            const int np = R.get<int>();
            M.pnames[l].resize(np); M.fvals[l].resize(np); M.ivals[l].resize(np); M.svals[l].resize(np); M.sstore[l].resize(np);
            for (int p = 0; p < np; ++p) {
                M.pnames[l][p] = R.str();
                const int type = R.get<int>(), n = R.get<int>();
                b200_param bp { nullptr, type, n, nullptr };
                if (type == 0) {
                    for (int k = 0; k < n; ++k) M.ivals[l][p].push_back(R.get<int>());
                } else if (type == 1) {
                    for (int k = 0; k < n; ++k) M.fvals[l][p].push_back(R.get<float>());
                } else {
                    for (int k = 0; k < n; ++k) M.sstore[l][p].push_back(R.str());
                }
                M.params[l].push_back(bp);
            }
        }
        const int nc = R.get<int>();
        for (int c = 0; c < 4 * nc; ++c)
            M.cs.push_back(R.str());
        // pointers after all the storage stopped moving
        for (int l = 0; l < nl; ++l) {
            for (size_t p = 0; p < M.params[l].size(); ++p) {
                b200_param& bp = M.params[l][p];
                bp.name        = M.pnames[l][p].c_str();
                if (bp.type == 0) bp.values = M.ivals[l][p].data();
                else if (bp.type == 1) bp.values = M.fvals[l][p].data();
                else {
                    for (auto& s : M.sstore[l][p]) M.svals[l][p].push_back(s.c_str());
                    bp.values = M.svals[l][p].data();
                }
            }
            M.layers.push_back({ M.oso[l].c_str(), M.lname[l].c_str(), (int)M.params[l].size(), M.params[l].data() });
        }
        for (int c = 0; c < nc; ++c)
            M.conns.push_back({ M.cs[4 * c].c_str(), M.cs[4 * c + 1].c_str(), M.cs[4 * c + 2].c_str(), M.cs[4 * c + 3].c_str() });
        memset(&descs[m], 0, sizeof descs[m]);
        descs[m].name = "material"; descs[m].nlayers = nl; descs[m].layers = M.layers.data();
        descs[m].nconnections = nc; descs[m].connections = M.conns.data(); descs[m].options = "";
    }
    b200_render* r = nullptr;
    if (b200_render_create(&S, nmat, descs.data(), argc > 3 ? argv[3] : "fma=0", &r) != B200_OK) {
        fprintf(stderr, "ERROR: %s\n", b200_last_error());
        return 1;
    }
    // two interleaved tile work sets, reassembled
    std::vector<int> sets[2];
    int k = 0;
    for (int y = 0; y < S.yres; y += 64)
        for (int x = 0; x < S.xres; x += 64, ++k) {
            int t[4] = { x, y, S.xres - x < 64 ? S.xres - x : 64, S.yres - y < 64 ? S.yres - y : 64 };
            sets[k & 1].insert(sets[k & 1].end(), t, t + 4);
        }
    std::vector<float> img((size_t)S.xres * S.yres * 3);
    long long paths = 0;
    double ms       = 0;
    for (int w = 0; w < 2; ++w) {
        const int nt = (int)sets[w].size() / 4;
        if (!nt)
            continue;
        long long npix = 0;
        for (int t = 0; t < nt; ++t) npix += (long long)sets[w][4 * t + 2] * sets[w][4 * t + 3];
        std::vector<float> strip((size_t)npix * 3);
        b200_render_stats st;
        if (b200_render_tiles(r, 0, nt, sets[w].data(), strip.data(), 0, &st) != B200_OK) {
            fprintf(stderr, "ERROR: %s\n", b200_last_error());
            return 1;
        }
        paths += st.paths;
        ms += st.device_ms;
        size_t i = 0;
        for (int t = 0; t < nt; ++t) {
            const int* q = &sets[w][4 * t];
            for (int y = q[1]; y < q[1] + q[3]; ++y)
                for (int x = q[0]; x < q[0] + q[2]; ++x, i += 3)
                    memcpy(&img[((size_t)y * S.xres + x) * 3], &strip[i], 12);
        }
    }
    printf("Rendered %lld paths in %.2f ms on the device (%.1f Mpaths/s)\n", paths, ms, paths / ms * 1e-3);
    FILE* f = fopen(argv[2], "wb");
    if (!f)
        return 1;
    fwrite(img.data(), 4, img.size(), f);
    fclose(f);
    b200_render_destroy(r);
    return 0;
}
