// Scene BVH builder of the testrender path: binned SAH, 16 bins, in-place partition,
// depth-first node numbering.  Behaviour of the reference's builder (src/testrender/bvh.cpp:
// build_bvh / the BuildNode work list) restated with every operation in float32 and in the
// same order, because the traversal order of the tree decides which of two equidistant hits a
// ray reports and the renders are compared bit for bit.
//
// Output layout (osl_b200.h b200_render_scene::bvh_nodes): 8 words per node
//   xmin xmax ymin ymax zmin zmax  child(u32)  nprims(u32)
// inner node: nprims == 0, children at child and child + 1; leaf: primitives
// indices[child .. child + nprims).

#include "../../../include/osl_b200.h"

#include <cstring>
#include <limits>
#include <vector>

namespace {

constexpr int NUM_BINS  = 16;
constexpr int MAX_DEPTH = 64;

struct F3 {
    float v[3];
};
inline float fmin2(float a, float b) { return b < a ? b : a; }
inline float fmax2(float a, float b) { return b > a ? b : a; }
struct Box {
    F3 lo, hi;
    void reset()
    {
        const float inf = std::numeric_limits<float>::infinity();
        for (int k = 0; k < 3; ++k) {
            lo.v[k] = inf;
            hi.v[k] = -inf;
        }
    }
    void grow(const F3& a, const F3& b)
    {
        for (int k = 0; k < 3; ++k) {
            lo.v[k] = fmin2(lo.v[k], a.v[k]);
            hi.v[k] = fmax2(hi.v[k], b.v[k]);
        }
    }
    void grow(const Box& b) { grow(b.lo, b.hi); }
};
inline float
half_area(const F3& lo, const F3& hi)
{
    const float d0 = hi.v[0] - lo.v[0], d1 = hi.v[1] - lo.v[1], d2 = hi.v[2] - lo.v[2];
    const float a = d0 * d1;
    const float b = d1 * d2;
    const float c = d2 * d0;
    const float ab = a + b;
    return ab + c;
}
struct Work {
    F3 cmin, cmax;  // bounds of the primitives' centroids
    int left, right, depth, node;
};

}  // namespace

extern "C" int
b200_build_bvh(const float* verts, int nverts, const int* triangles, int ntriangles, float* nodes, int max_nodes,
               unsigned* indices, int* nnodes)
{
    if (!verts || !triangles || !nodes || !indices || !nnodes || ntriangles <= 0 || max_nodes < 1)
        return B200_ERR_INVALID;
    const int n = ntriangles;
    std::vector<F3> bmin(n), bmax(n), cen(n);
    for (int t = 0; t < n; ++t) {
        Box b;
        b.reset();
        for (int c = 0; c < 3; ++c) {
            const int vi = triangles[3 * t + c];
            if (vi < 0 || vi >= nverts)
                return B200_ERR_INVALID;
            F3 p;
            memcpy(p.v, verts + 3 * (size_t)vi, sizeof p.v);
            b.grow(p, p);
        }
        bmin[t] = b.lo;
        bmax[t] = b.hi;
        for (int k = 0; k < 3; ++k)
            cen[t].v[k] = (b.lo.v[k] + b.hi.v[k]) * 0.5f;
        indices[t] = (unsigned)t;
    }
    int count = 1;
    auto set_node = [&](int i, const Box& b, unsigned child, unsigned nprims) {
        float* q = nodes + 8 * (size_t)i;
        q[0] = b.lo.v[0]; q[1] = b.hi.v[0]; q[2] = b.lo.v[1]; q[3] = b.hi.v[1]; q[4] = b.lo.v[2]; q[5] = b.hi.v[2];
        memcpy(q + 6, &child, 4);
        memcpy(q + 7, &nprims, 4);
    };
    Work cur;
    {
        Box root, c;
        root.reset();
        c.reset();
        for (int t = 0; t < n; ++t) {
            root.grow(bmin[t], bmax[t]);
            c.grow(cen[t], cen[t]);
        }
        set_node(0, root, 0, 0);
        cur = { c.lo, c.hi, 0, n, 1, 0 };
    }
    std::vector<Work> stack;
    const float bin_scale = (float)(0.999 * NUM_BINS);
    for (;;) {
        const int left = cur.left, right = cur.right, nprims = right - left;
        bool split = false;
        if (nprims > 1 && cur.depth < MAX_DEPTH) {
            float binf[3];
            for (int k = 0; k < 3; ++k) {
                const float ext = cur.cmax.v[k] - cur.cmin.v[k];
                binf[k]         = ext > 0 ? bin_scale / ext : 0.0f;
            }
            const float* nb = nodes + 8 * (size_t)cur.node;
            const F3 nlo = { { nb[0], nb[2], nb[4] } }, nhi = { { nb[1], nb[3], nb[5] } };
            const float inv_area = 1.0f / half_area(nlo, nhi);
            float best_cost      = (float)nprims;
            int best_axis = -1, best_bin = -1;
            for (int axis = 0; axis < 3; ++axis) {
                if (binf[axis] == 0)
                    continue;
                int cnt[NUM_BINS] = { 0 };
                Box bb[NUM_BINS];
                for (int i = 0; i < NUM_BINS; ++i)
                    bb[i].reset();
                for (int q = left; q < right; ++q) {
                    const unsigned prim = indices[q];
                    const float off     = cen[prim].v[axis] - cur.cmin.v[axis];
                    const float scaled  = off * binf[axis];
                    const int id        = (int)scaled;
                    cnt[id]++;
                    bb[id].grow(bmin[prim], bmax[prim]);
                }
                int numL[NUM_BINS];
                Box accl[NUM_BINS];
                int run = 0;
                Box acc;
                acc.reset();
                for (int i = 0; i < NUM_BINS; ++i) {
                    run += cnt[i];
                    numL[i] = run;
                    acc.grow(bb[i]);
                    accl[i] = acc;
                }
                Box rb = bb[NUM_BINS - 1];
                for (int i = NUM_BINS - 2; i >= 0; --i) {
                    if (numL[i] == 0 || numL[i] == nprims)
                        continue;
                    const float areaR = half_area(rb.lo, rb.hi);
                    const float areaL = half_area(accl[i].lo, accl[i].hi);
                    const float tl    = areaL * (float)numL[i];
                    const float tr    = areaR * (float)(nprims - numL[i]);
                    const float sum   = tl + tr;
                    const float scl   = inv_area * sum;
                    const float cost  = 4.0f + scl;
                    if (cost < best_cost) {
                        best_cost = cost;
                        best_axis = axis;
                        best_bin  = i;
                    }
                    rb.grow(bb[i]);
                }
            }
            if (best_axis >= 0) {
                int i = left, r = right;
                while (i < r) {
                    const unsigned prim = indices[i];
                    const float off     = cen[prim].v[best_axis] - cur.cmin.v[best_axis];
                    const float scaled  = off * binf[best_axis];
                    if ((int)scaled <= best_bin)
                        ++i;
                    else {
                        --r;
                        const unsigned tmp = indices[i];
                        indices[i]         = indices[r];
                        indices[r]         = tmp;
                    }
                }
                const int mid = r;
                if (count + 2 > max_nodes)
                    return B200_ERR_INVALID;
                const int nxt = count;
                count += 2;
                unsigned child = (unsigned)nxt, zero = 0;
                memcpy(nodes + 8 * (size_t)cur.node + 6, &child, 4);
                memcpy(nodes + 8 * (size_t)cur.node + 7, &zero, 4);
                Box lb, rb2, lc, rc;
                lb.reset(); rb2.reset(); lc.reset(); rc.reset();
                for (int q = left; q < mid; ++q) {
                    lb.grow(bmin[indices[q]], bmax[indices[q]]);
                    lc.grow(cen[indices[q]], cen[indices[q]]);
                }
                for (int q = mid; q < right; ++q) {
                    rb2.grow(bmin[indices[q]], bmax[indices[q]]);
                    rc.grow(cen[indices[q]], cen[indices[q]]);
                }
                set_node(nxt, lb, 0, 0);
                set_node(nxt + 1, rb2, 0, 0);
                stack.push_back({ rc.lo, rc.hi, mid, right, cur.depth + 1, nxt + 1 });
                cur   = { lc.lo, lc.hi, left, mid, cur.depth + 1, nxt };
                split = true;
            }
        }
        if (split)
            continue;
        unsigned child = (unsigned)left, np = (unsigned)nprims;
        memcpy(nodes + 8 * (size_t)cur.node + 6, &child, 4);
        memcpy(nodes + 8 * (size_t)cur.node + 7, &np, 4);
        if (stack.empty())
            break;
        cur = stack.back();
        stack.pop_back();
    }
    *nnodes = count;
    return B200_OK;
}
