// oracle_shadeops.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// Batch entry points over SoA planes for the scalar shadeop restatement in
// osl_oracle.h / osl_oracle_ops.h, with the same argument layout as
// b200_shadeop_noise / b200_shadeop_hash in include/osl_b200.h so a test can
// feed both sides the same buffers.  Reference signatures being restated:
// osl_<noise>_<codes> (src/liboslexec/opnoise.cpp:263-273, 468-473),
// osl_hash_* (src/liboslexec/builtindecl.h:169-173).
#include "osl_oracle_ops.h"

using namespace oslo;

namespace {

template<int KIND, int NC>
void
run_noise(int indim, bool derivs, long long n, const float* in, const float* period, float* out)
{
    for (long long i = 0; i < n; ++i) {
        float x[4] = { 0, 0, 0, 0 };
        for (int d = 0; d < indim; ++d)
            x[d] = in[d * n + i];
        if (KIND == N_CELL || KIND == N_HASH) {
            if (period)
                for (int d = 0; d < indim; ++d)
                    x[d] = pwrap(x[d], period[d]);
            float r[3];
            ihnoise_core<KIND, NC>(r, indim, x);
            for (int c = 0; c < NC; ++c)
                out[c * n + i] = r[c];
            if (derivs)
                for (int c = 0; c < 2 * NC; ++c)
                    out[(NC + c) * n + i] = 0.0f;
            continue;
        }
        int per[4] = { 1, 1, 1, 1 };
        if (period)
            for (int d = 0; d < indim; ++d)
                per[d] = iperiod(period[d]);
        const int* pp = period ? per : nullptr;
        if (derivs) {
            Df xd[4], r[3];
            for (int d = 0; d < indim; ++d)
                xd[d] = Df(x[d], in[(indim + d) * n + i], in[(2 * indim + d) * n + i]);
            if (KIND == N_SIMPLEX || KIND == N_USIMPLEX)
                noise_core<KIND, Df, NC>(r, indim, xd);
            else if (KIND == N_NOISE)
                perlin_nd<Df, NC, false>(r, indim, xd, pp);
            else
                perlin_nd<Df, NC, true>(r, indim, xd, pp);
            for (int c = 0; c < NC; ++c) {
                out[c * n + i]            = r[c].val;
                out[(NC + c) * n + i]     = r[c].dx;
                out[(2 * NC + c) * n + i] = r[c].dy;
            }
        } else {
            float r[3];
            if (KIND == N_SIMPLEX || KIND == N_USIMPLEX)
                noise_core<KIND, float, NC>(r, indim, x);
            else if (KIND == N_NOISE)
                perlin_nd<float, NC, false>(r, indim, x, pp);
            else
                perlin_nd<float, NC, true>(r, indim, x, pp);
            for (int c = 0; c < NC; ++c)
                out[c * n + i] = r[c];
        }
    }
}

}  // namespace

extern "C" int
oracle_noise(int kind, int outdim, int indim, int derivs, long long n, const float* in,
             const float* period, float* out)
{
    if (kind < 0 || kind > 5 || (outdim != 1 && outdim != 3) || indim < 1 || indim > 4)
        return 1;
#define GO(K)                                                            \
    if (outdim == 1)                                                     \
        run_noise<K, 1>(indim, derivs != 0, n, in, period, out);         \
    else                                                                 \
        run_noise<K, 3>(indim, derivs != 0, n, in, period, out);
    switch (kind) {
    case 0: GO(N_NOISE) break;
    case 1: GO(N_SNOISE) break;
    case 2: GO(N_CELL) break;
    case 3: GO(N_HASH) break;
    case 4: GO(N_SIMPLEX) break;
    default: GO(N_USIMPLEX) break;
    }
#undef GO
    return 0;
}

extern "C" int
oracle_hash(int indim, long long n, const float* in, int* out)
{
    for (long long i = 0; i < n; ++i) {
        float x[4] = { 0, 0, 0, 0 };
        for (int d = 0; d < indim; ++d)
            x[d] = in[d * n + i];
        out[i] = indim == 1   ? hash_f(x[0])
                 : indim == 2 ? hash_ff(x[0], x[1])
                 : indim == 3 ? hash_v(V3(x[0], x[1], x[2]))
                              : hash_vf(V3(x[0], x[1], x[2]), x[3]);
    }
    return 0;
}
