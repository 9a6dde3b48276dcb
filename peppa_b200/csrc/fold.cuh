// Device side of HingeFold (host_util.h): the scalar hinge loss from the similarity pass's CTA partials and the
// indicator counts (pig/loss.py:41-48: the mean over the N x N matrix of both clamps; the diagonal's margin - M_ii
// terms enter through the counts).  256 threads of one block, fixed summation order, fp64 accumulation.
#pragma once
#include "common.cuh"
#include "host_util.h"

namespace pb2 {

// tid in [0, 256): the calling threads; sh / sbad: 8 doubles / 8 ints of shared memory; bar_id: a named barrier free
// for these 256 threads.
__device__ __forceinline__ void hinge_loss_fold(const HingeFold& f, int tid, double* sh, int* sbad, uint32_t bar_id) {
    double acc = 0.0;
    int bad = 0;
    for (int i = tid; i < f.n_partials; i += 256) acc += (double)f.loss_partial[i];
    // four rows per thread and trip: their loads are in flight together; the sum keeps its order (i ascending)
    for (int64_t i0 = tid; i0 < f.n; i0 += 4 * 256) {
        float dg[4], x[4], y[4];
        int cnt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * 256;
            const bool ok = i < f.n;
            dg[u] = ok ? f.diag[i] : 0.f;
            cnt[u] = ok ? f.row_cnt[i] + f.col_cnt[i] : 0;
            x[u] = ok ? f.rinv_v[i] : 1.f;
            y[u] = ok ? f.rinv_a[i] : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (i0 + (int64_t)u * 256 < f.n) {
                acc += (double)(f.margin - dg[u]) * (double)cnt[u];
                bad |= !(fabsf(x[u]) <= 3.0e38f) || !(fabsf(y[u]) <= 3.0e38f);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((tid & 31) == 0) {
        sh[tid >> 5] = acc;
        sbad[tid >> 5] = bad;
    }
    named_bar_sync(bar_id, 256);
    if (tid == 0) {
        double t = 0.0;
        int b = 0;
        for (int w = 0; w < 8; ++w) {
            t += sh[w];
            b |= sbad[w];
        }
        f.loss_out[0] = b ? __int_as_float(0x7fc00000) : (float)(t * (double)f.coef);
    }
}

}  // namespace pb2
