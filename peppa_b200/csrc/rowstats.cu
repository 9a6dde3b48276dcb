// Light O(N*D) / O(N) kernels around the tensor-core passes: row norms (pig/util.py:11-12), paired
// dot products (the diagonal of pig/loss.py:43 and the positive's score of pig/metrics.py:8-20),
// merges of log-sum-exp partials, the normalisation Jacobian after the gradient GEMMs, and the
// elementwise contrastive(M) of pig/loss.py:41-48 for callers that hold a materialised matrix.
// All are HBM/L2-bound: one warp per row with 16-byte loads.
#include "common.cuh"
#include "fold.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kNegInf = -__builtin_huge_valf();

__device__ __forceinline__ void bf16x8_to_f32(const uint4& u, float (&f)[8]) {
    f[0] = __uint_as_float(u.x << 16);
    f[1] = __uint_as_float(u.x & 0xffff0000u);
    f[2] = __uint_as_float(u.y << 16);
    f[3] = __uint_as_float(u.y & 0xffff0000u);
    f[4] = __uint_as_float(u.z << 16);
    f[5] = __uint_as_float(u.z & 0xffff0000u);
    f[6] = __uint_as_float(u.w << 16);
    f[7] = __uint_as_float(u.w & 0xffff0000u);
}

// Eight consecutive row elements as fp32, for the three input dtypes the reference's callers use (fp32 tensors,
// fp16 under Lightning's `precision: 16`, bf16): the row-wise kernels read the TRUE input values -- only the tensor-
// core operands are 16-bit (an fp32 matrix goes there as a split-fp16 pair, pb2_split_f16).
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
    bf16x8_to_f32(*reinterpret_cast<const uint4*>(p), f);
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&f)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 t = __half22float2(h[e]);
        f[2 * e] = t.x;
        f[2 * e + 1] = t.y;
    }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------------------------- row norms
template <typename T>
__global__ void __launch_bounds__(256) row_norms_kernel(const T* __restrict__ x, int64_t n, int dim,
                                                        int64_t ld, float* __restrict__ rinv,
                                                        float* __restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n; r += nwarps) {
        const T* row = x + r * ld;
        float ss = 0.f;
        for (int k = lane * 8; k < dim; k += 256) {
            float f[8];
            load8(row + k, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
        }
        ss = warp_sum(ss);
        if (lane == 0) {
            const float nr = sqrtf(ss);
            if (norm) norm[r] = nr;
            if (rinv) rinv[r] = 1.0f / nr;  // no epsilon, like the reference: zero row -> inf -> NaN scores
        }
    }
}

// ------------------------------------------------------------------- normalised fp16 copy
// out = fp16(x * rinv): the B operand of the gradient GEMMs.  tcgen05 kind::f16 cannot mix an fp16
// A with a bf16 B, so the gradient matrix (fp16, exact small integers for the hinge loss) is paired
// with an fp16 copy of the normalised embeddings (|.| <= 1, 11-bit significand: rounding 2^-12).
template <typename T>
__global__ void __launch_bounds__(256)
    rows_scale_f16_kernel(const T* __restrict__ x, const float* __restrict__ rinv, int64_t n, int dim,
                          int64_t ld, __half* __restrict__ out, int64_t ld_out) {
    const int vec_per_row = dim / 8;
    const int64_t total = n * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row;
        const int d = (int)(i % vec_per_row) * 8;
        float f[8];
        load8(x + r * ld + d, f);
        const float s = rinv ? rinv[r] : 1.f;
        __half2 h[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(f[2 * e] * s, f[2 * e + 1] * s);
        *reinterpret_cast<uint4*>(out + r * ld_out + d) = *reinterpret_cast<const uint4*>(h);
    }
}

// Two-plane 8-bit embedding operand of the kind::i8 gradient GEMMs: q = round(x * rinv * 32512) in [-32512, 32512]
// (|x * rinv| <= 1 up to rounding; clamped), q = 256 hi + lo with hi in [-127, 127] (s8) and lo in [0, 255] (u8);
// out row = [hi plane (dim bytes) | lo plane (dim bytes)].  16 bits with one scale for the tensor: the quantisation
// step 3e-5 is ~7e-4 of a typical component of a 512-d unit vector and averages out over the >= 10^4 terms of a
// gradient row (measured: DESIGN section 4.2).
template <typename T>
__global__ void __launch_bounds__(256)
    rows_quant_i8_kernel(const T* __restrict__ x, const float* __restrict__ rinv, int64_t n, int dim, int64_t ld,
                         uint8_t* __restrict__ out, int64_t ld_out) {
    const int vec_per_row = dim / 8;
    const int64_t total = n * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row;
        const int d = (int)(i % vec_per_row) * 8;
        float f[8];
        load8(x + r * ld + d, f);
        const float s = (rinv ? rinv[r] : 1.f) * 32512.0f;
        uint32_t hi[2] = {0u, 0u}, lo[2] = {0u, 0u};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float qf = fminf(fmaxf(rintf(f[e] * s), -32512.0f), 32512.0f);  // NaN (zero-norm row) -> -32512: any finite value
            const int q = (int)qf;
            hi[e >> 2] |= ((uint32_t)(q >> 8) & 0xffu) << (8 * (e & 3));   // arithmetic shift = floor division
            lo[e >> 2] |= ((uint32_t)q & 0xffu) << (8 * (e & 3));
        }
        *reinterpret_cast<uint2*>(out + r * ld_out + d) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(out + r * ld_out + dim + d) = make_uint2(lo[0], lo[1]);
    }
}

// y0 = T(x0 * c), y1 = T(x1 * c) with c read on the device: the backward of a scalar loss whose gradients were
// produced in the forward (autograd hands grad_output over as a device scalar).  The saved gradients are fp32 and
// the product is rounded to the inputs' dtype LAST, like the reference's autograd under AMP: with a GradScaler
// (grad_output = 65536) an fp16 gradient of ~1e-7 must not pass through fp16 before the scale is applied.
// One launch for both gradients, 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(256)
    scale_pair_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int64_t n_vec, const float* __restrict__ coef,
                      T* __restrict__ y0, T* __restrict__ y1) {
    constexpr int kPer = 16 / sizeof(T);
    pdl_launch_dependents();
    pdl_wait();  // launched with programmatic stream serialization: scheduled while its predecessor drains
    const float c = *coef;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const bool second = i >= n_vec;
        const int64_t k = second ? i - n_vec : i;
        const float* src = (second ? x1 : x0) + k * kPer;
        T v[kPer];
#pragma unroll
        for (int e = 0; e < kPer; e += 4) {
            const float4 f = *reinterpret_cast<const float4*>(src + e);
            v[e] = (T)(f.x * c);
            v[e + 1] = (T)(f.y * c);
            v[e + 2] = (T)(f.z * c);
            v[e + 3] = (T)(f.w * c);
        }
        *reinterpret_cast<uint4*>((second ? y1 : y0) + k * kPer) = *reinterpret_cast<const uint4*>(v);
    }
}

// ----------------------------------------------------------------------------------- pair dot
template <typename T>
__global__ void __launch_bounds__(256)
    pair_dot_kernel(const T* __restrict__ x, const T* __restrict__ y,
                    const int64_t* __restrict__ ix, const int64_t* __restrict__ iy, const float* __restrict__ rinv_x,
                    const float* __restrict__ rinv_y, int64_t n, int dim, int64_t ldx, int64_t ldy,
                    float* __restrict__ out, float* __restrict__ dist_out, float* __restrict__ thr_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t k = warp; k < n; k += nwarps) {
        const int64_t rx = ix ? ix[k] : k, ry = iy ? iy[k] : k;
        const T* px = x + rx * ldx;
        const T* py = y + ry * ldy;
        float acc = 0.f;
        for (int d = lane * 8; d < dim; d += 256) {
            float a[8], b[8];
            load8(px + d, a);
            load8(py + d, b);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc = fmaf(a[e], b[e], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            // same operation order as the rank epilogue (sim.cu), no FMA contraction
            const float s = __fmul_rn(__fmul_rn(acc, rinv_x ? rinv_x[rx] : 1.f), rinv_y ? rinv_y[ry] : 1.f);
            if (out) out[k] = s;
            const float d = __fsub_rn(1.0f, s);
            if (dist_out) dist_out[k] = d;
            if (thr_out) thr_out[k] = rank_threshold(d);
        }
    }
}

// ---------------------------------------------------------------------------------- LSE merge
__global__ void __launch_bounds__(256) lse_merge_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                                        int n_parts, int64_t rows, float* __restrict__ lse,
                                                        int accumulate) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float m = kNegInf;
    for (int k = 0; k < n_parts; ++k) m = fmaxf(m, pmax[(int64_t)k * rows + r]);
    float s = 0.f;
    if (m > kNegInf)
        for (int k = 0; k < n_parts; ++k) s += psum[(int64_t)k * rows + r] * exp2f(pmax[(int64_t)k * rows + r] - m);
    float v = (m > kNegInf) ? (m + log2f(s)) * kLn2 : kNegInf;
    if (accumulate) {
        const float o = lse[r];
        const float hi = fmaxf(o, v), lo = fminf(o, v);
        v = (hi > kNegInf) ? hi + log1pf(expf(lo - hi)) : kNegInf;
    }
    lse[r] = v;
}

// partial sums of 2^(t - M) with one shift M for every partial (pb2_sim_lse_both): lse = (M + log2 sum) ln 2
__global__ void __launch_bounds__(64) lse_merge_const_kernel(const float* __restrict__ psum, int n_parts, int64_t n,
                                                              float shift, float* __restrict__ lse, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // four chains in a fixed order; sixteen loads in flight per thread (a 32768-wide merge has only 128 CTAs, so
    // the bandwidth has to come from memory-level parallelism inside the thread)
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int k = 0;
    for (; k + 16 <= n_parts; k += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldg(psum + (int64_t)(k + u) * n + i);
#pragma unroll
        for (int u = 0; u < 16; u += 4) {
            s0 += v[u];
            s1 += v[u + 1];
            s2 += v[u + 2];
            s3 += v[u + 3];
        }
    }
    for (; k + 4 <= n_parts; k += 4) {
        s0 += psum[(int64_t)k * n + i];
        s1 += psum[(int64_t)(k + 1) * n + i];
        s2 += psum[(int64_t)(k + 2) * n + i];
        s3 += psum[(int64_t)(k + 3) * n + i];
    }
    for (; k < n_parts; ++k) s0 += psum[(int64_t)k * n + i];
    const float s = (s0 + s1) + (s2 + s3);
    float v = (s > 0.f) ? (shift + log2f(s)) * kLn2 : (s == 0.f ? kNegInf : s);  // NaN / inf propagate
    if (accumulate) {
        const float o = lse[i];
        const float hi = fmaxf(o, v), lo = fminf(o, v);
        v = (v != v) ? v : ((hi > kNegInf) ? hi + log1pf(expf(lo - hi)) : kNegInf);
    }
    lse[i] = v;
}

// out[i] = log sum_k exp(parts[k, i]) (natural log): merges per-rank column log-sum-exp partials
__global__ void __launch_bounds__(256) lse_combine_kernel(const float* __restrict__ parts, int n_parts, int64_t n,
                                                          float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m = kNegInf;
    for (int k = 0; k < n_parts; ++k) m = fmaxf(m, parts[(int64_t)k * n + i]);
    float s = 0.f;
    if (m > kNegInf)
        for (int k = 0; k < n_parts; ++k) s += expf(parts[(int64_t)k * n + i] - m);
    out[i] = (m > kNegInf) ? m + logf(s) : kNegInf;
}

// fixed-order single-block sum -> deterministic
__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* __restrict__ p, int n, float alpha,
                                                            float* __restrict__ out) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += p[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v * alpha;
    }
}

// out (=|+=) alpha * ( sum partials + sum_k (margin - diag[k]) * cnt[k] ), double accumulation, fixed order
__global__ void __launch_bounds__(1024)
    hinge_loss_terms_kernel(const float* __restrict__ partials, int n_partials, const float* __restrict__ diag,
                            const int32_t* __restrict__ cnt, int64_t n, float margin, float alpha,
                            float* __restrict__ out, int accumulate) {
    __shared__ double sh[32];
    double acc = 0.0;
    if (partials)
        for (int i = threadIdx.x; i < n_partials; i += blockDim.x) acc += (double)partials[i];
    if (diag && cnt)
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += (double)(margin - diag[i]) * (double)cnt[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += sh[w];
        const float r = (float)(t * (double)alpha);
        out[0] = accumulate ? out[0] + r : r;
    }
}

__global__ void __launch_bounds__(1024) milnce_loss_kernel(const float* __restrict__ lse_row,
                                                           const float* __restrict__ lse_col,
                                                           const float* __restrict__ diag, int64_t n,
                                                           float* __restrict__ den, float* __restrict__ out) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = lse_row[i], b = lse_col[i];
        const float hi = fmaxf(a, b), lo = fminf(a, b);
        const float d = hi + log1pf(expf(lo - hi));
        den[i] = d;
        acc += d - diag[i];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = sh[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v / (float)n;
    }
}

// ------------------------------------------------------------------- split-fp16 operands (fp32 inputs)
// An fp32 matrix reaches the 16-bit tensor cores as x' = hi + lo with hi = fp16(x'), lo = fp16(x' - hi) (22 significant
// bits), where x' = x * r * 2^e is the row scaled (r = its 1/||x|| if given) and shifted so that its largest component
// lies in [512, 1024): fp16's range is never an issue and lo never goes subnormal for components that matter.  Then
//     <x'_i, y'_j> ~= <hi_x, hi_y> + <lo_x, hi_y> + <hi_x, lo_y>
// is ONE kind::f16 GEMM of contraction length 3 D over the concatenated rows X' = [hi | lo | hi] (side 0) and
// Y' = [hi | hi | lo] (side 1) -- no kernel of sim.cu changes, the fp32 accumulator sums the three products -- and the
// epilogue's per-row factors are the exact powers of two 2^-e (scale_out) instead of 1/||x||.  The dropped lo*lo term
// and lo's own rounding are ~2^-22 per product: ~2e-8 absolute on a cosine of 512-d unit vectors, far inside the 1e-6
// tie window of the rank parity bar (a bf16 split, 16 bits, would leave ~7e-7 and fail it -- measured).
__device__ __forceinline__ void split8_f16(const float (&x)[8], uint4& hi, uint4& lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        h[e] = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
        const float2 hf = __half22float2(h[e]);
        l[e] = __floats2half2_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
    }
    hi = *reinterpret_cast<const uint4*>(h);
    lo = *reinterpret_cast<const uint4*>(l);
}
// out row = [hi | lo | hi] (side 0) or [hi | hi | lo] (side 1), each part `dim` wide
__device__ __forceinline__ void store_split(__half* out_row, int dim, int d, int side, const uint4& hi, const uint4& lo) {
    *reinterpret_cast<uint4*>(out_row + d) = hi;
    *reinterpret_cast<uint4*>(out_row + dim + d) = side ? hi : lo;
    *reinterpret_cast<uint4*>(out_row + 2 * dim + d) = side ? lo : hi;
}
// exponent e with max * 2^e in [512, 1024) (0 for a zero / non-finite row: NaN and inf then propagate like the reference's)
__device__ __forceinline__ int split_exponent(float row_max) {
    if (!(row_max > 0.f) || !(row_max <= 3.0e38f)) return 0;
    return max(-120, min(120, 9 - ilogbf(row_max)));
}
// one warp per row
__global__ void __launch_bounds__(256)
    split_f16_kernel(const float* __restrict__ x, const float* __restrict__ rinv, int64_t n, int dim, int64_t ld, int side,
                     __half* __restrict__ out, int64_t ld_out, float* __restrict__ scale_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n; r += nwarps) {
        const float* row = x + r * ld;
        const float ri = rinv ? rinv[r] : 1.f;
        float m = 0.f;
        for (int d = lane * 8; d < dim; d += 256) {
            float f[8];
            load8(row + d, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) m = fmaxf(m, fabsf(f[e] * ri));  // fmaxf drops NaN: a NaN row keeps e = 0 ...
        }
        m = warp_max(m);
        const int e = split_exponent(m);
        const float s = ldexpf(ri, e);                                   // ... and NaN / inf reach the operand through s
        for (int d = lane * 8; d < dim; d += 256) {
            float f[8];
            load8(row + d, f);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] *= s;
            uint4 hi, lo;
            split8_f16(f, hi, lo);
            store_split(out + r * ld_out, dim, d, side, hi, lo);
        }
        if (lane == 0) scale_out[r] = ldexpf(1.f, -e);
    }
}

// ------------------------------------------------------------------- fused small-batch path
// One launch before the similarity pass of a training step (pig/loss.py:33-39 at batch size ~1k,
// where launches dominate): per row i norms of V_i and A_i, the diagonal score, the fp16 normalised
// copies for the gradient GEMMs, and zeroing of the count / partial buffers.
// kRegs (dim <= 512): a row pair stays in registers between the norms and the normalised copies (one memory round
// trip; the step at batch ~1k is latency bound).
template <typename T, bool kRegs>
__global__ void __launch_bounds__(256)
    hinge_prep_kernel(const T* __restrict__ v, const T* __restrict__ a, int64_t n, int dim,
                      int64_t ldv, int64_t lda, float* __restrict__ rinv_v, float* __restrict__ rinv_a,
                      float* __restrict__ diag, __half* __restrict__ vh, __half* __restrict__ ah,
                      int32_t* __restrict__ row_cnt, int32_t* __restrict__ col_cnt, float* __restrict__ loss_partial,
                      int n_partials, __half* __restrict__ vx, __half* __restrict__ ax, float* __restrict__ scale_v,
                      float* __restrict__ scale_a, const float* __restrict__ rinv_v_in, const float* __restrict__ rinv_a_in) {
    // Wait first, trigger second: the workspace may still be read by the previous step's kernels, and the similarity
    // pass that follows reads v / a before ITS wait (OperandsReadyScope) -- it may only be launched once everything
    // ahead of this kernel in the stream, i.e. the producer of v and a, is complete.
    pdl_wait();
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_partials; i += (int64_t)gridDim.x * blockDim.x)
        loss_partial[i] = 0.f;
    for (int64_t r = warp; r < n; r += nwarps) {
        const T* vr = v + r * ldv;
        const T* ar = a + r * lda;
        float sv = 0.f, sa = 0.f, dot = 0.f, mv = 0.f, ma = 0.f;
        auto accumulate = [&](const float (&x)[8], const float (&y)[8]) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                sv = fmaf(x[e], x[e], sv);
                sa = fmaf(y[e], y[e], sa);
                dot = fmaf(x[e], y[e], dot);
                mv = fmaxf(mv, fabsf(x[e]));
                ma = fmaxf(ma, fabsf(y[e]));
            }
        };
        [[maybe_unused]] float xs[2][8], ys[2][8];
        if constexpr (kRegs) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (lane * 8 + 256 * i < dim) {
                    load8(vr + lane * 8 + 256 * i, xs[i]);
                    load8(ar + lane * 8 + 256 * i, ys[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (lane * 8 + 256 * i < dim) accumulate(xs[i], ys[i]);
        } else {
            for (int d = lane * 8; d < dim; d += 256) {
                float x[8], y[8];
                load8(vr + d, x);
                load8(ar + d, y);
                accumulate(x, y);
            }
        }
        sv = warp_sum(sv);
        sa = warp_sum(sa);
        dot = warp_sum(dot);
        // 1/||row|| handed over by the producer of the rows (the encoder tail, pb2_project_normalize) is used as it is
        const float rv = rinv_v_in ? rinv_v_in[r] : 1.0f / sqrtf(sv), ra = rinv_a_in ? rinv_a_in[r] : 1.0f / sqrtf(sa);
        // fp32 inputs: split-fp16 tensor-core operands of the NORMALISED rows (see split_f16_kernel)
        int ev = 0, ea = 0;
        if (vx) {
            ev = split_exponent(warp_max(mv) * rv);
            ea = split_exponent(warp_max(ma) * ra);
        }
        const float s_v = ldexpf(rv, ev), s_a = ldexpf(ra, ea);
        if (lane == 0) {
            rinv_v[r] = rv;
            rinv_a[r] = ra;
            diag[r] = __fmul_rn(__fmul_rn(dot, rv), ra);
            row_cnt[r] = 0;
            col_cnt[r] = 0;
            if (vx) {
                scale_v[r] = ldexpf(1.f, -ev);
                scale_a[r] = ldexpf(1.f, -ea);
            }
        }
        auto emit = [&](int d, float (&x)[8], float (&y)[8]) {
            __half2 hx[4], hy[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hx[e] = __floats2half2_rn(x[2 * e] * rv, x[2 * e + 1] * rv);
                hy[e] = __floats2half2_rn(y[2 * e] * ra, y[2 * e + 1] * ra);
            }
            *reinterpret_cast<uint4*>(vh + r * dim + d) = *reinterpret_cast<const uint4*>(hx);
            *reinterpret_cast<uint4*>(ah + r * dim + d) = *reinterpret_cast<const uint4*>(hy);
            if (vx) {  // split-fp16 tensor-core operands [n, 3 dim] (S = V A^T: V is side 0, A side 1)
                uint4 hi, lo;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    x[e] *= s_v;
                    y[e] *= s_a;
                }
                split8_f16(x, hi, lo);
                store_split(vx + r * 3 * dim, dim, d, 0, hi, lo);
                split8_f16(y, hi, lo);
                store_split(ax + r * 3 * dim, dim, d, 1, hi, lo);
            }
        };
        if constexpr (kRegs) {
#pragma unroll
            for (int i = 0; i < 2; ++i)
                if (lane * 8 + 256 * i < dim) emit(lane * 8 + 256 * i, xs[i], ys[i]);
        } else {
            for (int d = lane * 8; d < dim; d += 256) {
                float x[8], y[8];
                load8(vr + d, x);
                load8(ar + d, y);
                emit(d, x, y);
            }
        }
    }
}

// One launch after the gradient GEMMs: rows [0, n) -> dV, rows [n, 2n) -> dA (normalisation Jacobian +
// diagonal term, as hinge_finish_kernel), and block 0 finishes the scalar loss from the CTA partials
// and the indicator counts (NaN if any row norm is zero, like the reference's 0/0).
template <typename T>
__device__ __forceinline__ void store8(T* dst, const float (&o)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* dst, const float (&o)[8]) {
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* dst, const float (&o)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(o[2 * e], o[2 * e + 1]);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}
template <>
__device__ __forceinline__ void store8<__half>(__half* dst, const float (&o)[8]) {
    __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(o[2 * e], o[2 * e + 1]);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

// kRegs (dim <= 512): the row's x, y and p stay in registers between the dot product and the result -- every load of a
// row is in flight at once, one memory round trip instead of three (the step at batch ~1k is latency bound).
// coef_dev (optional): autograd's grad_output, applied in fp32 to the finished fp32 gradient, the rounding to TOut last
// (bit for bit what hinge_finish2 -> fp32 -> pb2_scale_pair produces).  fold.loss_out != nullptr: block 0 only folds the
// scalar loss (serial fp64 work that would otherwise sit behind a block's rows, on the step's critical path at batch
// ~1k; first in the grid, so that it runs beside the first wave of a large grid) and the other blocks share the rows.
template <typename T, typename TOut, bool kRegs>
__global__ void __launch_bounds__(256)
    hinge_finish2_kernel(const float* __restrict__ p_v, const float* __restrict__ p_a,
                         const T* __restrict__ v, const T* __restrict__ a, int64_t n, int dim,
                         int64_t ldv, int64_t lda, const float* __restrict__ rinv_v, const float* __restrict__ rinv_a,
                         const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ col_cnt, float coef,
                         const float* __restrict__ coef_dev, const HingeFold fold, TOut* __restrict__ d_v,
                         TOut* __restrict__ d_a) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int first = fold.loss_out ? 1 : 0;
    if (first && blockIdx.x == 0) {
        __shared__ double sh[8];
        __shared__ int sbad[8];
        hinge_loss_fold(fold, (int)threadIdx.x, sh, sbad, 1);
        return;
    }
    const float cdev = coef_dev ? coef_dev[0] : 1.f;
    const int64_t warp = ((int64_t)blockIdx.x - first) * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)(gridDim.x - first) * (blockDim.x >> 5);
    for (int64_t rr = warp; rr < 2 * n; rr += nwarps) {
        const bool is_v = rr < n;
        const int64_t r = is_v ? rr : rr - n;
        const float* pr = (is_v ? p_v : p_a) + r * dim;
        const T* xr = is_v ? v + r * ldv : a + r * lda;
        const T* yr = is_v ? a + r * lda : v + r * ldv;
        const float rx = is_v ? rinv_v[r] : rinv_a[r];
        const float ry = is_v ? rinv_a[r] : rinv_v[r];
        const float gd = -(float)(row_cnt[r] + col_cnt[r]) * ry;
        TOut* out = (is_v ? d_v : d_a) + r * dim;
        auto result = [&](float (&o)[8], const float (&x)[8], const float (&y)[8], const float (&pv)[8], float dot) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = coef * rx * (fmaf(gd, y[e], pv[e]) - x[e] * rx * dot);
            if (coef_dev) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __fmul_rn(o[e], cdev);
            }
        };
        if constexpr (kRegs) {
            float x[2][8], y[2][8], pv[2][8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int d = lane * 8 + 256 * i;
                if (d < dim) {
                    load8(xr + d, x[i]);
                    load8(yr + d, y[i]);
                    load8(pr + d, pv[i]);
                }
            }
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (lane * 8 + 256 * i < dim) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dot = fmaf(fmaf(gd, y[i][e], pv[i][e]), x[i][e] * rx, dot);
                }
            }
            dot = warp_sum(dot);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int d = lane * 8 + 256 * i;
                if (d < dim) {
                    float o[8];
                    result(o, x[i], y[i], pv[i], dot);
                    store8<TOut>(out + d, o);
                }
            }
            continue;
        }
        float dot = 0.f;
        for (int d = lane * 8; d < dim; d += 256) {
            float x[8], y[8], pv[8];
            load8(xr + d, x);
            load8(yr + d, y);
            load8(pr + d, pv);
#pragma unroll
            for (int e = 0; e < 8; ++e) dot = fmaf(fmaf(gd, y[e], pv[e]), x[e] * rx, dot);
        }
        dot = warp_sum(dot);
        for (int d = lane * 8; d < dim; d += 256) {
            float x[8], y[8], pv[8], o[8];
            load8(xr + d, x);
            load8(yr + d, y);
            load8(pr + d, pv);
            result(o, x, y, pv, dot);
            store8<TOut>(out + d, o);
        }
    }
}

// the fold as a launch of its own (pb2_hinge_forward when the gradient products do not fit one grid)
__global__ void __launch_bounds__(256) hinge_fold_kernel(const HingeFold fold) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ double sh[8];
    __shared__ int sbad[8];
    hinge_loss_fold(fold, (int)threadIdx.x, sh, sbad, 1);
}

// ------------------------------------------------------------------------------ hinge finish
// One warp per row; kRegs (dim <= 512): the row (p, x, y) is held in registers between the dot product and the
// result (one pass over HBM / L2 instead of two), else re-read.
template <typename T, bool kRegs>
__global__ void __launch_bounds__(256)
    hinge_finish_kernel(const float* __restrict__ p, int64_t ld_p, const T* __restrict__ x,
                        const T* __restrict__ y, const float* __restrict__ rinv_x,
                        const float* __restrict__ rinv_y,
                        const int32_t* __restrict__ row_cnt, const int32_t* __restrict__ col_cnt, int64_t rows,
                        int dim, int64_t ldx, int64_t ldy, float coef_host, const float* __restrict__ coef_dev,
                        float* __restrict__ grad, int64_t ld_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const float coef = coef_host * (coef_dev ? coef_dev[0] : 1.f);
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float rx = rinv_x[r];
        const float gd = -(float)(row_cnt[r] + col_cnt[r]) * rinv_y[r];
        const float* pr = p + r * ld_p;
        const T* xr = x + r * ldx;
        const T* yr = y + r * ldy;
        if constexpr (kRegs) {
            float a[2][8], b[2][8], pv[2][8];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int d = lane * 8 + 256 * i;
                if (d < dim) {
                    load8(xr + d, a[i]);
                    load8(yr + d, b[i]);
                    load8(pr + d, pv[i]);
                }
            }
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (lane * 8 + 256 * i < dim) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) dot = fmaf(fmaf(gd, b[i][e], pv[i][e]), a[i][e] * rx, dot);
                }
            }
            dot = warp_sum(dot);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int d = lane * 8 + 256 * i;
                if (d < dim) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float g = fmaf(gd, b[i][e], pv[i][e]);
                        o[e] = coef * rx * (g - a[i][e] * rx * dot);
                    }
                    float* gr = grad + r * ld_grad + d;
                    *reinterpret_cast<float4*>(gr) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(gr + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
            continue;
        }
        float dot = 0.f;
        for (int d = lane * 8; d < dim; d += 256) {
            float a[8], b[8];
            load8(xr + d, a);
            load8(yr + d, b);
            const float4 p0 = *reinterpret_cast<const float4*>(pr + d);
            const float4 p1 = *reinterpret_cast<const float4*>(pr + d + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) dot = fmaf(fmaf(gd, b[e], pv[e]), a[e] * rx, dot);
        }
        dot = warp_sum(dot);
        for (int d = lane * 8; d < dim; d += 256) {
            float a[8], b[8];
            load8(xr + d, a);
            load8(yr + d, b);
            const float4 p0 = *reinterpret_cast<const float4*>(pr + d);
            const float4 p1 = *reinterpret_cast<const float4*>(pr + d + 4);
            const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float g = fmaf(gd, b[e], pv[e]);
                o[e] = coef * rx * (g - a[e] * rx * dot);
            }
            float* gr = grad + r * ld_grad + d;
            *reinterpret_cast<float4*>(gr) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(gr + 4) = make_float4(o[4], o[5], o[6], o[7]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
    milnce_finish_kernel(const float* __restrict__ p, int64_t ld_p, const T* __restrict__ y, int64_t rows,
                         int dim, int64_t ldy, float coef_host, const float* __restrict__ coef_dev,
                         float* __restrict__ grad, int64_t ld_grad) {
    const float coef = coef_host * (coef_dev ? coef_dev[0] : 1.f);
    const int vec_per_row = dim / 8;
    const int64_t total = rows * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row;
        const int d = (int)(i % vec_per_row) * 8;
        float b[8];
        load8(y + r * ldy + d, b);
        const float4 p0 = *reinterpret_cast<const float4*>(p + r * ld_p + d);
        const float4 p1 = *reinterpret_cast<const float4*>(p + r * ld_p + d + 4);
        const float s = 1.0f / 8192.0f;
        float* gr = grad + r * ld_grad + d;
        *reinterpret_cast<float4*>(gr) = make_float4(coef * (p0.x * s - b[0]), coef * (p0.y * s - b[1]),
                                                     coef * (p0.z * s - b[2]), coef * (p0.w * s - b[3]));
        *reinterpret_cast<float4*>(gr + 4) = make_float4(coef * (p1.x * s - b[4]), coef * (p1.y * s - b[5]),
                                                         coef * (p1.z * s - b[6]), coef * (p1.w * s - b[7]));
    }
}

// MIL-NCE with K candidates per clip (pig/loss.py:19-25 views x as [N, N, K]): the positive term of row r
// is sum_k w[r * group + k] * y[(r * group + k) / y_div] with w = softmax_k of the K paired logits.
//   video side: group = K, y_div = 1 (the clip's K audio rows);  audio side: group = 1, y_div = K (its video).
template <typename T>
__global__ void __launch_bounds__(256)
    milnce_finish_k_kernel(const float* __restrict__ p, int64_t ld_p, const T* __restrict__ y,
                           const float* __restrict__ w, int64_t rows, int group, int y_div, int dim, int64_t ldy,
                           float coef_host, const float* __restrict__ coef_dev, float* __restrict__ grad,
                           int64_t ld_grad) {
    const float coef = coef_host * (coef_dev ? coef_dev[0] : 1.f);
    const int vec_per_row = dim / 8;
    const int64_t total = rows * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row;
        const int d = (int)(i % vec_per_row) * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = 0; k < group; ++k) {
            const int64_t c = r * group + k;
            float b[8];
            load8(y + (c / y_div) * ldy + d, b);
            const float wk = w[c];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(wk, b[j], acc[j]);
        }
        const float4 p0 = *reinterpret_cast<const float4*>(p + r * ld_p + d);
        const float4 p1 = *reinterpret_cast<const float4*>(p + r * ld_p + d + 4);
        const float s = 1.0f / 8192.0f;
        float* gr = grad + r * ld_grad + d;
        *reinterpret_cast<float4*>(gr) = make_float4(coef * (p0.x * s - acc[0]), coef * (p0.y * s - acc[1]),
                                                     coef * (p0.z * s - acc[2]), coef * (p0.w * s - acc[3]));
        *reinterpret_cast<float4*>(gr + 4) = make_float4(coef * (p1.x * s - acc[4]), coef * (p1.y * s - acc[5]),
                                                         coef * (p1.z * s - acc[6]), coef * (p1.w * s - acc[7]));
    }
}

// ------------------------------------------------------------------- resampled recall (f1)
// pig/metrics.py:54-77 draws n_samples subsets of `size` clips and ranks each 100 x 100 sub-matrix with
// its own GEMM + argsort loop.  Here the G x G score matrix is computed once (pb2_sim_matrix) and every
// (sample, query) pair counts, inside its subset, the candidates strictly closer than its positive:
// rank[s, j] = #{ c != j : fl32(1 - S[ix_j, ix_c]) < fl32(1 - S[ix_j, ix_j]) }.  One warp per (sample, query).
__global__ void __launch_bounds__(256)
    subset_rank_kernel(const float* __restrict__ S, int64_t ld, const int64_t* __restrict__ idx, int n_samples,
                       int size, int32_t* __restrict__ rank) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t total = (int64_t)n_samples * size;
    for (int64_t w = warp; w < total; w += nwarps) {
        const int64_t s = w / size;
        const int j = (int)(w % size);
        const int64_t* ix = idx + s * size;
        const float* row = S + ix[j] * ld;
        const float dpos = __fsub_rn(1.0f, row[ix[j]]);
        int cnt = 0;
        for (int c = lane; c < size; c += 32)
            cnt += (c != j && __fsub_rn(1.0f, row[ix[c]]) < dpos) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) rank[w] = cnt;
    }
}

// ---------------------------------------------------------------------- contrastive(M) on a matrix
// pass 1: one block per row: loss partial, indicator counts, off-diagonal gradient entries.
__global__ void __launch_bounds__(256)
    contrastive_rows_kernel(const float* __restrict__ m, int64_t n, int64_t ld, float margin,
                            float* __restrict__ loss_partial, int32_t* __restrict__ row_cnt,
                            int32_t* __restrict__ col_cnt, float* __restrict__ grad, int64_t ld_grad, float coef) {
    __shared__ float shf[8];
    __shared__ int shi[8];
    float total = 0.f;
    for (int64_t i = blockIdx.x; i < n; i += gridDim.x) {
        const float di = m[i * ld + i];
        float l = 0.f;
        int rc = 0;
        for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
            if (j == i) continue;
            const float v = m[i * ld + j];
            const float zc = margin + v - m[j * ld + j];
            const float zr = margin + v - di;
            const bool ic = zc >= 0.f, ir = zr >= 0.f;
            l += (ic ? zc : 0.f) + (ir ? zr : 0.f);
            rc += ir;
            if (ic) atomicAdd(col_cnt + j, 1);
            if (grad) grad[i * ld_grad + j] = coef * ((ic ? 1.f : 0.f) + (ir ? 1.f : 0.f));
        }
        l = warp_sum(l);
        for (int o = 16; o > 0; o >>= 1) rc += __shfl_xor_sync(0xffffffffu, rc, o);
        if ((threadIdx.x & 31) == 0) {
            shf[threadIdx.x >> 5] = l;
            shi[threadIdx.x >> 5] = rc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            int c = 0;
            for (int w = 0; w < 8; ++w) {
                s += shf[w];
                c += shi[w];
            }
            total += s;
            row_cnt[i] = c;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_partial[blockIdx.x] = total;
}
__global__ void contrastive_diag_kernel(int64_t n, const int32_t* __restrict__ row_cnt,
                                        const int32_t* __restrict__ col_cnt, float* __restrict__ grad,
                                        int64_t ld_grad, float coef) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grad[i * ld_grad + i] = -coef * (float)(row_cnt[i] + col_cnt[i]);
}
__global__ void scale_by_dev_kernel(float* __restrict__ g, int64_t n, int64_t ld, const float* __restrict__ s) {
    const float f = s[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * n; i += (int64_t)gridDim.x * blockDim.x)
        g[(i / n) * ld + (i % n)] *= f;
}

static int grid_for_warps(int64_t rows) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((rows + 7) / 8, (int64_t)sm_count() * 8));
}

}  // namespace pb2

using namespace pb2;

static bool vec_ok(const void* p, int64_t ld_elems, int elem_bytes) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld_elems * elem_bytes) % 16 == 0;
}
static int elem_bytes(int dtype) { return dtype == PB2_F32 ? 4 : ((dtype == PB2_BF16 || dtype == PB2_F16) ? 2 : 0); }
// run `...` with T bound to the element type of `dtype` (bf16 / fp16 / fp32 rows)
#define PB2_ROWS_DISPATCH(dtype, ...)                       \
    do {                                                    \
        if ((dtype) == PB2_BF16) {                          \
            using T = __nv_bfloat16;                        \
            __VA_ARGS__;                                    \
        } else if ((dtype) == PB2_F16) {                    \
            using T = __half;                               \
            __VA_ARGS__;                                    \
        } else {                                            \
            using T = float;                                \
            __VA_ARGS__;                                    \
        }                                                   \
    } while (0)

extern "C" int pb2_row_norms(const void* x, int dtype, int64_t n, int dim, int64_t ld, float* rinv, float* norm,
                             void* stream) {
    if (n <= 0) return PB2_OK;
    const int es = elem_bytes(dtype);
    if (!x || !es || dim <= 0 || dim % 8 != 0 || !vec_ok(x, ld, es))
        return set_error(PB2_ERR_ARG, "row_norms: need bf16 / fp16 / fp32 rows, dim %% 8 == 0, 16-byte aligned");
    PB2_ROWS_DISPATCH(dtype, row_norms_kernel<T><<<grid_for_warps(n), 256, 0, (cudaStream_t)stream>>>((const T*)x, n, dim, ld,
                                                                                                    rinv, norm));
    return check_launch("row_norms");
}

extern "C" int pb2_split_f16(const float* x, const float* rinv, int64_t n, int dim, int64_t ld, int side, void* out,
                             int64_t ld_out, float* scale_out, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!x || !out || !scale_out || dim <= 0 || dim % 8 != 0 || !vec_ok(x, ld, 4) || !vec_ok(out, ld_out, 2) ||
        ld_out < 3 * (int64_t)dim || (side != 0 && side != 1))
        return set_error(PB2_ERR_ARG, "split_f16: need fp32 rows, dim %% 8 == 0, 16-byte aligned, ld_out >= 3 dim, side 0 / 1");
    split_f16_kernel<<<grid_for_warps(n), 256, 0, (cudaStream_t)stream>>>(x, rinv, n, dim, ld, side, (__half*)out, ld_out,
                                                                         scale_out);
    return check_launch("split_f16");
}

extern "C" int pb2_rows_scale_f16(const void* x, int dtype, const float* rinv, int64_t n, int dim, int64_t ld, void* out,
                                  int64_t ld_out, void* stream) {
    if (n <= 0) return PB2_OK;
    const int es = elem_bytes(dtype);
    if (!x || !out || !es || dim <= 0 || dim % 8 != 0 || !vec_ok(x, ld, es) || !vec_ok(out, ld_out, 2))
        return set_error(PB2_ERR_ARG, "rows_scale_f16: need bf16 / fp16 / fp32 rows, dim %% 8 == 0, 16-byte aligned");
    const int64_t total = n * (dim / 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8));
    PB2_ROWS_DISPATCH(dtype, rows_scale_f16_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, rinv, n, dim, ld,
                                                                                            (__half*)out, ld_out));
    return check_launch("rows_scale_f16");
}

extern "C" int pb2_rows_quant_i8(const void* x, int dtype, const float* rinv, int64_t n, int dim, int64_t ld, void* out,
                                 int64_t ld_out, void* stream) {
    if (n <= 0) return PB2_OK;
    const int es = elem_bytes(dtype);
    if (!x || !out || !es || dim <= 0 || dim % 8 != 0 || !vec_ok(x, ld, es) || !vec_ok(out, ld_out, 1) || ld_out < 2 * (int64_t)dim)
        return set_error(PB2_ERR_ARG, "rows_quant_i8: need bf16 / fp16 / fp32 rows, dim %% 8 == 0, 16-byte aligned, ld_out >= 2 dim");
    const int64_t total = n * (dim / 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8));
    PB2_ROWS_DISPATCH(dtype, rows_quant_i8_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, rinv, n, dim, ld,
                                                                                           (uint8_t*)out, ld_out));
    return check_launch("rows_quant_i8");
}

extern "C" int pb2_scale_pair(const float* x0, const float* x1, int64_t n_elems, int out_dtype, const float* coef,
                              void* y0, void* y1, void* stream) {
    if (n_elems <= 0) return PB2_OK;
    const int es = out_dtype == PB2_F32 ? 4 : 2;
    auto ok = [](const void* p) { return p && (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!ok(x0) || !ok(x1) || !ok(y0) || !ok(y1) || !coef || (n_elems * es) % 16 != 0 ||
        (out_dtype != PB2_F32 && out_dtype != PB2_BF16 && out_dtype != PB2_F16))
        return set_error(PB2_ERR_ARG, "scale_pair: need 16-byte aligned fp32 inputs and bf16 / fp16 / fp32 outputs of whole 16-byte vectors");
    const int64_t n_vec = n_elems * es / 16;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((2 * n_vec + 255) / 256, (int64_t)sm_count() * 8));
    cudaStream_t st = (cudaStream_t)stream;
    // the backward of the launch-bound training step: its launch overlaps the tail of whatever precedes it in the stream
    // (autograd's ones_like fill, or hinge_finish2 itself)
    PdlScope pdl;
    cudaError_t e;
    if (out_dtype == PB2_F32)
        e = launch_ex(scale_pair_kernel<float>, (unsigned)grid, 256u, (size_t)0, st, 1, x0, x1, n_vec, coef, (float*)y0, (float*)y1);
    else if (out_dtype == PB2_BF16)
        e = launch_ex(scale_pair_kernel<__nv_bfloat16>, (unsigned)grid, 256u, (size_t)0, st, 1, x0, x1, n_vec, coef,
                      (__nv_bfloat16*)y0, (__nv_bfloat16*)y1);
    else
        e = launch_ex(scale_pair_kernel<__half>, (unsigned)grid, 256u, (size_t)0, st, 1, x0, x1, n_vec, coef, (__half*)y0,
                      (__half*)y1);
    const int rc = check_cuda(e, "scale_pair");
    if (rc) return rc;
    return check_launch("scale_pair");
}

extern "C" int pb2_pair_dot(const void* x, const void* y, int dtype, const int64_t* ix, const int64_t* iy,
                            const float* rinv_x, const float* rinv_y, int64_t n, int dim, int64_t ldx, int64_t ldy,
                            float* out, float* dist_out, float* thr_out, void* stream) {
    if (n <= 0) return PB2_OK;
    const int es = elem_bytes(dtype);
    if (!x || !y || !es || dim <= 0 || dim % 8 != 0 || !vec_ok(x, ldx, es) || !vec_ok(y, ldy, es))
        return set_error(PB2_ERR_ARG, "pair_dot: need bf16 / fp16 / fp32 rows, dim %% 8 == 0, 16-byte aligned");
    PB2_ROWS_DISPATCH(dtype, pair_dot_kernel<T><<<grid_for_warps(n), 256, 0, (cudaStream_t)stream>>>(
                                 (const T*)x, (const T*)y, ix, iy, rinv_x, rinv_y, n, dim, ldx, ldy, out, dist_out, thr_out));
    return check_launch("pair_dot");
}

extern "C" int pb2_subset_rank(const float* scores, int64_t ld, const int64_t* idx, int n_samples, int size,
                               int32_t* rank, void* stream) {
    if (n_samples <= 0 || size <= 0) return PB2_OK;
    if (!scores || !idx || !rank) return set_error(PB2_ERR_ARG, "subset_rank: null");
    subset_rank_kernel<<<grid_for_warps((int64_t)n_samples * size), 256, 0, (cudaStream_t)stream>>>(scores, ld, idx, n_samples,
                                                                                                 size, rank);
    return check_launch("subset_rank");
}

extern "C" int pb2_lse_merge(const float* part_max, const float* part_sum, int n_parts, int64_t rows, float* lse,
                             int accumulate, void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!part_max || !part_sum || !lse || n_parts <= 0) return set_error(PB2_ERR_ARG, "lse_merge: bad arguments");
    lse_merge_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(part_max, part_sum, n_parts,
                                                                                      rows, lse, accumulate);
    return check_launch("lse_merge");
}

extern "C" int pb2_lse_merge_const(const float* part_sum, int n_parts, int64_t n, float bound, float* lse,
                                   int accumulate, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!part_sum || !lse || n_parts <= 0) return set_error(PB2_ERR_ARG, "lse_merge_const: bad arguments");
    lse_merge_const_kernel<<<(unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream>>>(
        part_sum, n_parts, n, bound * 1.4426950408889634f, lse, accumulate);
    return check_launch("lse_merge_const");
}

extern "C" int pb2_lse_combine(const float* parts, int n_parts, int64_t n, float* out, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!parts || !out || n_parts <= 0) return set_error(PB2_ERR_ARG, "lse_combine: bad arguments");
    lse_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(parts, n_parts, n, out);
    return check_launch("lse_combine");
}

extern "C" int pb2_sum_partials(const float* partials, int n, float alpha, float* out, void* stream) {
    if (!partials || !out || n < 0) return set_error(PB2_ERR_ARG, "sum_partials: bad arguments");
    sum_partials_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n, alpha, out);
    return check_launch("sum_partials");
}

extern "C" int pb2_hinge_loss_terms(const float* partials, int n_partials, const float* diag, const int32_t* cnt,
                                    int64_t n, float margin, float alpha, float* out, int accumulate, void* stream) {
    if (!out || n_partials < 0 || n < 0 || ((diag == nullptr) != (cnt == nullptr)))
        return set_error(PB2_ERR_ARG, "hinge_loss_terms: bad arguments");
    hinge_loss_terms_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partials, n_partials, diag, cnt, n, margin, alpha, out,
                                                                accumulate);
    return check_launch("hinge_loss_terms");
}

extern "C" int pb2_milnce_loss(const float* lse_row, const float* lse_col, const float* diag, int64_t n, float* den,
                               float* out, void* stream) {
    if (n <= 0 || !lse_row || !lse_col || !diag || !den || !out)
        return set_error(PB2_ERR_ARG, "milnce_loss: bad arguments");
    milnce_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lse_row, lse_col, diag, n, den, out);
    return check_launch("milnce_loss");
}

extern "C" int pb2_hinge_finish(const float* p, int64_t ld_p, const void* x, const void* y, int dtype, const float* rinv_x,
                                const float* rinv_y, const int32_t* row_cnt,
                                const int32_t* col_cnt, int64_t rows, int dim, int64_t ldx, int64_t ldy,
                                float coef_host, const float* coef_dev, float* grad_x, int64_t ld_grad,
                                void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!p || !x || !y || !rinv_x || !rinv_y || !row_cnt || !col_cnt || !grad_x)
        return set_error(PB2_ERR_ARG, "hinge_finish: null");
    const int es = elem_bytes(dtype);
    if (!es || dim % 8 != 0 || !vec_ok(x, ldx, es) || !vec_ok(y, ldy, es) || !vec_ok(p, ld_p, 4) || !vec_ok(grad_x, ld_grad, 4))
        return set_error(PB2_ERR_ARG, "hinge_finish: dtype / alignment");
    if (dim <= 512)
        PB2_ROWS_DISPATCH(dtype, hinge_finish_kernel<T, true><<<grid_for_warps(rows), 256, 0, (cudaStream_t)stream>>>(
                                 p, ld_p, (const T*)x, (const T*)y, rinv_x, rinv_y, row_cnt, col_cnt, rows, dim, ldx, ldy,
                                 coef_host, coef_dev, grad_x, ld_grad));
    else
        PB2_ROWS_DISPATCH(dtype, hinge_finish_kernel<T, false><<<grid_for_warps(rows), 256, 0, (cudaStream_t)stream>>>(
                                 p, ld_p, (const T*)x, (const T*)y, rinv_x, rinv_y, row_cnt, col_cnt, rows, dim, ldx, ldy,
                                 coef_host, coef_dev, grad_x, ld_grad));
    return check_launch("hinge_finish");
}

extern "C" int pb2_hinge_prep(const void* v, const void* a, int dtype, int64_t n, int dim, int64_t ldv, int64_t lda,
                              float* rinv_v, float* rinv_a, float* diag, void* vh, void* ah, int32_t* row_cnt,
                              int32_t* col_cnt, float* loss_partial, int n_partials, void* v_split, void* a_split,
                              float* scale_v, float* scale_a, const float* rinv_v_in, const float* rinv_a_in, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!v || !a || !rinv_v || !rinv_a || !diag || !vh || !ah || !row_cnt || !col_cnt || !loss_partial)
        return set_error(PB2_ERR_ARG, "hinge_prep: null");
    const int es = elem_bytes(dtype);
    if (!es || dim % 8 != 0 || !vec_ok(v, ldv, es) || !vec_ok(a, lda, es) || !vec_ok(vh, dim, 2) || !vec_ok(ah, dim, 2))
        return set_error(PB2_ERR_ARG, "hinge_prep: dtype / alignment");
    if ((v_split == nullptr) != (a_split == nullptr) ||
        (v_split && (dtype != PB2_F32 || !scale_v || !scale_a || !vec_ok(v_split, 3 * (int64_t)dim, 2) ||
                     !vec_ok(a_split, 3 * (int64_t)dim, 2))))
        return set_error(PB2_ERR_ARG, "hinge_prep: the split-fp16 outputs and their scales go together, for fp32 rows only");
    cudaError_t e;
#define PB2_PREP(REGS)                                                                                                     \
    PB2_ROWS_DISPATCH(dtype, e = launch_ex(hinge_prep_kernel<T, REGS>, (unsigned)grid_for_warps(n), 256u, (size_t)0,       \
                                           (cudaStream_t)stream, 1, (const T*)v, (const T*)a, n, dim, ldv, lda, rinv_v, rinv_a, \
                                           diag, (__half*)vh, (__half*)ah, row_cnt, col_cnt, loss_partial, n_partials,     \
                                           (__half*)v_split, (__half*)a_split, scale_v, scale_a, rinv_v_in, rinv_a_in))
    if (dim <= 512) PB2_PREP(true);
    else PB2_PREP(false);
#undef PB2_PREP
    int rc = check_cuda(e, "hinge_prep");
    if (rc) return rc;
    return check_launch("hinge_prep");
}

int pb2::hinge_finish2_ex(const float* p_v, const float* p_a, const void* v, const void* a, int dtype, int64_t n, int dim,
                          int64_t ldv, int64_t lda, const float* rinv_v, const float* rinv_a, const int32_t* row_cnt,
                          const int32_t* col_cnt, float coef, const float* coef_dev, const HingeFold& fold, void* d_v, void* d_a,
                          int out_dtype, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!p_v || !p_a || !v || !a || !rinv_v || !rinv_a || !row_cnt || !col_cnt || !d_v || !d_a)
        return set_error(PB2_ERR_ARG, "hinge_finish2: null");
    const int ob = out_dtype == PB2_F32 ? 4 : 2;
    const int es = elem_bytes(dtype);
    if (!es || dim % 8 != 0 || !vec_ok(v, ldv, es) || !vec_ok(a, lda, es) || !vec_ok(p_v, dim, 4) || !vec_ok(p_a, dim, 4) ||
        !vec_ok(d_v, dim, ob) || !vec_ok(d_a, dim, ob))
        return set_error(PB2_ERR_ARG, "hinge_finish2: dtype / alignment");
    const unsigned grid = (unsigned)grid_for_warps(2 * n) + (fold.loss_out ? 1u : 0u);
    cudaError_t e;
#define PB2_FIN2_(TO, REGS)                                                                                                   \
    PB2_ROWS_DISPATCH(dtype, e = launch_ex(hinge_finish2_kernel<T, TO, REGS>, grid, 256u, (size_t)0, (cudaStream_t)stream, 1, \
                                           p_v, p_a, (const T*)v, (const T*)a, n, dim, ldv, lda, rinv_v, rinv_a, row_cnt,     \
                                           col_cnt, coef, coef_dev, fold, (TO*)d_v, (TO*)d_a))
#define PB2_FIN2(TO)                         \
    do {                                     \
        if (dim <= 512) PB2_FIN2_(TO, true); \
        else PB2_FIN2_(TO, false);           \
    } while (0)
    if (out_dtype == PB2_F32) PB2_FIN2(float);
    else if (out_dtype == PB2_BF16) PB2_FIN2(__nv_bfloat16);
    else if (out_dtype == PB2_F16) PB2_FIN2(__half);
    else return set_error(PB2_ERR_ARG, "hinge_finish2: unknown output dtype");
#undef PB2_FIN2
#undef PB2_FIN2_
    int rc = check_cuda(e, "hinge_finish2");
    if (rc) return rc;
    return check_launch("hinge_finish2");
}

int pb2::hinge_fold(const HingeFold& fold, void* stream) {
    if (!fold.loss_out) return PB2_OK;
    const int rc = check_cuda(launch_ex(hinge_fold_kernel, 1u, 256u, (size_t)0, (cudaStream_t)stream, 1, fold), "hinge_fold");
    if (rc) return rc;
    return check_launch("hinge_fold");
}

extern "C" int pb2_hinge_finish2(const float* p_v, const float* p_a, const void* v, const void* a, int dtype, int64_t n, int dim,
                                 int64_t ldv, int64_t lda, const float* rinv_v, const float* rinv_a, const float* diag,
                                 const int32_t* row_cnt, const int32_t* col_cnt, const float* loss_partial,
                                 int n_partials, float margin, float coef, float* loss_out, void* d_v, void* d_a,
                                 int out_dtype, void* stream) {
    if (n <= 0) return PB2_OK;
    if (!diag || !loss_partial || !loss_out) return set_error(PB2_ERR_ARG, "hinge_finish2: null");
    HingeFold fold;
    fold.loss_partial = loss_partial;
    fold.n_partials = n_partials;
    fold.diag = diag;
    fold.row_cnt = row_cnt;
    fold.col_cnt = col_cnt;
    fold.rinv_v = rinv_v;
    fold.rinv_a = rinv_a;
    fold.n = n;
    fold.margin = margin;
    fold.coef = coef;
    fold.loss_out = loss_out;
    return hinge_finish2_ex(p_v, p_a, v, a, dtype, n, dim, ldv, lda, rinv_v, rinv_a, row_cnt, col_cnt, coef, nullptr, fold, d_v, d_a,
                            out_dtype, stream);
}

extern "C" int pb2_milnce_finish(const float* p, int64_t ld_p, const void* y, int dtype, int64_t rows, int dim, int64_t ldy,
                                 float coef_host, const float* coef_dev, float* grad_x, int64_t ld_grad,
                                 void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!p || !y || !grad_x) return set_error(PB2_ERR_ARG, "milnce_finish: null");
    const int es = elem_bytes(dtype);
    if (!es || dim % 8 != 0 || !vec_ok(y, ldy, es) || !vec_ok(p, ld_p, 4) || !vec_ok(grad_x, ld_grad, 4))
        return set_error(PB2_ERR_ARG, "milnce_finish: dtype / alignment");
    const int64_t total = rows * (dim / 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8));
    PB2_ROWS_DISPATCH(dtype, milnce_finish_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(p, ld_p, (const T*)y, rows, dim, ldy,
                                                                                           coef_host, coef_dev, grad_x, ld_grad));
    return check_launch("milnce_finish");
}

extern "C" int pb2_milnce_finish_k(const float* p, int64_t ld_p, const void* y, int dtype, const float* w, int64_t rows, int group,
                                   int y_div, int dim, int64_t ldy, float coef_host, const float* coef_dev,
                                   float* grad_x, int64_t ld_grad, void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!p || !y || !w || !grad_x) return set_error(PB2_ERR_ARG, "milnce_finish_k: null");
    if (group < 1 || y_div < 1) return set_error(PB2_ERR_ARG, "milnce_finish_k: group and y_div must be >= 1");
    const int es = elem_bytes(dtype);
    if (!es || dim % 8 != 0 || !vec_ok(y, ldy, es) || !vec_ok(p, ld_p, 4) || !vec_ok(grad_x, ld_grad, 4))
        return set_error(PB2_ERR_ARG, "milnce_finish_k: dtype / alignment");
    const int64_t total = rows * (dim / 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8));
    PB2_ROWS_DISPATCH(dtype, milnce_finish_k_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
                                 p, ld_p, (const T*)y, w, rows, group, y_div, dim, ldy, coef_host, coef_dev, grad_x, ld_grad));
    return check_launch("milnce_finish_k");
}

extern "C" int pb2_contrastive_matrix(const float* m, int64_t n, int64_t ld, float margin, float* loss_partial,
                                      int n_partials, float* grad_m, int64_t ld_grad, float coef_host,
                                      const float* coef_dev, void* stream) {
    // workspace for the indicator counts lives at the tail of loss_partial: [n_partials | 2n int32]
    if (n <= 0 || !m || !loss_partial) return set_error(PB2_ERR_ARG, "contrastive_matrix: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::min<int64_t>(n, n_partials);
    int32_t* cnt = reinterpret_cast<int32_t*>(loss_partial + n_partials);
    int rc = check_cuda(cudaMemsetAsync(loss_partial, 0, sizeof(float) * n_partials + sizeof(int32_t) * 2 * n, st),
                        "contrastive_matrix memset");
    if (rc) return rc;
    contrastive_rows_kernel<<<grid, 256, 0, st>>>(m, n, ld, margin, loss_partial, cnt, cnt + n, grad_m, ld_grad,
                                                  coef_host);
    rc = check_launch("contrastive_rows");
    if (rc || !grad_m) return rc;
    contrastive_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, cnt, cnt + n, grad_m, ld_grad, coef_host);
    rc = check_launch("contrastive_diag");
    if (rc || !coef_dev) return rc;
    scale_by_dev_kernel<<<std::min<int64_t>((n * n + 255) / 256, (int64_t)sm_count() * 8), 256, 0, st>>>(
        grad_m, n, ld_grad, coef_dev);
    return check_launch("contrastive_scale");
}
