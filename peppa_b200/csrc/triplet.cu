// Kernel (c): anchor / positive / negative cosine-gap scoring.
//
// Replaces pig/metrics.py:45-52 (triplet_accuracy): two F.cosine_similarity passes, a
// subtraction and a sign -- here one pass over the three [T, D] streams.  Optional int64
// row indices fuse the gathers of pig/triplet.py:71-73,89-91 (audio[pos], video[pos],
// video[neg]) so no gathered copies are materialised.
//
// HBM-bound: algorithmic bytes per triplet = 3 * D * sizeof(elem) + 4 (3076 B at D=512
// bf16).  One warp owns a triplet; every lane issues all of its 16-byte loads for the
// triplet (and for the next one) before consuming any, so each SM keeps hundreds of KB in
// flight; five dot products are accumulated in fp32 and warp-reduced with shuffles.
#include "common.cuh"
#include "peppa_b200.h"
#include "host_util.h"

namespace pb2 {

template <typename T>
struct Elem;
template <>
struct Elem<__nv_bfloat16> {
    static constexpr int kVec = 8;
    __device__ static void unpack(const uint4& u, float (&f)[8]) {
        // bf16 -> fp32 is a 16-bit shift
        f[0] = __uint_as_float(u.x << 16);
        f[1] = __uint_as_float(u.x & 0xffff0000u);
        f[2] = __uint_as_float(u.y << 16);
        f[3] = __uint_as_float(u.y & 0xffff0000u);
        f[4] = __uint_as_float(u.z << 16);
        f[5] = __uint_as_float(u.z & 0xffff0000u);
        f[6] = __uint_as_float(u.w << 16);
        f[7] = __uint_as_float(u.w & 0xffff0000u);
    }
    __device__ static float load1(const void* p, int64_t i) {
        return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    }
};
template <>
struct Elem<__half> {
    static constexpr int kVec = 8;
    __device__ static void unpack(const uint4& u, float (&f)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 t = __half22float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    __device__ static float load1(const void* p, int64_t i) {
        return __half2float(reinterpret_cast<const __half*>(p)[i]);
    }
};
template <>
struct Elem<float> {
    static constexpr int kVec = 4;
    __device__ static void unpack(const uint4& u, float (&f)[4]) {
        f[0] = __uint_as_float(u.x);
        f[1] = __uint_as_float(u.y);
        f[2] = __uint_as_float(u.z);
        f[3] = __uint_as_float(u.w);
    }
    __device__ static float load1(const void* p, int64_t i) { return reinterpret_cast<const float*>(p)[i]; }
};

struct Dots {
    float ap, an, aa, pp, nn;
};

// F.cosine_similarity semantics (torch 2.x): each norm is clamped to eps = 1e-8 from below.
__device__ __forceinline__ float finish_triplet(const Dots& d, int discrete) {
    const float eps = 1e-8f;
    const float na = fmaxf(sqrtf(d.aa), eps), np_ = fmaxf(sqrtf(d.pp), eps), nn_ = fmaxf(sqrtf(d.nn), eps);
    const float gap = d.ap / (na * np_) - d.an / (na * nn_);
    if (!discrete) return gap;
    // (sign(gap) + 1) / 2 -> {0, 0.5, 1}; NaN propagates like torch.sign.
    return gap > 0.f ? 1.f : (gap < 0.f ? 0.f : (gap == 0.f ? 0.5f : gap));
}

__device__ __forceinline__ Dots warp_reduce(Dots d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d.ap += __shfl_xor_sync(0xffffffffu, d.ap, o);
        d.an += __shfl_xor_sync(0xffffffffu, d.an, o);
        d.aa += __shfl_xor_sync(0xffffffffu, d.aa, o);
        d.pp += __shfl_xor_sync(0xffffffffu, d.pp, o);
        d.nn += __shfl_xor_sync(0xffffffffu, d.nn, o);
    }
    return d;
}

__device__ __forceinline__ int64_t row_of(const int64_t* idx, int64_t t) { return idx ? idx[t] : t; }

// kVPL = 16-byte vectors per lane per row (D = kVPL * 32 * kVec).  Two triplets in flight per warp.
template <typename T, int kVPL>
__global__ void __launch_bounds__(256) triplet_vec_kernel(const T* __restrict__ a, const T* __restrict__ p,
                                                          const T* __restrict__ n, const int64_t* __restrict__ ia,
                                                          const int64_t* __restrict__ ip,
                                                          const int64_t* __restrict__ in_, int64_t T_, int64_t ld,
                                                          int discrete, float* __restrict__ out) {
    constexpr int kVec = Elem<T>::kVec;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    constexpr int kPair = 2;

    for (int64_t t0 = warp * kPair; t0 < T_; t0 += nwarps * kPair) {
        uint4 va[kPair][kVPL], vp[kPair][kVPL], vn[kPair][kVPL];
#pragma unroll
        for (int u = 0; u < kPair; ++u) {
            const int64_t t = (t0 + u < T_) ? t0 + u : t0;  // tail: re-read a valid row, result discarded
            const T* ra = a + row_of(ia, t) * ld;
            const T* rp = p + row_of(ip, t) * ld;
            const T* rn = n + row_of(in_, t) * ld;
#pragma unroll
            for (int v = 0; v < kVPL; ++v) {
                const int off = (v * 32 + lane) * kVec;
                va[u][v] = ld_stream_v4(ra + off);
                vp[u][v] = ld_stream_v4(rp + off);
                vn[u][v] = ld_stream_v4(rn + off);
            }
        }
#pragma unroll
        for (int u = 0; u < kPair; ++u) {
            Dots d = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int v = 0; v < kVPL; ++v) {
                float fa[kVec], fp[kVec], fn[kVec];
                Elem<T>::unpack(va[u][v], fa);
                Elem<T>::unpack(vp[u][v], fp);
                Elem<T>::unpack(vn[u][v], fn);
#pragma unroll
                for (int e = 0; e < kVec; ++e) {
                    d.ap = fmaf(fa[e], fp[e], d.ap);
                    d.an = fmaf(fa[e], fn[e], d.an);
                    d.aa = fmaf(fa[e], fa[e], d.aa);
                    d.pp = fmaf(fp[e], fp[e], d.pp);
                    d.nn = fmaf(fn[e], fn[e], d.nn);
                }
            }
            d = warp_reduce(d);
            if (lane == 0 && t0 + u < T_) out[t0 + u] = finish_triplet(d, discrete);
        }
    }
}

// Any D, any alignment: one warp per triplet, scalar loads strided by lane.
template <typename T>
__global__ void __launch_bounds__(256) triplet_generic_kernel(const T* __restrict__ a, const T* __restrict__ p,
                                                              const T* __restrict__ n, const int64_t* __restrict__ ia,
                                                              const int64_t* __restrict__ ip,
                                                              const int64_t* __restrict__ in_, int64_t T_, int D,
                                                              int64_t ld, int discrete, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = warp; t < T_; t += nwarps) {
        const int64_t ra = row_of(ia, t) * ld, rp = row_of(ip, t) * ld, rn = row_of(in_, t) * ld;
        Dots d = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k = lane; k < D; k += 32) {
            const float x = Elem<T>::load1(a, ra + k), y = Elem<T>::load1(p, rp + k), z = Elem<T>::load1(n, rn + k);
            d.ap = fmaf(x, y, d.ap);
            d.an = fmaf(x, z, d.an);
            d.aa = fmaf(x, x, d.aa);
            d.pp = fmaf(y, y, d.pp);
            d.nn = fmaf(z, z, d.nn);
        }
        d = warp_reduce(d);
        if (lane == 0) out[t] = finish_triplet(d, discrete);
    }
}

template <typename T>
static int launch_triplet(const void* a, const void* p, const void* n, const int64_t* ia, const int64_t* ip,
                          const int64_t* in_, int64_t T_, int D, int64_t ld, int discrete, float* out,
                          cudaStream_t st) {
    constexpr int kVec = Elem<T>::kVec;
    const int sms = sm_count();
    const int block = 256, warps_per_block = block / 32;
    const int64_t want = (T_ + 2 * warps_per_block - 1) / (2 * warps_per_block);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)sms * 8));
    const bool aligned = (((uintptr_t)a | (uintptr_t)p | (uintptr_t)n) % 16 == 0) && ((ld * sizeof(T)) % 16 == 0);
    const int vpl = (D % (32 * kVec) == 0) ? D / (32 * kVec) : 0;
    const T *A = (const T*)a, *P = (const T*)p, *N = (const T*)n;
#define PB2_LAUNCH_VPL(V)                                                                                   \
    triplet_vec_kernel<T, V><<<grid, block, 0, st>>>(A, P, N, ia, ip, in_, T_, ld, discrete, out)
    if (aligned && vpl == 1) PB2_LAUNCH_VPL(1);
    else if (aligned && vpl == 2) PB2_LAUNCH_VPL(2);
    else if (aligned && vpl == 3) PB2_LAUNCH_VPL(3);
    else if (aligned && vpl == 4) PB2_LAUNCH_VPL(4);
    else {
        const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>((T_ + warps_per_block - 1) / warps_per_block,
                                                                   (int64_t)sms * 8));
        triplet_generic_kernel<T><<<g2, block, 0, st>>>(A, P, N, ia, ip, in_, T_, D, ld, discrete, out);
    }
#undef PB2_LAUNCH_VPL
    return check_launch("triplet_score");
}

}  // namespace pb2

extern "C" int pb2_triplet_score(const void* anchor, const void* positive, const void* negative,
                                 const int64_t* anchor_idx, const int64_t* positive_idx, const int64_t* negative_idx,
                                 int64_t n_triplets, int dim, int64_t ld, int dtype, int discrete, float* out,
                                 void* stream) {
    using namespace pb2;
    if (n_triplets < 0 || dim <= 0 || ld < dim) return set_error(PB2_ERR_ARG, "triplet_score: bad sizes");
    if (n_triplets == 0) return PB2_OK;
    if (!anchor || !positive || !negative || !out) return set_error(PB2_ERR_ARG, "triplet_score: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case PB2_BF16:
            return launch_triplet<__nv_bfloat16>(anchor, positive, negative, anchor_idx, positive_idx, negative_idx,
                                                 n_triplets, dim, ld, discrete, out, st);
        case PB2_F16:
            return launch_triplet<__half>(anchor, positive, negative, anchor_idx, positive_idx, negative_idx,
                                          n_triplets, dim, ld, discrete, out, st);
        case PB2_F32:
            return launch_triplet<float>(anchor, positive, negative, anchor_idx, positive_idx, negative_idx,
                                         n_triplets, dim, ld, discrete, out, st);
    }
    return set_error(PB2_ERR_ARG, "triplet_score: unknown dtype");
}
