// Encoder tail: the last two stages of both reference encoders,
//     project = nn.Linear(n_in, 512)   then   F.normalize(x, p=2, dim=1)
// (pig/models.py:96-109 Wav2VecEncoder, :130-150 R3DEncoder), fused into ONE tcgen05 kernel that emits what
// the scoring kernels consume: L2-normalised bf16 rows and the fp32 1/||row|| of those rounded rows (the
// re-normalisation of pig/util.py:11-12 without another pass over the embeddings).
//
//   y = x W^T + b            x [rows, n_in] bf16, W [n_out, n_in] bf16 (nn.Linear layout), fp32 accumulate
//   e = y / max(||y||, eps)  eps = 1e-12 (F.normalize);  out = bf16(e)
//   rinv = 1 / ||out||       norm = ||y||  (the backward's normalisation Jacobian needs it)
//
// One CTA owns 128 rows and ALL n_out <= 512 output features: the 128 x 512 fp32 accumulator is the whole
// of TMEM, so the row norm is complete inside the CTA and y never exists in HBM.  x tiles [128 x 64] and
// W tiles [n_out x 64] stream through a two-stage TMA ring; one thread issues the MMAs (two of N <= 256
// per k-step); four epilogue warps (thread == row) read the accumulator twice -- sum of squares, then
// scale / round / stage -- and the bf16 tile leaves through swizzled slabs and TMA stores.
#include "common.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {
namespace proj {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int kEpiWarps = 4, kMmaWarp = 4, kTmaWarp = 5;
constexpr int kThreads = 6 * 32;
constexpr int kMaxOut = 512;
constexpr int kXBytes = BM * BK * 2;           // 16 KiB
constexpr int kWBytes = kMaxOut * BK * 2;      // 64 KiB (rows beyond n_out are zero-filled by TMA)
constexpr int kStageBytes = kXBytes + kWBytes;
constexpr int kStages = 2;
constexpr int kSlabBytes = 32 * 64 * 2;        // one warp's [32 rows x 64 bf16] staging slab
constexpr int kOutBytes = kEpiWarps * 2 * kSlabBytes;
constexpr int kBiasBytes = kMaxOut * 4;
constexpr int kSmem = kStages * kStageBytes + kOutBytes + kBiasBytes + 256;

struct Args {
    int64_t rows;
    int n_out, kblocks;
    uint32_t tx_bytes;  // bytes one stage's TMA loads deliver
    const float* bias;  // may be null
    float eps;
    float* rinv;        // may be null
    float* norm;        // may be null
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(kThreads, 1)
    project_normalize_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                             const __grid_constant__ CUtensorMap tm_out, const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* out_stage = smem + kStages * kStageBytes;
    float* bias_s = reinterpret_cast<float*>(out_stage + kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + kBiasBytes);
    uint64_t* full = bars;                // [kStages]
    uint64_t* empty = bars + kStages;     // [kStages]
    uint64_t* acc_full = empty + kStages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_tiles = (a.rows + BM - 1) / BM;
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kEpiWarps);
        fence_mbar_init();
    }
    if (warp == kTmaWarp) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < kMaxOut; i += kThreads) bias_s[i] = (a.bias && i < a.n_out) ? a.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n1 = a.n_out > 256 ? 256 : a.n_out, n2 = a.n_out - n1;  // N of the two MMAs of a k-step

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + stage * kStageBytes;
                    uint8_t* sw = sx + kXBytes;
                    mbar_arrive_expect_tx(full + stage, a.tx_bytes);
                    tma_load_2d(sx, &tm_x, full + stage, kb * BK, (int32_t)(t * BM), kEvictFirst);
                    tma_load_2d(sw, &tm_w, full + stage, kb * BK, 0, kEvictLast);
                    if (n2 > 0) tma_load_2d(sw + kWBytes / 2, &tm_w, full + stage, kb * BK, 256, kEvictLast);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            const uint32_t idesc1 = make_idesc(BM, (uint32_t)n1, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            const uint32_t idesc2 = n2 > 0 ? make_idesc(BM, (uint32_t)n2, kFmtBF16, kFmtBF16, kMajorK, kMajorK) : 0u;
            const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            constexpr uint32_t kStageLo = kStageBytes >> 4, kWLo = kXBytes >> 4, kW2Lo = (kXBytes + kWBytes / 2) >> 4,
                               kKLo = (UK * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0, lo = lo0;
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        umma_f16_lohi(tmem_base, lo + k * kKLo, lo + kWLo + k * kKLo, desc_hi, idesc1, acc);
                        if (n2 > 0) umma_f16_lohi(tmem_base + 256, lo + k * kKLo, lo + kW2Lo + k * kKLo, desc_hi, idesc2, acc);
                    }
                    umma_commit(empty + stage);
                    lo += kStageLo;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                        lo = lo0;
                    }
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue: thread == row, two passes over the 128 x n_out accumulator
        const int quad = warp & 3;
        uint8_t* slab = out_stage + warp * 2 * kSlabBytes;
        uint32_t n_slab = 0;
        const int n_chunks = a.n_out / 32;
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const int64_t row = t * BM + quad * 32 + lane;
            mbar_wait(acc_full, (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
            float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_lane + ch * 32, v);
                tmem_ld_wait();
                const float* b = bias_s + ch * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float y0 = __uint_as_float(v[j]) + b[j], y1 = __uint_as_float(v[j + 1]) + b[j + 1];
                    const float y2 = __uint_as_float(v[j + 2]) + b[j + 2], y3 = __uint_as_float(v[j + 3]) + b[j + 3];
                    ss0 = fmaf(y0, y0, ss0);
                    ss1 = fmaf(y1, y1, ss1);
                    ss2 = fmaf(y2, y2, ss2);
                    ss3 = fmaf(y3, y3, ss3);
                }
            }
            const float nrm = sqrtf((ss0 + ss1) + (ss2 + ss3));
            const float scale = 1.0f / fmaxf(nrm, a.eps);  // F.normalize: x / max(||x||, eps)
            float q0 = 0.f, q1 = 0.f;
            for (int ch = 0; ch < n_chunks; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_lane + ch * 32, v);
                tmem_ld_wait();
                const float* b = bias_s + ch * 32;
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const float e0 = (__uint_as_float(v[j]) + b[j]) * scale, e1 = (__uint_as_float(v[j + 1]) + b[j + 1]) * scale;
                    const uint32_t pk = pack_bf16(e0, e1);
                    packed[j >> 1] = pk;
                    const float r0 = __uint_as_float(pk << 16), r1 = __uint_as_float(pk & 0xffff0000u);
                    q0 = fmaf(r0, r0, q0);
                    q1 = fmaf(r1, r1, q1);
                }
                uint8_t* sl = slab + (n_slab & 1) * kSlabBytes;
                if ((ch & 1) == 0) {  // the slab about to be rewritten must have been read by its TMA store
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
                uint8_t* srow = sl + lane * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c16 = ((ch & 1) * 4 + k) ^ (lane & 7);  // 128-byte swizzle
                    *reinterpret_cast<uint4*>(srow + c16 * 16) =
                        make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                }
                if (ch & 1) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm_out, sl, (ch - 1) * 32, (int32_t)(t * BM + quad * 32));
                        tma_store_commit();
                    }
                    ++n_slab;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            if (row < a.rows) {
                if (a.rinv) a.rinv[row] = 1.0f / sqrtf(q0 + q1);  // no epsilon: the scoring kernels' convention
                if (a.norm) a.norm[row] = nrm;
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace proj
}  // namespace pb2

using namespace pb2;

extern "C" int pb2_project_normalize(const void* x, const void* w, const float* bias, int64_t rows, int n_in, int n_out,
                                     int64_t ldx, int64_t ldw, float eps, void* out, int64_t ld_out, float* rinv,
                                     float* norm, void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!x || !w || !out) return set_error(PB2_ERR_ARG, "project_normalize: null");
    if (n_in <= 0 || n_in % 64 != 0) return set_error(PB2_ERR_ARG, "project_normalize: n_in must be a positive multiple of 64");
    if (n_out <= 0 || n_out % 64 != 0 || n_out > proj::kMaxOut)
        return set_error(PB2_ERR_ARG, "project_normalize: n_out must be a multiple of 64, at most 512");
    if (ld_out % 8 != 0 || ld_out < n_out) return set_error(PB2_ERR_ARG, "project_normalize: ld_out %% 8 == 0, >= n_out");
    CUtensorMap tx, tw, to;
    int rc = make_tmap_2d(&tx, x, 2, (uint64_t)rows, (uint64_t)n_in, (uint64_t)ldx * 2, proj::BM, proj::BK);
    if (rc) return rc;
    const int w_box = n_out < 256 ? n_out : 256;  // W rows per TMA box (a second box covers rows 256..511)
    rc = make_tmap_2d(&tw, w, 2, (uint64_t)n_out, (uint64_t)n_in, (uint64_t)ldw * 2, (uint32_t)w_box, proj::BK);
    if (rc) return rc;
    rc = make_tmap_2d(&to, out, 2, (uint64_t)rows, (uint64_t)n_out, (uint64_t)ld_out * 2, 32, 64);
    if (rc) return rc;
    proj::Args a;
    a.rows = rows;
    a.n_out = n_out;
    a.kblocks = n_in / proj::BK;
    a.tx_bytes = (uint32_t)(proj::kXBytes + w_box * proj::BK * 2 * (n_out > 256 ? 2 : 1));
    a.bias = bias;
    a.eps = eps;
    a.rinv = rinv;
    a.norm = norm;
    static bool configured = false;
    if (!configured) {
        rc = check_cuda(cudaFuncSetAttribute(proj::project_normalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             proj::kSmem),
                        "project_normalize");
        if (rc) return rc;
        configured = true;
    }
    const int64_t n_tiles = (rows + proj::BM - 1) / proj::BM;
    const int grid = (int)std::min<int64_t>(n_tiles, sm_count());
    proj::project_normalize_kernel<<<grid, proj::kThreads, proj::kSmem, (cudaStream_t)stream>>>(tx, tw, to, a);
    return check_launch("project_normalize");
}
