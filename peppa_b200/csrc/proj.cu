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
// of TMEM, so the row norm is complete inside the CTA and y never exists in HBM.  CTAs work in pairs
// (tcgen05 cta_group::2, one MMA of M = 256 per two SMs).  Two kernels:
//   * resident-x phased kernel (n_in <= 512, n_out % 256 == 0 -- the reference's Linear(512, 512)): the pair's x tile
//     stays in shared memory, W streams, and the accumulator is produced in two phases of 256 columns with their own
//     full / empty barriers, so the two-pass epilogue of one phase runs under the MMAs of the other;
//   * interleaved kernel (every other shape): x and W tiles stream through a three-stage TMA ring, two MMAs of N <= 256
//     per k-step fill all of TMEM, then the epilogue runs (MMA and epilogue do not overlap).
// Eight epilogue warps (thread == row, two column halves per TMEM lane quadrant) read the accumulator twice -- sum of
// squares, then scale / round / stage -- with packed fp32 math; the bf16 tile leaves through 128-byte-swizzled slabs
// and TMA stores.
#include "common.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {
namespace proj {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int kEpiWarps = 8, kMmaWarp = 8, kTmaWarp = 9;  // the column-split / phased variants: two epilogue warps per TMEM lane quadrant
constexpr int kThreads = 10 * 32;
// The product kernel is templated on kParts = epilogue warps per TMEM lane quadrant (each takes a contiguous run of
// 64-column slabs): 4 * kParts epilogue warps, then the MMA warp, then the TMA warp.
template <int kParts> struct Cfg {
    static constexpr int kEpi = 4 * kParts, kMma = kEpi, kTma = kEpi + 1, kThr = (kEpi + 2) * 32;
    static constexpr int kSlabs = kParts == 2 ? 2 : 1;   // staging slabs per warp (64 KiB in all either way)
    static constexpr int kRed = 2 * kParts * 128 * 4;     // per-row partial sums: [ss | q][part][row]
};
constexpr int kMaxOut = 512;
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one MMA of M = 256 spans the 128 rows of both CTAs; each CTA
// stages its own x rows and HALF of W's rows for each of the two MMAs of a k-step, so a stage is 48 KiB instead
// of 80 and W crosses L2 -> SM once per pair (the kernel was bound by that traffic, not by the tensor pipe).
constexpr int kXBytes = BM * BK * 2;           // 16 KiB
constexpr int kWHalf = 128 * BK * 2;           // 16 KiB: this CTA's rows of one MMA's W tile (N <= 256 per MMA)
constexpr int kWBytes = 2 * kWHalf;
constexpr int kStageBytes = kXBytes + kWBytes;
constexpr int kStages = 3;
constexpr int kWBox = 32;                      // W rows per TMA box
constexpr int kSlabBytes = 32 * 64 * 2;        // one warp's [32 rows x 64 bf16] staging slab
constexpr int kOutBytes = kEpiWarps * 2 * kSlabBytes;
constexpr int kBiasBytes = 0;                   // the bias is read through L1 (broadcast loads): shared memory is full
constexpr int kRedBytes = 2 * 2 * BM * 4;      // per-row partial sums of the two column halves: [ss | q][half][row]
constexpr int kSmem = kStages * kStageBytes + kOutBytes + kBiasBytes + 2 * kRedBytes + 256;  // room for kParts = 4

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

template <int kParts>
__global__ void __launch_bounds__(Cfg<kParts>::kThr, 1)
    project_normalize_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                             const __grid_constant__ CUtensorMap tm_out, const Args a) {
    using C = Cfg<kParts>;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* out_stage = smem + kStages * kStageBytes;
    float* red_s = reinterpret_cast<float*>(out_stage + kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red_s) + C::kRed);
    uint64_t* full = bars;                // [kStages]
    uint64_t* empty = bars + kStages;     // [kStages]
    uint64_t* acc_full = empty + kStages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();  // 0 = leader (issues the MMAs)
    const int64_t n_tiles = (a.rows + 2 * BM - 1) / (2 * BM);  // pair tiles of 256 rows
    const int64_t unit0 = blockIdx.x / 2, n_units = gridDim.x / 2;
    if (warp == C::kTma && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == C::kMma && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 2 * C::kEpi);  // the leader's collects both CTAs' epilogues
        fence_mbar_init();
    }
    if (warp == C::kTma) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n1 = a.n_out > 256 ? 256 : a.n_out, n2 = a.n_out - n1;  // N of the two MMAs of a k-step

    if (warp == C::kTma) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int h1 = n1 / 2, h2 = n2 / 2;  // this CTA's W rows per MMA
            for (int64_t t = unit0; t < n_tiles; t += n_units) {
                const int32_t xrow = (int32_t)((t * 2 + crank) * BM);
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + stage * kStageBytes;
                    uint8_t* sw = sx + kXBytes;
                    if (crank == 0) mbar_arrive_expect_tx(full + stage, a.tx_bytes);  // both CTAs' bytes
                    const uint32_t lbar = mapa_u32(smem_u32(full + stage), 0);
                    tma_load_2d_pair(sx, &tm_x, lbar, kb * BK, xrow, kEvictFirst);
                    for (int r0 = 0; r0 < h1; r0 += kWBox)
                        tma_load_2d_pair(sw + r0 * (BK * 2), &tm_w, lbar, kb * BK, (int32_t)crank * h1 + r0, kEvictLast);
                    for (int r0 = 0; r0 < h2; r0 += kWBox)
                        tma_load_2d_pair(sw + kWHalf + r0 * (BK * 2), &tm_w, lbar, kb * BK, 256 + (int32_t)crank * h2 + r0, kEvictLast);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == C::kMma) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc1 = make_idesc(2 * BM, (uint32_t)n1, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            const uint32_t idesc2 = n2 > 0 ? make_idesc(2 * BM, (uint32_t)n2, kFmtBF16, kFmtBF16, kMajorK, kMajorK) : 0u;
            const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            constexpr uint32_t kStageLo = kStageBytes >> 4, kWLo = kXBytes >> 4, kW2Lo = (kXBytes + kWHalf) >> 4,
                               kKLo = (UK * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0, lo = lo0;
            int64_t it = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
                mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
                tc_fence_after();
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        umma_f16_pair_lohi(tmem_base, lo + k * kKLo, lo + kWLo + k * kKLo, desc_hi, desc_hi, idesc1, acc);
                        if (n2 > 0)
                            umma_f16_pair_lohi(tmem_base + 256, lo + k * kKLo, lo + kW2Lo + k * kKLo, desc_hi, desc_hi, idesc2, acc);
                    }
                    umma_commit_pair(empty + stage);
                    lo += kStageLo;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                        lo = lo0;
                    }
                }
                umma_commit_pair(acc_full);
            }
        }
    } else {
        // ===== epilogue: thread == row; the kParts warps of a TMEM lane quadrant take contiguous runs of 64-column
        // slabs; two passes over the accumulator (sum of squares, then scale / round / stage), the parts' per-row
        // partial sums meet in shared memory in a fixed order (deterministic).  With the single 128 x 512 accumulator
        // stage the epilogue is NOT overlapped with the next tile's MMAs, so its length counts in full: kParts = 2
        // hides the TMEM round trips with two register buffers per warp, kParts = 4 with four warps per scheduler.
        const int quad = warp & 3, part = warp >> 2;
        uint8_t* slab = out_stage + warp * C::kSlabs * kSlabBytes;
        uint32_t n_slab = 0;
        const int pairs = a.n_out / 64, base = pairs / kParts, rem = pairs % kParts;
        const int ch0 = 2 * (part * base + (part < rem ? part : rem));
        const int ch1 = ch0 + 2 * (base + (part < rem ? 1 : 0));
        const int r = quad * 32 + lane;
        const bool has_bias = a.bias != nullptr;
        int64_t it = 0;
        for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
            const int64_t row0 = (t * 2 + crank) * BM, row = row0 + r;
            mbar_wait(acc_full, (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
            float2 ssa = make_float2(0.f, 0.f), ssb = make_float2(0.f, 0.f);
            auto pass1 = [&](const uint32_t (&v)[32], int ch) {
                const float4* b4 = reinterpret_cast<const float4*>(a.bias + ch * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float2 y01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y));
                    const float2 y23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w));
                    ssa = __ffma2_rn(y01, y01, ssa);
                    ssb = __ffma2_rn(y23, y23, ssb);
                }
            };
            float2 qa = make_float2(0.f, 0.f);
            float2 sc2;
            auto pass2 = [&](const uint32_t (&v)[32], int ch) {
                const float4* b4 = reinterpret_cast<const float4*>(a.bias + ch * 32);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float2 e01 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y)), sc2);
                    const float2 e23 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w)), sc2);
                    const uint32_t p0 = pack_bf16(e01.x, e01.y), p1 = pack_bf16(e23.x, e23.y);
                    packed[2 * j] = p0;
                    packed[2 * j + 1] = p1;
                    const float2 r01 = make_float2(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u));
                    const float2 r23 = make_float2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
                    qa = __ffma2_rn(r01, r01, qa);
                    qa = __ffma2_rn(r23, r23, qa);
                }
                const int cp = (ch - ch0) & 1;
                uint8_t* sl = slab + (C::kSlabs == 2 ? (n_slab & 1) : 0u) * kSlabBytes;
                if (cp == 0) {  // the slab about to be rewritten must have been read by its TMA store
                    if (lane == 0) tma_store_wait_read<C::kSlabs - 1>();
                    __syncwarp();
                }
                uint8_t* srow = sl + lane * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c16 = (cp * 4 + k) ^ (lane & 7);  // 128-byte swizzle
                    *reinterpret_cast<uint4*>(srow + c16 * 16) =
                        make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                }
                if (cp) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm_out, sl, (ch - 1) * 32, (int32_t)(row0 + quad * 32));
                        tma_store_commit();
                    }
                    ++n_slab;
                }
            };
            float nrm;
            auto row_scale = [&]() {  // after the barrier: the parts' sums of squares in a fixed order
                float ss = red_s[r];
#pragma unroll
                for (int p = 1; p < kParts; ++p) ss += red_s[p * BM + r];
                nrm = sqrtf(ss);
                const float scale = 1.0f / fmaxf(nrm, a.eps);  // F.normalize: x / max(||x||, eps)
                sc2 = make_float2(scale, scale);
            };
            if constexpr (kParts == 2) {
                // Both passes keep the TMEM load of chunk c + 1 in flight while chunk c is processed (two register
                // buffers, like sim.cu).
                uint32_t va[32], vb[32];
                if (ch0 < ch1) tmem_ld32(t_lane + ch0 * 32, va);
#pragma unroll 1
                for (int ch = ch0; ch < ch1; ch += 2) {  // each part holds an even number of chunks
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(t_lane + (ch + 1) * 32, vb);
                    pass1(va, ch);
                    tmem_ld_wait();
                    __syncwarp();
                    if (ch + 2 < ch1) tmem_ld32(t_lane + (ch + 2) * 32, va);
                    pass1(vb, ch + 1);
                }
                red_s[part * BM + r] = (ssa.x + ssa.y) + (ssb.x + ssb.y);
                if (ch0 < ch1) tmem_ld32(t_lane + ch0 * 32, va);  // pass 2's first chunk travels across the barrier
                named_bar_sync(1, C::kEpi * 32);
                row_scale();
#pragma unroll 1
                for (int ch = ch0; ch < ch1; ch += 2) {
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(t_lane + (ch + 1) * 32, vb);
                    pass2(va, ch);
                    tmem_ld_wait();
                    __syncwarp();
                    if (ch + 2 < ch1) tmem_ld32(t_lane + (ch + 2) * 32, va);
                    pass2(vb, ch + 1);
                }
            } else {
                uint32_t va[32];
#pragma unroll 1
                for (int ch = ch0; ch < ch1; ++ch) {
                    tmem_ld32(t_lane + ch * 32, va);
                    tmem_ld_wait();
                    pass1(va, ch);
                }
                red_s[part * BM + r] = (ssa.x + ssa.y) + (ssb.x + ssb.y);
                if (ch0 < ch1) tmem_ld32(t_lane + ch0 * 32, va);  // pass 2's first chunk travels across the barrier
                named_bar_sync(1, C::kEpi * 32);
                row_scale();
#pragma unroll 1
                for (int ch = ch0; ch < ch1; ++ch) {
                    tmem_ld_wait();
                    __syncwarp();
                    pass2(va, ch);
                    if (ch + 1 < ch1) tmem_ld32(t_lane + (ch + 1) * 32, va);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty), 0));  // the leader's MMA warp
            red_s[(kParts + part) * BM + r] = qa.x + qa.y;
            named_bar_sync(1, C::kEpi * 32);
            if (part == 0 && row < a.rows) {
                float q = red_s[kParts * BM + r];
#pragma unroll
                for (int p = 1; p < kParts; ++p) q += red_s[(kParts + p) * BM + r];
                if (a.rinv) a.rinv[row] = 1.0f / sqrtf(q);  // no epsilon: the scoring kernels' convention
                if (a.norm) a.norm[row] = nrm;
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still signals it
    if (warp == C::kTma) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}


// ---------------------------------------------------------------------------------------------------------
// Resident-x phased kernel (n_in <= 512, n_out % 128 == 0: the reference's Linear(512, 512) tails).
//
// What bounds the interleaved kernel above is that its two halves cannot overlap: the 128 x 512 fp32 accumulator is
// all of TMEM, so a tile costs an MMA phase (8192 tensor cycles, and 384 KB per SM through L2 at the chip's
// ~43 B/cycle/SM: ~9 k cycles) PLUS a two-pass epilogue that reads 2 x 256 KB out of TMEM at its 64 B/cycle (>= 8192
// cycles, whatever the number of epilogue warps: sixteen warps measured the same as eight).  Here the accumulator is
// produced in PHASES of 128 output columns, one after the other, each with its own full / empty barrier:
//     MMA_0 | MMA_1 | MMA_2 | MMA_3           | MMA_0' | MMA_1' ...
//           | p1_0  | p1_1  | p1_2  p1_3 p2_0 | p2_1   | p2_2 ...
// pass 1 (sum of squares) of phase q runs under MMA_{q+1}; the next tile's MMA_q starts as soon as pass 2 (scale /
// round / store) has drained phase q.  Only pass 1 of the last phase and pass 2 of the first are exposed.  The phased
// variant of the measurement build streamed x once per phase and lost to L2 what it won; here the pair's x tile
// (128 x n_in bf16 per CTA, <= 128 KiB) STAYS in shared memory for all phases -- slot kb is refilled with the next
// tile's k-block as soon as the last phase has consumed it -- and only W streams (8 KiB stages: this CTA's 64 rows of
// a phase's 128, one k-block; eight stages in flight), so a tile moves the same 384 KB per SM as before.
constexpr int kR_MaxKb = 8;                        // n_in <= 512
constexpr int kR_WBytes = 64 * 1024;               // the W ring
constexpr int kR_Epi = 8, kR_Mma = 8, kR_TmaW = 9, kR_TmaX = 10, kR_Threads = 11 * 32;
constexpr int kR_OutBytes = kR_Epi * kSlabBytes;   // one staging slab per epilogue warp
constexpr int kR_Smem = kR_MaxKb * kXBytes + kR_WBytes + kR_OutBytes + kRedBytes + 512;
static_assert(kR_Smem <= 227 * 1024, "resident encoder tail: shared memory");

// kPW: output columns per phase = N of its MMAs.  256 is the product; 128 (four phases, more overlap on paper) is a
// measurement option and 35 % SLOWER: an M = 256 pair MMA of N = 128 reads 6 KB of operands from each CTA's shared
// memory per 64 tensor cycles (96 B/cycle, against 64 B/cycle at N = 256), which together with the TMA fill and the
// epilogue's staging exceeds the 128 B/cycle a shared memory delivers -- the tensor pipe waits for its A operand.
template <int kPW>
__global__ void __launch_bounds__(kR_Threads, 1)
    project_normalize_resident_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                      const __grid_constant__ CUtensorMap tm_out, const Args a) {
    constexpr int kWStage = (kPW / 2) * BK * 2;    // this CTA's half of a phase's W rows, one k-block
    constexpr int kWStages = kR_WBytes / kWStage;
    constexpr int kMaxPhases = kMaxOut / kPW;
    constexpr int kSL = kPW / 128;                 // 64-column slabs per phase and warp
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* xs = smem;
    uint8_t* ws = xs + kR_MaxKb * kXBytes;
    uint8_t* out_stage = ws + kR_WBytes;
    float* red_s = reinterpret_cast<float*>(out_stage + kR_OutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red_s) + kRedBytes);
    uint64_t* x_full = bars;                      // [kR_MaxKb]
    uint64_t* x_empty = x_full + kR_MaxKb;        // [kR_MaxKb]
    uint64_t* w_full = x_empty + kR_MaxKb;        // [kWStages]
    uint64_t* w_empty = w_full + kWStages;        // [kWStages]
    uint64_t* acc_full = w_empty + kWStages;      // [kMaxPhases]
    uint64_t* acc_empty = acc_full + kMaxPhases;  // [kMaxPhases]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kMaxPhases);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int64_t n_tiles = (a.rows + 2 * BM - 1) / (2 * BM);
    const int64_t unit0 = blockIdx.x / 2, n_units = gridDim.x / 2;
    const int n_phases = a.n_out / kPW, n_kb = a.kblocks;
    if (warp == kR_TmaW && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == kR_Mma && lane == 0) {
        for (int s = 0; s < kR_MaxKb; ++s) {
            mbar_init(x_full + s, 1);
            mbar_init(x_empty + s, 1);
        }
        for (int s = 0; s < kWStages; ++s) {
            mbar_init(w_full + s, 1);
            mbar_init(w_empty + s, 1);
        }
        for (int q = 0; q < kMaxPhases; ++q) {
            mbar_init(acc_full + q, 1);
            mbar_init(acc_empty + q, 2 * kR_Epi);  // the leader's collects both CTAs' epilogues
        }
        fence_mbar_init();
    }
    if (warp == kR_TmaW) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kR_TmaW) {
        if (lane == 0) {  // W: [phase][k-block] stages, this CTA's half of the phase's rows
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units)
                for (int q = 0; q < n_phases; ++q)
                    for (int kb = 0; kb < n_kb; ++kb) {
                        mbar_wait(w_empty + stage, phase ^ 1);
                        if (crank == 0) mbar_arrive_expect_tx(w_full + stage, 2 * kWStage);  // both CTAs' bytes
                        tma_load_2d_pair(ws + stage * kWStage, &tm_w, mapa_u32(smem_u32(w_full + stage), 0), kb * BK,
                                         q * kPW + (int32_t)crank * (kPW / 2), kEvictLast);
                        if (++stage == kWStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
        }
    } else if (warp == kR_TmaX) {
        if (lane == 0) {  // x: slot kb is refilled once the last phase of the previous tile has consumed it
            int64_t it = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
                const int32_t xrow = (int32_t)((t * 2 + crank) * BM);
                for (int kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(x_empty + kb, (uint32_t)(it & 1) ^ 1);
                    if (crank == 0) mbar_arrive_expect_tx(x_full + kb, 2 * kXBytes);
                    tma_load_2d_pair(xs + kb * kXBytes, &tm_x, mapa_u32(smem_u32(x_full + kb), 0), kb * BK, xrow, kEvictFirst);
                }
            }
        }
    } else if (warp == kR_Mma) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = make_idesc(2 * BM, (uint32_t)kPW, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            const uint64_t d0 = make_smem_desc(smem_u32(xs), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), xlo0 = (uint32_t)d0;
            const uint32_t wlo0 = (uint32_t)make_smem_desc(smem_u32(ws), 16, 1024);
            constexpr uint32_t kXLo = kXBytes >> 4, kWLo = kWStage >> 4, kKLo = (UK * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0, wlo = wlo0;
            int64_t it = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
                for (int q = 0; q < n_phases; ++q) {
                    mbar_wait(acc_empty + q, (uint32_t)(it & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)(q * kPW);
                    uint32_t xlo = xlo0;
                    for (int kb = 0; kb < n_kb; ++kb) {
                        if (q == 0) mbar_wait(x_full + kb, (uint32_t)(it & 1));
                        mbar_wait(w_full + stage, phase);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_f16_pair_lohi(d_tmem, xlo + k * kKLo, wlo + k * kKLo, desc_hi, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair(w_empty + stage);
                        if (q == n_phases - 1) umma_commit_pair(x_empty + kb);
                        xlo += kXLo;
                        wlo += kWLo;
                        if (++stage == kWStages) {
                            stage = 0;
                            phase ^= 1;
                            wlo = wlo0;
                        }
                    }
                    umma_commit_pair(acc_full + q);
                }
            }
        }
    } else {
        // ===== epilogue: thread == row; of a phase's kPW columns warps 0-3 take the first half (kSL 64-column slabs), warps
        // 4-7 the second; the halves' per-row partial sums meet in shared memory in a fixed order (deterministic).  A
        // slab's two chunks are loaded and waited for together; one 4 KiB staging slab per warp.  (Tried and dropped:
        // chunk-wise register ping-pong with 2 KiB half slabs under the 64-byte swizzle, so that a store's read of the
        // slab overlaps the next chunk's arithmetic -- correct, 8 % slower: twice the TMA stores.)
        const int quad = warp & 3, sub = warp >> 2;
        uint8_t* sl = out_stage + warp * kSlabBytes;
        const int r = quad * 32 + lane;
        const bool has_bias = a.bias != nullptr;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        const int n_slabs = n_phases * kSL;   // this warp's slabs of a tile; slab i: phase i / kSL, columns below
        auto slab_col = [&](int i) { return (i / kSL) * kPW + (sub * kSL + i % kSL) * 64; };
        int64_t it = 0;
        for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
            const int64_t row0 = (t * 2 + crank) * BM, row = row0 + r;
            const uint32_t par = (uint32_t)(it & 1);
            float2 ssa = make_float2(0.f, 0.f), ssb = make_float2(0.f, 0.f);
            uint32_t va[32], vb[32];
            auto pass1 = [&](const uint32_t (&v)[32], int col) {
                const float4* b4 = reinterpret_cast<const float4*>(a.bias + col);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float2 y01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y));
                    const float2 y23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w));
                    ssa = __ffma2_rn(y01, y01, ssa);
                    ssb = __ffma2_rn(y23, y23, ssb);
                }
            };
            // ---- pass 1: sum of squares, phase q under the MMAs of phase q + 1
#pragma unroll 1
            for (int i = 0; i < n_slabs; ++i) {
                if (i % kSL == 0) {
                    mbar_wait(acc_full + i / kSL, par);
                    tc_fence_after();
                }
                const int col = slab_col(i);
                tmem_ld32(t_lane + col, va);
                tmem_ld32(t_lane + col + 32, vb);
                tmem_ld_wait();
                pass1(va, col);
                pass1(vb, col + 32);
            }
            red_s[sub * BM + r] = (ssa.x + ssa.y) + (ssb.x + ssb.y);
            tmem_ld32(t_lane + slab_col(0), va);  // pass 2's first loads travel across the barrier
            tmem_ld32(t_lane + slab_col(0) + 32, vb);
            named_bar_sync(1, kR_Epi * 32);
            const float nrm = sqrtf(red_s[r] + red_s[BM + r]);
            const float scale = 1.0f / fmaxf(nrm, a.eps);  // F.normalize: x / max(||x||, eps)
            const float2 sc2 = make_float2(scale, scale);
            float2 qa = make_float2(0.f, 0.f);
            auto pass2 = [&](const uint32_t (&v)[32], int col, int cp) {
                const float4* b4 = reinterpret_cast<const float4*>(a.bias + col);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float2 e01 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y)), sc2);
                    const float2 e23 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w)), sc2);
                    const uint32_t p0 = pack_bf16(e01.x, e01.y), p1 = pack_bf16(e23.x, e23.y);
                    packed[2 * j] = p0;
                    packed[2 * j + 1] = p1;
                    const float2 r01 = make_float2(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u));
                    const float2 r23 = make_float2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
                    qa = __ffma2_rn(r01, r01, qa);
                    qa = __ffma2_rn(r23, r23, qa);
                }
                if (cp == 0) {  // the slab's previous store must have read it (waited for after this chunk's arithmetic; before it: same time)
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                }
                uint8_t* srow = sl + lane * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c16 = (cp * 4 + k) ^ (lane & 7);  // 128-byte swizzle
                    *reinterpret_cast<uint4*>(srow + c16 * 16) =
                        make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                }
            };
            // ---- pass 2: scale / round / stage / store; phase q goes back to the next tile's MMAs as soon as it is read
#pragma unroll 1
            for (int i = 0; i < n_slabs; ++i) {
                const int col = slab_col(i);
                tmem_ld_wait();
                if (i % kSL == kSL - 1) tc_fence_before();
                __syncwarp();
                if (lane == 0 && i % kSL == kSL - 1)
                    mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty + i / kSL), 0));  // the phase is in registers
                __syncwarp();
                pass2(va, col, 0);
                pass2(vb, col + 32, 1);
                if (i + 1 < n_slabs) {
                    tmem_ld32(t_lane + slab_col(i + 1), va);
                    tmem_ld32(t_lane + slab_col(i + 1) + 32, vb);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_out, sl, col, (int32_t)(row0 + quad * 32));
                    tma_store_commit();
                }
            }
            red_s[2 * BM + sub * BM + r] = qa.x + qa.y;
            named_bar_sync(1, kR_Epi * 32);
            if (sub == 0 && row < a.rows) {
                if (a.rinv) a.rinv[row] = 1.0f / sqrtf(red_s[2 * BM + r] + red_s[3 * BM + r]);  // no epsilon: the scoring kernels' convention
                if (a.norm) a.norm[row] = nrm;
            }
        }
    
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still signals it
    if (warp == kR_TmaW) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Column-split variant (n_out % 128 == 0; measured option, see g_variant below): the two CTAs of a cluster take the SAME 128 rows and half of the
// output columns each.  A CTA's accumulator is then 128 x n_out/2 <= 256 TMEM columns, i.e. TWO stages fit:
// the MMA of row tile t+1 overlaps the two-pass epilogue of tile t (the single-stage kernel above serialises
// them: ncu showed its tensor pipe 39 % busy).  Each CTA loads half of the x tile and multicasts it to both;
// W rows are private.  The row statistics (sum of squares of y, then of the rounded e) are completed across the
// pair through distributed shared memory: each CTA stores its per-row partial into the peer's buffer and
// arrives on the peer's mbarrier; partials are added as own + peer in both CTAs (commutative: identical bits).
constexpr int kS_XBytes = BM * BK * 2;               // 16 KiB (two multicast halves of 64 rows)
constexpr int kS_WBytes = 256 * BK * 2;              // 32 KiB: this CTA's W rows (n_out / 2 <= 256)
constexpr int kS_StageBytes = kS_XBytes + kS_WBytes;
constexpr int kS_Stages = 3;
constexpr int kS_XchgBytes = 2 * 2 * BM * 4;         // [ss | q][tile parity][row] partials written by the peer
constexpr int kS_BiasBytes = 256 * 4;                // this CTA's half of the bias
constexpr int kS_Smem = kS_Stages * kS_StageBytes + kOutBytes + kRedBytes + kS_XchgBytes + kS_BiasBytes + 256;

// One fp32 into the peer CTA's shared memory; its completion is counted (4 bytes) on the peer's mbarrier -- the
// distributed-shared-memory producer/consumer primitive: no release fence (MEMBAR.GPU) and no L1 invalidation
// (CCTL.IVALL), which a release-arrive / cluster fence pair costs twice per tile.
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr),
                 "r"(__float_as_uint(v)), "r"(cluster_bar)
                 : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
    project_normalize_split_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                   const __grid_constant__ CUtensorMap tm_out, const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* out_stage = smem + kS_Stages * kS_StageBytes;
    float* red_s = reinterpret_cast<float*>(out_stage + kOutBytes);
    float* xchg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(red_s) + kRedBytes);  // [kind][parity][row]
    float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(xchg) + kS_XchgBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(bias_s) + kS_BiasBytes);
    uint64_t* full = bars;                      // [kS_Stages]
    uint64_t* empty = full + kS_Stages;         // [kS_Stages]  count 2: both CTAs' MMAs have read the stage
    uint64_t* acc_full = empty + kS_Stages;     // [2]
    uint64_t* acc_empty = acc_full + 2;         // [2]
    uint64_t* xbar = acc_empty + 2;             // [kind][parity]: the peer's partial has landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbar + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank(), peer = crank ^ 1u;
    const int n_half = a.n_out / 2;             // this CTA's output columns [crank * n_half, +n_half)
    const int64_t n_tiles = (a.rows + BM - 1) / BM;
    const int64_t unit0 = blockIdx.x / 2, n_units = gridDim.x / 2;
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kS_Stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 2);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc_full + s, 1);
            mbar_init(acc_empty + s, kEpiWarps);
        }
        for (int s = 0; s < 4; ++s) mbar_init(xbar + s, 1);  // armed locally with the bytes the peer will deliver
        fence_mbar_init();
    }
    if (warp == kTmaWarp) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 256; i += kThreads) bias_s[i] = (a.bias && i < n_half) ? a.bias[crank * n_half + i] : 0.f;
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx = (uint32_t)(kS_XBytes + n_half * BK * 2);
            for (int64_t t = unit0; t < n_tiles; t += n_units) {
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + stage * kS_StageBytes;
                    uint8_t* sw = sx + kS_XBytes;
                    mbar_arrive_expect_tx(full + stage, tx);  // both x halves (own + the peer's multicast) + own W rows
                    tma_load_2d_mcast(sx + crank * (kS_XBytes / 2), &tm_x, full + stage, kb * BK, (int32_t)(t * BM + crank * 64),
                                      (uint16_t)3, kEvictFirst);
                    for (int r0 = 0; r0 < n_half; r0 += kWBox)
                        tma_load_2d(sw + r0 * (BK * 2), &tm_w, full + stage, kb * BK, (int32_t)crank * n_half + r0, kEvictLast);
                    if (++stage == kS_Stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, (uint32_t)n_half, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            constexpr uint32_t kStageLo = kS_StageBytes >> 4, kWLo = kS_XBytes >> 4, kKLo = (UK * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0, lo = lo0;
            int64_t it = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
                const int as = (int)(it & 1);
                mbar_wait(acc_empty + as, (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_f16_lohi(d_tmem, lo + k * kKLo, lo + kWLo + k * kKLo, desc_hi, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_mcast(empty + stage, (uint16_t)3);  // the x tile lives in both CTAs
                    lo += kStageLo;
                    if (++stage == kS_Stages) {
                        stage = 0;
                        phase ^= 1;
                        lo = lo0;
                    }
                }
                umma_commit(acc_full + as);
            }
        }
    } else {
        const int quad = warp & 3, half = warp >> 2;
        uint8_t* slab = out_stage + warp * 2 * kSlabBytes;
        uint32_t n_slab = 0;
        const int n_chunks = n_half / 32;                                       // even: n_out % 128 == 0
        const int split = (n_chunks + 1) / 2 / 2 * 2;
        const int ch0 = half == 0 ? 0 : split, ch1 = half == 0 ? split : n_chunks;
        const int r = quad * 32 + lane;
        // exchange of a per-row partial with the peer CTA: kind 0 = sum of squares of y, kind 1 = of the rounded e
        auto exchange = [&](int kind, int par, uint32_t wait_parity, float own) -> float {
            float* buf = xchg + (kind * 2 + par) * BM;
            uint64_t* bar = xbar + kind * 2 + par;
            if (half == 0) st_async_f32(mapa_u32(smem_u32(buf + r), peer), own, mapa_u32(smem_u32(bar), peer));
            if (warp == 0 && lane == 0) mbar_arrive_expect_tx(bar, BM * 4);  // one fp32 per row from the peer
            mbar_wait(bar, wait_parity);
            return own + *reinterpret_cast<volatile float*>(buf + r);
        };
        int64_t it = 0;
        for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
            const int as = (int)(it & 1);
            const uint32_t par2 = (uint32_t)((it >> 1) & 1);
            const int64_t row0 = t * BM, row = row0 + r;
            mbar_wait(acc_full + as, par2);
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * 256);
            float2 ssa = make_float2(0.f, 0.f), ssb = make_float2(0.f, 0.f);
            for (int ch = ch0; ch < ch1; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_lane + ch * 32, v);
                tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + ch * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = b4[j];
                    const float2 y01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y));
                    const float2 y23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w));
                    ssa = __ffma2_rn(y01, y01, ssa);
                    ssb = __ffma2_rn(y23, y23, ssb);
                }
            }
            red_s[half * BM + r] = (ssa.x + ssa.y) + (ssb.x + ssb.y);
            named_bar_sync(1, kEpiWarps * 32);
            const float nrm = sqrtf(exchange(0, as, par2, red_s[r] + red_s[BM + r]));
            const float scale = 1.0f / fmaxf(nrm, a.eps);  // F.normalize: x / max(||x||, eps)
            const float2 sc2 = make_float2(scale, scale);
            float2 qa = make_float2(0.f, 0.f);
            for (int ch = ch0; ch < ch1; ++ch) {
                uint32_t v[32];
                tmem_ld32(t_lane + ch * 32, v);
                tmem_ld_wait();
                const float4* b4 = reinterpret_cast<const float4*>(bias_s + ch * 32);
                uint32_t packed[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = b4[j];
                    const float2 e01 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y)), sc2);
                    const float2 e23 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w)), sc2);
                    const uint32_t p0 = pack_bf16(e01.x, e01.y), p1 = pack_bf16(e23.x, e23.y);
                    packed[2 * j] = p0;
                    packed[2 * j + 1] = p1;
                    const float2 r01 = make_float2(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u));
                    const float2 r23 = make_float2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
                    qa = __ffma2_rn(r01, r01, qa);
                    qa = __ffma2_rn(r23, r23, qa);
                }
                const int cp = (ch - ch0) & 1;
                uint8_t* sl = slab + (n_slab & 1) * kSlabBytes;
                if (cp == 0) {
                    if (lane == 0) tma_store_wait_read<1>();
                    __syncwarp();
                }
                uint8_t* srow = sl + lane * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c16 = (cp * 4 + k) ^ (lane & 7);
                    *reinterpret_cast<uint4*>(srow + c16 * 16) =
                        make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                }
                if (cp) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tm_out, sl, (int32_t)crank * n_half + (ch - 1) * 32, (int32_t)(row0 + quad * 32));
                        tma_store_commit();
                    }
                    ++n_slab;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);
            red_s[2 * BM + half * BM + r] = qa.x + qa.y;
            named_bar_sync(1, kEpiWarps * 32);
            const float q = exchange(1, as, par2, red_s[2 * BM + r] + red_s[3 * BM + r]);
            if (crank == 0 && half == 0 && row < a.rows) {
                if (a.rinv) a.rinv[row] = 1.0f / sqrtf(q);  // no epsilon: the scoring kernels' convention
                if (a.norm) a.norm[row] = nrm;
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit while its peer still multicasts into / signals it
    if (warp == kTmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------
#ifdef PB2_MEASURE
// Phased variant (round 2 experiment, measurement build only: pb2_debug_proj_variant(3)): the two column halves of the
// accumulator are produced ONE AFTER THE OTHER
// instead of interleaved -- MMA_A = x W[0:256]^T over all of K into TMEM columns [0, 256), then MMA_B = x W[256:512]^T
// into [256, 512) (the x k-tiles are streamed twice; they hit L2 the second time) -- so that each half has its own
// full / empty barrier and the epilogue overlaps the tensor pipe although the tile owns all of TMEM:
//     MMA_A(t) | MMA_B(t)          | MMA_A(t+1)        | MMA_B(t+1) ...
//              | pass1_A  pass1_B  | pass2_A  pass2_B  |
// pass 1 (sum of squares) of half A runs under MMA_B, and the next tile's MMA_A starts as soon as pass 2 (scale /
// round / store) has drained half A, i.e. under pass 2 of half B.  All eight epilogue warps work on the same half
// (two warps per TMEM lane quadrant, each a contiguous run of 64-column slabs).  A stage of the TMA ring is one x
// tile + this CTA's half of ONE W tile (32 KiB), five stages.  Correct (same tests as the product kernel) and NOT
// faster: 0.645 ms against 0.652 ms for 2^20 rows -- streaming x twice costs what the overlap buys (W already
// crosses L2 -> SM once per tile and pair: 768 KB per 134 MFLOP; with x twice 1 MB), so the product keeps the
// interleaved kernel above.
constexpr int kP_StageBytes = kXBytes + kWHalf;
constexpr int kP_Stages = 5;
constexpr int kP_Smem = kP_Stages * kP_StageBytes + kOutBytes + kRedBytes + 256;
static_assert(kP_Smem <= 227 * 1024, "phased encoder tail: shared memory");

__global__ void __launch_bounds__(kThreads, 1)
    project_normalize_phased_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                                    const __grid_constant__ CUtensorMap tm_out, const Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* out_stage = smem + kP_Stages * kP_StageBytes;
    float* red_s = reinterpret_cast<float*>(out_stage + kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red_s) + kRedBytes);
    uint64_t* full = bars;                 // [kP_Stages]
    uint64_t* empty = bars + kP_Stages;    // [kP_Stages]
    uint64_t* acc_full = empty + kP_Stages;   // [2]: halves A, B
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int64_t n_tiles = (a.rows + 2 * BM - 1) / (2 * BM);
    const int64_t unit0 = blockIdx.x / 2, n_units = gridDim.x / 2;
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kP_Stages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int h = 0; h < 2; ++h) {
            mbar_init(acc_full + h, 1);
            mbar_init(acc_empty + h, 2 * kEpiWarps);  // the leader's collects both CTAs' epilogues
        }
        fence_mbar_init();
    }
    if (warp == kTmaWarp) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int n_half[2] = {a.n_out > 256 ? 256 : a.n_out, a.n_out > 256 ? a.n_out - 256 : 0};  // columns of halves A, B

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units) {
                const int32_t xrow = (int32_t)((t * 2 + crank) * BM);
                for (int h = 0; h < 2; ++h) {
                    const int rows_w = n_half[h] / 2;  // this CTA's W rows of the half
                    if (rows_w == 0) continue;
                    const uint32_t tx = 2u * (uint32_t)(kXBytes + rows_w * BK * 2);  // both CTAs' bytes of a stage
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(empty + stage, phase ^ 1);
                        uint8_t* sx = smem + stage * kP_StageBytes;
                        uint8_t* sw = sx + kXBytes;
                        if (crank == 0) mbar_arrive_expect_tx(full + stage, tx);
                        const uint32_t lbar = mapa_u32(smem_u32(full + stage), 0);
                        tma_load_2d_pair(sx, &tm_x, lbar, kb * BK, xrow, h == 0 ? kEvictNormal : kEvictFirst);
                        for (int r0 = 0; r0 < rows_w; r0 += kWBox)
                            tma_load_2d_pair(sw + r0 * (BK * 2), &tm_w, lbar, kb * BK, h * 256 + (int32_t)crank * rows_w + r0, kEvictLast);
                        if (++stage == kP_Stages) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && crank == 0) {
            uint32_t idesc[2];
            idesc[0] = make_idesc(2 * BM, (uint32_t)n_half[0], kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            idesc[1] = n_half[1] > 0 ? make_idesc(2 * BM, (uint32_t)n_half[1], kFmtBF16, kFmtBF16, kMajorK, kMajorK) : 0u;
            const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            constexpr uint32_t kStageLo = kP_StageBytes >> 4, kWLo = kXBytes >> 4, kKLo = (UK * 2) >> 4;
            int stage = 0;
            uint32_t phase = 0, lo = lo0;
            int64_t it = 0;
            for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
                for (int h = 0; h < 2; ++h) {
                    if (n_half[h] == 0) continue;
                    mbar_wait(acc_empty + h, (uint32_t)(it & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < a.kblocks; ++kb) {
                        mbar_wait(full + stage, phase);
                        tc_fence_after();
#pragma unroll
                        for (int k = 0; k < BK / UK; ++k)
                            umma_f16_pair_lohi(tmem_base + h * 256, lo + k * kKLo, lo + kWLo + k * kKLo, desc_hi, desc_hi, idesc[h],
                                               (kb | k) != 0 ? 1u : 0u);
                        umma_commit_pair(empty + stage);
                        lo += kStageLo;
                        if (++stage == kP_Stages) {
                            stage = 0;
                            phase ^= 1;
                            lo = lo0;
                        }
                    }
                    umma_commit_pair(acc_full + h);
                }
            }
        }
    } else {
        // ===== epilogue: thread == row; in each half the quadrant's two warps take a contiguous run of 64-column slabs
        const int quad = warp & 3, sub = warp >> 2;
        uint8_t* slab = out_stage + warp * 2 * kSlabBytes;
        uint32_t n_slab = 0;
        int ch0[2], ch1[2];  // this warp's 32-column chunks of halves A and B (even counts: slabs are two chunks)
        for (int h = 0; h < 2; ++h) {
            const int pairs = n_half[h] / 64, first = (pairs + 1) / 2;
            ch0[h] = sub == 0 ? 0 : 2 * first;
            ch1[h] = sub == 0 ? 2 * first : 2 * pairs;
        }
        const int r = quad * 32 + lane;
        const bool has_bias = a.bias != nullptr;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        int64_t it = 0;
        for (int64_t t = unit0; t < n_tiles; t += n_units, ++it) {
            const int64_t row0 = (t * 2 + crank) * BM, row = row0 + r;
            const uint32_t par = (uint32_t)(it & 1);
            float2 ssa = make_float2(0.f, 0.f), ssb = make_float2(0.f, 0.f);
            uint32_t va[32], vb[32];
            // ---- pass 1: sum of squares, half A under MMA_B
            for (int h = 0; h < 2; ++h) {
                if (n_half[h] == 0) continue;
                mbar_wait(acc_full + h, par);
                tc_fence_after();
                const int c0 = ch0[h], c1 = ch1[h];
                auto pass1 = [&](const uint32_t (&v)[32], int ch) {
                    const float4* b4 = reinterpret_cast<const float4*>(a.bias + h * 256 + ch * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float2 y01 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y));
                        const float2 y23 = __fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w));
                        ssa = __ffma2_rn(y01, y01, ssa);
                        ssb = __ffma2_rn(y23, y23, ssb);
                    }
                };
                if (c0 < c1) tmem_ld32(t_lane + h * 256 + c0 * 32, va);
#pragma unroll 1
                for (int ch = c0; ch < c1; ch += 2) {
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(t_lane + h * 256 + (ch + 1) * 32, vb);
                    pass1(va, ch);
                    tmem_ld_wait();
                    __syncwarp();
                    if (ch + 2 < c1) tmem_ld32(t_lane + h * 256 + (ch + 2) * 32, va);
                    pass1(vb, ch + 1);
                }
            }
            red_s[sub * BM + r] = (ssa.x + ssa.y) + (ssb.x + ssb.y);
            named_bar_sync(1, kEpiWarps * 32);
            const float nrm = sqrtf(red_s[r] + red_s[BM + r]);
            const float scale = 1.0f / fmaxf(nrm, a.eps);  // F.normalize: x / max(||x||, eps)
            const float2 sc2 = make_float2(scale, scale);
            float2 qa = make_float2(0.f, 0.f);
            // ---- pass 2: scale / round / stage / store; half A is released to the next tile's MMA_A before half B is read
            for (int h = 0; h < 2; ++h) {
                if (n_half[h] == 0) continue;
                const int c0 = ch0[h], c1 = ch1[h];
                auto pass2 = [&](const uint32_t (&v)[32], int ch) {
                    const float4* b4 = reinterpret_cast<const float4*>(a.bias + h * 256 + ch * 32);
                    uint32_t packed[16];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = has_bias ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        const float2 e01 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), make_float2(b.x, b.y)), sc2);
                        const float2 e23 = __fmul2_rn(__fadd2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), make_float2(b.z, b.w)), sc2);
                        const uint32_t p0 = pack_bf16(e01.x, e01.y), p1 = pack_bf16(e23.x, e23.y);
                        packed[2 * j] = p0;
                        packed[2 * j + 1] = p1;
                        const float2 r01 = make_float2(__uint_as_float(p0 << 16), __uint_as_float(p0 & 0xffff0000u));
                        const float2 r23 = make_float2(__uint_as_float(p1 << 16), __uint_as_float(p1 & 0xffff0000u));
                        qa = __ffma2_rn(r01, r01, qa);
                        qa = __ffma2_rn(r23, r23, qa);
                    }
                    const int cp = (ch - c0) & 1;
                    uint8_t* sl = slab + (n_slab & 1) * kSlabBytes;
                    if (cp == 0) {  // the slab about to be rewritten must have been read by its TMA store
                        if (lane == 0) tma_store_wait_read<1>();
                        __syncwarp();
                    }
                    uint8_t* srow = sl + lane * 128;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int c16 = (cp * 4 + k) ^ (lane & 7);  // 128-byte swizzle
                        *reinterpret_cast<uint4*>(srow + c16 * 16) =
                            make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
                    }
                    if (cp) {
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tm_out, sl, h * 256 + (ch - 1) * 32, (int32_t)(row0 + quad * 32));
                            tma_store_commit();
                        }
                        ++n_slab;
                    }
                };
                if (c0 < c1) tmem_ld32(t_lane + h * 256 + c0 * 32, va);
#pragma unroll 1
                for (int ch = c0; ch < c1; ch += 2) {
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(t_lane + h * 256 + (ch + 1) * 32, vb);
                    pass2(va, ch);
                    tmem_ld_wait();
                    __syncwarp();
                    if (ch + 2 < c1) tmem_ld32(t_lane + h * 256 + (ch + 2) * 32, va);
                    pass2(vb, ch + 1);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty + h), 0));  // this half may be overwritten
            }
            red_s[2 * BM + sub * BM + r] = qa.x + qa.y;
            named_bar_sync(1, kEpiWarps * 32);
            if (sub == 0 && row < a.rows) {
                if (a.rinv) a.rinv[row] = 1.0f / sqrtf(red_s[2 * BM + r] + red_s[3 * BM + r]);  // no epsilon: the scoring kernels' convention
                if (a.norm) a.norm[row] = nrm;
            }
        }
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still signals it
    if (warp == kTmaWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

#endif  // PB2_MEASURE

// pb2_debug_proj_variant (measurement build): 0 = the product's choice (the resident-x phased kernel when n_in <= 512 and
// n_out % 256 == 0, else the interleaved kernel), 1 = the interleaved kernel always, 2 = column split (n_out % 128 == 0),
// 3 = phased halves with x streamed twice, 4 = interleaved with sixteen epilogue warps, 5 = resident x with 128-column
// phases.  2^20 rows 512 -> 512, variants alternating on one board (tools/ab_tail.py): interleaved 0.62 ms alone /
// 0.75 ms sustained, resident-256 0.56 / 0.61 ms, resident-128 0.74 / 0.88, sixteen warps 0.63 / 0.70, column split 0.68,
// phased halves 0.645.
PB2_KNOB g_variant = 0;

}  // namespace proj
}  // namespace pb2

using namespace pb2;

#ifdef PB2_MEASURE
extern "C" int pb2_debug_proj_variant(int v) {
    proj::g_variant = v;
    return PB2_OK;
}
#endif

extern "C" int pb2_project_normalize(const void* x, const void* w, const float* bias, int64_t rows, int n_in, int n_out,
                                     int64_t ldx, int64_t ldw, float eps, void* out, int64_t ld_out, float* rinv,
                                     float* norm, void* stream) {
    if (rows <= 0) return PB2_OK;
    if (!x || !w || !out) return set_error(PB2_ERR_ARG, "project_normalize: null");
    if (n_in <= 0 || n_in % 64 != 0) return set_error(PB2_ERR_ARG, "project_normalize: n_in must be a positive multiple of 64");
    if (n_out <= 0 || n_out % 64 != 0 || n_out > proj::kMaxOut)
        return set_error(PB2_ERR_ARG, "project_normalize: n_out must be a multiple of 64, at most 512");
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return set_error(PB2_ERR_ARG, "project_normalize: bias must be 16-byte aligned");
    if (ld_out % 8 != 0 || ld_out < n_out) return set_error(PB2_ERR_ARG, "project_normalize: ld_out %% 8 == 0, >= n_out");
    CUtensorMap tx, tw, to;
    int rc = make_tmap_2d(&tx, x, 2, (uint64_t)rows, (uint64_t)n_in, (uint64_t)ldx * 2, proj::BM, proj::BK);
    if (rc) return rc;
    rc = make_tmap_2d(&tw, w, 2, (uint64_t)n_out, (uint64_t)n_in, (uint64_t)ldw * 2, proj::kWBox, proj::BK);
    if (rc) return rc;
    rc = make_tmap_2d(&to, out, 2, (uint64_t)rows, (uint64_t)n_out, (uint64_t)ld_out * 2, 32, 64);
    if (rc) return rc;
    proj::Args a;
    a.rows = rows;
    a.n_out = n_out;
    a.kblocks = n_in / proj::BK;
    a.tx_bytes = (uint32_t)(2 * proj::kXBytes + n_out * proj::BK * 2);  // both CTAs: two x tiles + all W rows once
    a.bias = bias;
    a.eps = eps;
    a.rinv = rinv;
    a.norm = norm;
    if (n_out % 128 == 0 && proj::g_variant == 2) {  // column-split pairs: two TMEM stages, epilogue overlaps the MMA
        CUtensorMap tx2;
        rc = make_tmap_2d(&tx2, x, 2, (uint64_t)rows, (uint64_t)n_in, (uint64_t)ldx * 2, 64, proj::BK);
        if (rc) return rc;
        static PerDeviceOnce configured2;
        rc = ensure_dynamic_smem(configured2, proj::project_normalize_split_kernel, proj::kS_Smem, "project_normalize");
        if (rc) return rc;
        const int64_t tiles = (rows + proj::BM - 1) / proj::BM;
        const int grid2 = 2 * (int)std::min<int64_t>(tiles, sm_count() / 2);
        rc = check_cuda(launch_ex(proj::project_normalize_split_kernel, (unsigned)grid2, (unsigned)proj::kThreads,
                                  (size_t)proj::kS_Smem, (cudaStream_t)stream, 2, tx2, tw, to, a),
                        "project_normalize launch");
        if (rc) return rc;
        return check_launch("project_normalize");
    }
    const int64_t n_tiles = (rows + 2 * proj::BM - 1) / (2 * proj::BM);
    const int grid = 2 * (int)std::min<int64_t>(n_tiles, sm_count() / 2);
#ifdef PB2_MEASURE
    if (proj::g_variant == 3) {  // phased halves: the epilogue overlaps the tensor pipe (experiment)
        static PerDeviceOnce configured3;
        rc = ensure_dynamic_smem(configured3, proj::project_normalize_phased_kernel, proj::kP_Smem, "project_normalize");
        if (rc) return rc;
        rc = check_cuda(launch_ex(proj::project_normalize_phased_kernel, (unsigned)grid, (unsigned)proj::kThreads,
                                  (size_t)proj::kP_Smem, (cudaStream_t)stream, 2, tx, tw, to, a),
                        "project_normalize launch");
        if (rc) return rc;
        return check_launch("project_normalize");
    }
#endif
    // n_in <= 512 and whole phases: x resident in shared memory, accumulator produced in phases (the product for the
    // reference's Linear(512, 512) tails); everything else takes the interleaved kernel
    const int pw = proj::g_variant == 5 ? 128 : 256;
    if ((proj::g_variant == 0 || proj::g_variant == 5) && a.kblocks <= proj::kR_MaxKb && n_out % pw == 0) {
        CUtensorMap twp;
        rc = make_tmap_2d(&twp, w, 2, (uint64_t)n_out, (uint64_t)n_in, (uint64_t)ldw * 2, pw / 2, proj::BK);
        if (rc) return rc;
        static PerDeviceOnce configured_r256;
        if (pw == 256) {
            rc = ensure_dynamic_smem(configured_r256, proj::project_normalize_resident_kernel<256>, proj::kR_Smem, "project_normalize");
            if (rc) return rc;
            rc = check_cuda(launch_ex(proj::project_normalize_resident_kernel<256>, (unsigned)grid, (unsigned)proj::kR_Threads,
                                      (size_t)proj::kR_Smem, (cudaStream_t)stream, 2, tx, twp, to, a),
                            "project_normalize launch");
        }
#ifdef PB2_MEASURE
        else {
            static PerDeviceOnce configured_r128;
            rc = ensure_dynamic_smem(configured_r128, proj::project_normalize_resident_kernel<128>, proj::kR_Smem, "project_normalize");
            if (rc) return rc;
            rc = check_cuda(launch_ex(proj::project_normalize_resident_kernel<128>, (unsigned)grid, (unsigned)proj::kR_Threads,
                                      (size_t)proj::kR_Smem, (cudaStream_t)stream, 2, tx, twp, to, a),
                            "project_normalize launch");
        }
#endif
        if (rc) return rc;
        return check_launch("project_normalize");
    }
#ifdef PB2_MEASURE
    if (proj::g_variant == 4) {  // sixteen epilogue warps (four per TMEM lane quadrant and scheduler)
        static PerDeviceOnce configured4;
        rc = ensure_dynamic_smem(configured4, proj::project_normalize_kernel<4>, proj::kSmem, "project_normalize");
        if (rc) return rc;
        rc = check_cuda(launch_ex(proj::project_normalize_kernel<4>, (unsigned)grid, (unsigned)proj::Cfg<4>::kThr,
                                  (size_t)proj::kSmem, (cudaStream_t)stream, 2, tx, tw, to, a),
                        "project_normalize launch");
        if (rc) return rc;
        return check_launch("project_normalize");
    }
#endif
    static PerDeviceOnce configured;
    rc = ensure_dynamic_smem(configured, proj::project_normalize_kernel<2>, proj::kSmem, "project_normalize");
    if (rc) return rc;
    rc = check_cuda(launch_ex(proj::project_normalize_kernel<2>, (unsigned)grid, (unsigned)proj::Cfg<2>::kThr,
                              (size_t)proj::kSmem, (cudaStream_t)stream, 2, tx, tw, to, a),
                    "project_normalize launch");
    if (rc) return rc;
    return check_launch("project_normalize");
}
