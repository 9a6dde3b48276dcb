// Backward GEMMs of the contrastive losses on the tensor cores:
//     out[M, dim] (=|+=) alpha * op(G) * Z        G fp16 gradient matrix, Z fp16 (normalised) embeddings
// i.e. the autograd backward of torch.matmul in pig/util.py:13 / pig/loss.py:19
// (dX = G Y, dY = G^T X).  G comes from the fused forward epilogue (sim.cu).  tcgen05 kind::f16 cannot mix
// an fp16 A with a bf16 B (illegal instruction, measured), so Z is an fp16 copy of the embeddings
// (pb2_rows_scale_f16); both operand formats are runtime fields of the instruction descriptor.
//
// Operand layouts for tcgen05.mma (kind::f16, fp32 accumulate in TMEM):
//   A = G   (transpose == 0): K-major, one TMA box [128 rows x 64 k] per stage.
//   A = G^T (transpose != 0): MN-major, two TMA boxes [64 k x 64 m] per stage.
//   B = Z^T always MN-major (Z is [K, dim] row-major): BN/64 TMA boxes [64 k x 64 n] per stage.
// Same warp-specialised pipeline as sim.cu: 8 epilogue warps, MMA warp, TMA warp, 2 TMEM stages.
//
// Work decomposition.  A gradient-matrix block of 32768 x 32768 gives 256 x 2 output tiles of 128 x 256 with
// 512 k-blocks each: 512 long tiles on 148 SMs = 3.46 waves, i.e. 13 % of the machine idles in the last
// wave.  With a caller workspace the kernel therefore runs STREAM-K: the (row block, k-block) units are cut
// into equal contiguous ranges, one per GROUP of n_cb CTAs (the CTAs of a group walk the same range for
// the n_cb column tiles, so a G tile is still fetched from HBM once and hit in L2 by its neighbours).  A
// row block cut between two groups is finished deterministically: the group holding its tail (always that
// group's FIRST segment) parks the raw accumulators in the workspace and raises a flag; the group holding
// its head (its LAST segment, reached ~a whole range later) adds them in its epilogue.  No atomics on the
// output, fixed summation order, bit-reproducible.
#include "common.cuh"
#include "fold.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {
namespace gg {

constexpr int BM = 128, BK = 64, UK = 16;
// epilogue warps 0-7, MMA issuer 8, TMA producer + TMEM alloc 9: the warp arbiter prefers the highest
// eligible warp id, so the tensor pipeline's warps get the highest ids (see sim.cu)
constexpr int kEpiWarp0 = 0, kEpiWarps = 8, kMmaWarp = 8, kTmaWarp = 9;
constexpr int kThreads = (kEpiWarps + 2) * 32;
constexpr int kBoxBytes = 64 * 64 * 2;  // one [64 x 64] 16-bit box

struct Args {
    int64_t m, k;  // output rows, contraction length
    int n_rb, n_cb;
    int64_t n_tiles;
    int kblocks;
    float alpha;
    int accumulate;
    float* out;
    int64_t ld_out;
    // descriptor strides (bytes); runtime so the self-test can probe alternatives
    uint32_t mn_lbo, mn_sbo, mn_kstep;
    uint32_t idesc;
    // stream-K (n_groups > 0): group g = blockIdx / n_cb owns units [total*g/n_groups, total*(g+1)/n_groups)
    int n_groups;
    uint32_t* flags;  // [grid] arrival counters, zero between launches
    float4* parts;    // [grid][BN/4][128] raw accumulators of a CTA's tail segment
    int z_lo_col;     // kind::i8 variant: first column of the low plane inside the two-plane embedding operand
};

template <int BN, int kCtas>
struct Smem {
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = (BN / kCtas / 64) * kBoxBytes;  // a CTA of a pair stages half of B's columns
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (200 * 1024) / kStageBytes > 8 ? 8 : (200 * 1024) / kStageBytes;
    static constexpr int kTileBytes = kStages * kStageBytes;
    static constexpr int kTotal = 1024 + kTileBytes + 256;
};

// Calls f(rb, cb, kb0, kb1) for every segment of this CTA (or CTA pair), in the same order for all warp
// roles.  rb counts row blocks of BM * kCtas rows.
template <int kCtas, class F>
__device__ __forceinline__ void for_each_segment(const Args& a, int block, int nblocks, F&& f) {
    const int unit = block / kCtas, n_units = nblocks / kCtas;
    if (a.n_groups > 0) {
        const int g = unit / a.n_cb, cb = unit % a.n_cb;
        if (g >= a.n_groups) return;
        const int64_t total = (int64_t)a.n_rb * a.kblocks;
        int64_t u = total * g / a.n_groups;
        const int64_t u1 = total * (g + 1) / a.n_groups;
        while (u < u1) {
            const int rb = (int)(u / a.kblocks), kb0 = (int)(u % a.kblocks);
            const int kb1 = (int)min((int64_t)a.kblocks, kb0 + (u1 - u));
            f(rb, cb, kb0, kb1);
            u += kb1 - kb0;
        }
    } else {
        for (int64_t t = unit; t < a.n_tiles; t += n_units) f((int)(t / a.n_cb), (int)(t % a.n_cb), 0, a.kblocks);
    }
}

// kCtas == 2: CTA pairs (cluster of 2) drive one tcgen05.mma.cta_group::2 of M = 256, N = BN per k-step;
// each CTA stages its own 128 rows of A and BN/2 columns of B, accumulates its 128 x BN in its own TMEM
// and runs its own epilogue.  See common.cuh for the protocol.
// `block` of `nblocks`: this CTA's index within ITS product (the dual launch runs two products in one grid).
template <bool kTranspose, int BN, int kCtas>
__device__ __forceinline__ void gg_body(const CUtensorMap& tm_g, const CUtensorMap& tm_z, const Args& a, const int block,
                                        const int nblocks) {
    using L = Smem<BN, kCtas>;
    constexpr int BNL = BN / kCtas;  // B columns staged by this CTA
    // BN == 512 (pairs only): the 128 x 512 fp32 accumulator is all of TMEM, one stage -- the epilogue is not
    // overlapped, which costs ~1 % at the contraction lengths this variant is chosen for
    constexpr int kAccStages = BN == 512 ? 1 : 2;
    constexpr int kUmmaN = BN > 256 ? 256 : BN;  // N of one tcgen05.mma; BN == 512 issues two per k-step
    static_assert(BN <= 256 || kCtas == 2, "512-wide tiles need CTA pairs");
    // 128-byte-swizzled TMA/UMMA tiles need 1024-byte alignment; the kernel has no static shared
    // memory, so the dynamic segment starts at the (aligned) base of the CTA's shared window.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + L::kStages;
    uint64_t* acc_full = bars + 2 * L::kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = kCtas == 2 ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_g);
        tma_prefetch_desc(&tm_z);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc_full + s, 1);
            mbar_init(acc_empty + s, kEpiWarps * kCtas);  // the leader's collects both CTAs' epilogues
        }
        fence_mbar_init();
    }
    if (warp == kTmaWarp) {
        if (kCtas == 2) tmem_alloc_pair(tmem_slot, kAccStages * BN);
        else tmem_alloc(tmem_slot, kAccStages * BN);
    }
    pdl_launch_dependents();
    tc_fence_before();
    if (kCtas == 2) cluster_sync_all();  // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // the prologue above overlapped the previous kernel's tail; global memory from here on

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for_each_segment<kCtas>(a, block, nblocks, [&](int rb, int cb, int kb0, int kb1) {
                const int row0 = (rb * kCtas + (int)crank) * BM;
                // this CTA's B columns: chunk q of 64 -> MMA q / (kUmmaN / kCtas / 64), half `crank` of its N
                constexpr int kChunksPerMma = kUmmaN / kCtas / 64;
                auto bcol = [&](int q) {
                    return cb * BN + (q / kChunksPerMma) * kUmmaN + (int)crank * (kUmmaN / kCtas) + (q % kChunksPerMma) * 64;
                };
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sa = smem + stage * L::kStageBytes;
                    uint8_t* sb = sa + L::kABytes;
                    if (kCtas == 1) {
                        mbar_arrive_expect_tx(full + stage, L::kStageBytes);
                        if (!kTranspose) {
                            tma_load_2d(sa, &tm_g, full + stage, kb * BK, row0, kEvictFirst);
                        } else {
                            tma_load_2d(sa, &tm_g, full + stage, row0, kb * BK, kEvictFirst);
                            tma_load_2d(sa + kBoxBytes, &tm_g, full + stage, row0 + 64, kb * BK, kEvictFirst);
                        }
#pragma unroll
                        for (int cchunk = 0; cchunk < BNL / 64; ++cchunk)
                            tma_load_2d(sb + cchunk * kBoxBytes, &tm_z, full + stage, bcol(cchunk), kb * BK, kEvictLast);
                    } else {
                        // both CTAs' bytes are counted on the leader's barrier
                        if (crank == 0) mbar_arrive_expect_tx(full + stage, L::kStageBytes * kCtas);
                        const uint32_t lbar = mapa_u32(smem_u32(full + stage), 0);
                        if (!kTranspose) {
                            tma_load_2d_pair(sa, &tm_g, lbar, kb * BK, row0, kEvictFirst);
                        } else {
                            tma_load_2d_pair(sa, &tm_g, lbar, row0, kb * BK, kEvictFirst);
                            tma_load_2d_pair(sa + kBoxBytes, &tm_g, lbar, row0 + 64, kb * BK, kEvictFirst);
                        }
#pragma unroll
                        for (int cchunk = 0; cchunk < BNL / 64; ++cchunk)
                            tma_load_2d_pair(sb + cchunk * kBoxBytes, &tm_z, lbar, bcol(cchunk), kb * BK, kEvictLast);
                    }
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            });
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && crank == 0) {
            const uint32_t idesc = a.idesc;
            // running descriptor low words (start address >> 4 | LBO); the high words are loop invariants
            const uint64_t dk = make_smem_desc(smem_u32(smem), 16, 1024);               // K-major G tile
            const uint64_t dm = make_smem_desc(smem_u32(smem), a.mn_lbo, a.mn_sbo);     // MN-major tiles
            const uint32_t a_hi = (uint32_t)((kTranspose ? dm : dk) >> 32), b_hi = (uint32_t)(dm >> 32);
            const uint32_t a_lo0 = (uint32_t)(kTranspose ? dm : dk), b_lo0 = (uint32_t)dm + (L::kABytes >> 4);
            const uint32_t a_kstep = kTranspose ? (a.mn_kstep >> 4) : ((UK * 2) >> 4), b_kstep = a.mn_kstep >> 4;
            constexpr uint32_t kStageLo = L::kStageBytes >> 4;
            constexpr uint32_t kMmaBLo = ((kUmmaN / kCtas / 64) * kBoxBytes) >> 4;  // B chunks of MMA j follow those of j-1
            int stage = 0;
            uint32_t phase = 0, a_lo = a_lo0, b_lo = b_lo0;
            int64_t it = 0;
            for_each_segment<kCtas>(a, block, nblocks, [&](int, int, int kb0, int kb1) {
                const int as = (int)(it % kAccStages);
                mbar_wait(acc_empty + as, (uint32_t)((it / kAccStages) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
#pragma unroll
                        for (int j = 0; j < BN / kUmmaN; ++j) {
                            if (kCtas == 2)
                                umma_f16_pair_lohi(d_tmem + j * kUmmaN, a_lo + k * a_kstep, b_lo + j * kMmaBLo + k * b_kstep, a_hi,
                                                   b_hi, idesc, acc);
                            else
                                umma_f16_pair_lohi_1(d_tmem + j * kUmmaN, a_lo + k * a_kstep, b_lo + j * kMmaBLo + k * b_kstep,
                                                     a_hi, b_hi, idesc, acc);
                        }
                    }
                    if (kCtas == 2) umma_commit_pair(empty + stage);
                    else umma_commit(empty + stage);
                    a_lo += kStageLo;
                    b_lo += kStageLo;
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                        a_lo = a_lo0;
                        b_lo = b_lo0;
                    }
                }
                if (kCtas == 2) umma_commit_pair(acc_full + as);
                else umma_commit(acc_full + as);
                ++it;
            });
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - kEpiWarp0) >> 2;
        constexpr int kChunksPerHalf = BN / 64;
        int64_t it = 0;
        for_each_segment<kCtas>(a, block, nblocks, [&](int rb, int cb, int kb0, int kb1) {
            const int as = (int)(it % kAccStages);
            const int64_t row = ((int64_t)rb * kCtas + crank) * BM + quad * 32 + lane;
            // stream-K: a tail segment parks its accumulators; a head segment folds its neighbour's in
            const bool park = kb0 > 0, fold = kb1 < a.kblocks;
            const uint32_t peer = (uint32_t)block + (uint32_t)(a.n_cb * kCtas);  // same tile slot, next group
            if (fold) {
                if (lane == 0) {
                    // The peer is a CTA of this grid that may not be resident yet when another kernel (an NCCL
                    // reduction overlapped on a second stream, another tenant) holds SMs: being late is legal, so
                    // this wait backs off instead of trapping after the couple of seconds a lost mbarrier arrive
                    // gets.  Only a wait of minutes -- a protocol bug -- ends the kernel rather than hang the GPU.
                    const long long t0 = clock64();
                    uint32_t ns = 100;
                    while (ld_acquire_u32(a.flags + peer) < (uint32_t)kEpiWarps) {
                        __nanosleep(ns);
                        if (ns < 8000) ns <<= 1;
                        if (clock64() - t0 > 64 * PB2_WAIT_TIMEOUT_CYCLES) {
                            printf("pb2: grad_gemm stream-K flag wait gave up after minutes (block %d)\n", block);
                            __trap();
                        }
                    }
                }
                __syncwarp();
            }
            mbar_wait(acc_full + as, (uint32_t)((it / kAccStages) & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int ch = 0; ch < kChunksPerHalf; ++ch) {
                const int cbase = (half * kChunksPerHalf + ch) * 32;
                uint32_t v[32];
                tmem_ld32(t_lane + cbase, v);
                tmem_ld_wait();
                if (park) {  // [BN/4][128] float4: consecutive lanes (rows) write consecutive 16 bytes
                    float4* dst = a.parts + ((size_t)block * (BN / 4) + cbase / 4) * BM + quad * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j * BM] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                  __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    continue;
                }
                if (fold) {
                    const float4* src = a.parts + ((size_t)peer * (BN / 4) + cbase / 4) * BM + quad * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 o = __ldcg(src + j * BM);
                        v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + o.x);
                        v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + o.y);
                        v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + o.z);
                        v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + o.w);
                    }
                }
                if (row < a.m) {
                    float* dst = a.out + row * a.ld_out + (int64_t)cb * BN + cbase;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(__uint_as_float(v[j]) * a.alpha, __uint_as_float(v[j + 1]) * a.alpha,
                                               __uint_as_float(v[j + 2]) * a.alpha, __uint_as_float(v[j + 3]) * a.alpha);
                        if (a.accumulate) {
                            const float4 old = *reinterpret_cast<const float4*>(dst + j);
                            o.x += old.x;
                            o.y += old.y;
                            o.z += old.z;
                            o.w += old.w;
                        }
                        *reinterpret_cast<float4*>(dst + j) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (kCtas == 2) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty + as), 0));
                else mbar_arrive(acc_empty + as);
            }
            if (park) {  // publish: every lane's stores, then one release-increment per warp
                __threadfence();
                __syncwarp();
                if (lane == 0) red_release_add_u32(a.flags + block, 1u);
            }
            if (fold) {  // all epilogue warps are past their reads of the peer's slot: re-arm its flag
                named_bar_sync(1, kEpiWarps * 32);
                if (warp == kEpiWarp0 && lane == 0) a.flags[peer] = 0u;
            }
            ++it;
        });
    }
    tc_fence_before();
    if (kCtas == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still signals it
    else __syncthreads();
    if (warp == kTmaWarp) {
        tc_fence_after();
        if (kCtas == 2) tmem_dealloc_pair(tmem_base, kAccStages * BN);
        else tmem_dealloc(tmem_base, kAccStages * BN);
    }
}

template <bool kTranspose, int BN, int kCtas>
__global__ void __launch_bounds__(kThreads, 1)
    grad_gemm_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z,
                     const Args a) {
    gg_body<kTranspose, BN, kCtas>(tm_g, tm_z, a, (int)blockIdx.x, (int)gridDim.x);
}

// Both backward products of one gradient-matrix block in ONE launch: CTAs [0, n0) compute out0 = G Z0, the
// rest out1 = G^T Z1.  A batch-1k training step is launch bound (each of these GEMMs is ~64 short tiles), and
// the two products are independent, so they share a grid instead of queueing behind each other.
// fold.loss_out != nullptr (pb2_hinge_forward): one more CTA at the end of the grid folds the step's scalar loss from
// what the similarity pass left behind, beside the products instead of in a launch of its own.
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
    grad_gemm_dual_kernel(const __grid_constant__ CUtensorMap tm_g0, const __grid_constant__ CUtensorMap tm_z0,
                          const __grid_constant__ CUtensorMap tm_g1, const __grid_constant__ CUtensorMap tm_z1, const Args a0,
                          const Args a1, const int n0, const int n1, const HingeFold fold) {
    if ((int)blockIdx.x < n0) {
        gg_body<false, BN, 1>(tm_g0, tm_z0, a0, (int)blockIdx.x, n0);
    } else if ((int)blockIdx.x < n0 + n1) {
        gg_body<true, BN, 1>(tm_g1, tm_z1, a1, (int)blockIdx.x - n0, n1);
    } else {
        extern __shared__ __align__(1024) uint8_t smem[];
        pdl_launch_dependents();
        pdl_wait();
        if (threadIdx.x < 256)
            hinge_loss_fold(fold, (int)threadIdx.x, reinterpret_cast<double*>(smem), reinterpret_cast<int*>(smem + 64), 1);
    }
}


// ----------------------------------------------------------------------------------------------------------
// kind::i8 variant: the hinge gradient matrix holds exactly {0, 1, 2}, so it travels as ONE byte per entry (u8,
// written by sim.cu's HingePolicyT<.., true>) -- half the HBM bytes on the way out of the similarity pass and on
// both ways into the gradient products.  The embedding operand is q = round(xhat * 32512) = 256 hi + lo as two
// 8-bit planes [K, 2 dim] = [hi (s8) | lo (u8)] (pb2_rows_quant_i8: 16 bits, one scale for the tensor); the two
// products G hi and G lo accumulate EXACTLY in s32 (|G hi| <= 2 * 127 * K < 2^31 for any block), and the epilogue
// joins them: out (+)= alpha / 32512 * (256 acc_hi + acc_lo).  tcgen05.mma.kind::i8 issues 2.3x the MACs per second of
// kind::f16 from the same operand bytes (tools/ubench/i8mma.cu), so two planes cost no more tensor time than the fp16
// product they replace.
//
// CTA pairs only (cta_group::2, M = 256); a pair tile is 256 rows x 256 output columns, its two s32 accumulators
// (hi plane at TMEM columns [0, 256), lo plane at [256, 512)) fill the 128 x 512 TMEM of each CTA: one accumulator
// stage, like the 512-wide fp16 tile.  A stage of the TMA ring is 128 k-bytes deep:
//   A = G   (transpose == 0): K-major   [128 rows x 128 k] u8          16 KiB  (+32 B per K = 32 instruction)
//   A = G^T (transpose != 0): MN-major  [128 k-rows x 128 m] u8        16 KiB  (+4096 B per instruction, SBO 1024)
//   B planes, always MN-major: [128 k-rows x 128 n] s8 and u8           2 x 16 KiB (this CTA's half of the 256 columns)
// Stream-K, the parked accumulators and the fold are those of gg_body; the fold adds integers, so a cut row block is
// bit-identical to an uncut one.
namespace i8 {
constexpr int kBK8 = 128, kUK8 = 32, BN_OUT = 256;
constexpr int kTileA = BM * kBK8, kTileB = (BN_OUT / 2) * kBK8;
constexpr int kStageBytes = kTileA + 2 * kTileB;
constexpr int kStages = 4;
constexpr int kTileBytes = kStages * kStageBytes;
constexpr int kTotal = 1024 + kTileBytes + 256;
constexpr int kTmemCols = 512;
constexpr float kScale = 32512.0f;  // 127 * 256: q = round(xhat * kScale) in [-32512, 32512] = 256 * [-127, 127] + [0, 255]
}  // namespace i8

template <bool kTranspose>
__global__ void __launch_bounds__(kThreads, 1)
    grad_gemm_i8_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_z, const Args a) {
    using namespace i8;
    constexpr int kCtas = 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + kStages;
    uint64_t* acc_full = bars + 2 * kStages;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

    const int block = (int)blockIdx.x, nblocks = (int)gridDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();  // 0 = leader (issues the MMAs)
    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_g);
        tma_prefetch_desc(&tm_z);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kEpiWarps * kCtas);  // the leader's collects both CTAs' epilogues
        fence_mbar_init();
    }
    if (warp == kTmaWarp) tmem_alloc_pair(tmem_slot, kTmemCols);
    pdl_launch_dependents();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == kTmaWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for_each_segment<kCtas>(a, block, nblocks, [&](int rb, int cb, int kb0, int kb1) {
                const int row0 = (rb * kCtas + (int)crank) * BM;
                const int ncol = cb * BN_OUT + (int)crank * (BN_OUT / 2);  // this CTA's half of the tile's output columns
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sa = smem + stage * kStageBytes;
                    uint8_t* sb_hi = sa + kTileA;
                    uint8_t* sb_lo = sb_hi + kTileB;
                    if (crank == 0) mbar_arrive_expect_tx(full + stage, kStageBytes * kCtas);
                    const uint32_t lbar = mapa_u32(smem_u32(full + stage), 0);
                    if (!kTranspose) tma_load_2d_pair(sa, &tm_g, lbar, kb * kBK8, row0, kEvictFirst);
                    else tma_load_2d_pair(sa, &tm_g, lbar, row0, kb * kBK8, kEvictFirst);
                    tma_load_2d_pair(sb_hi, &tm_z, lbar, ncol, kb * kBK8, kEvictLast);
                    tma_load_2d_pair(sb_lo, &tm_z, lbar, a.z_lo_col + ncol, kb * kBK8, kEvictLast);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            });
        }
    } else if (warp == kMmaWarp) {
        if (lane == 0 && crank == 0) {
            constexpr uint32_t kAMajor = kTranspose ? kMajorMN : kMajorK;
            constexpr uint32_t idesc_hi = make_idesc_i8(BM * kCtas, BN_OUT, kFmtU8, kFmtS8, kAMajor, kMajorMN);
            constexpr uint32_t idesc_lo = make_idesc_i8(BM * kCtas, BN_OUT, kFmtU8, kFmtU8, kAMajor, kMajorMN);
            // K-major G tile: 128-byte rows, 8-row atoms (SBO 1024).  MN-major tiles: 128 bytes along M / N per k-row,
            // 8 k-rows per 1024-byte atom (SBO 1024); a tile is ONE 128-element chunk along M / N, so LBO (the distance
            // between such chunks) is never used -- set to the tile size.
            const uint64_t dk = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint64_t dm = make_smem_desc(smem_u32(smem), 16384, 1024);
            const uint32_t a_hi = (uint32_t)((kTranspose ? dm : dk) >> 32), b_hi = (uint32_t)(dm >> 32);
            const uint32_t a_lo0 = (uint32_t)(kTranspose ? dm : dk), b_lo0 = (uint32_t)dm + (kTileA >> 4);
            constexpr uint32_t a_kstep = kTranspose ? (kUK8 * 128) >> 4 : kUK8 >> 4;  // 32 k-rows of 128 B, or 32 bytes along a row
            constexpr uint32_t b_kstep = (kUK8 * 128) >> 4;
            constexpr uint32_t kStageLo = kStageBytes >> 4, kPlaneLo = kTileB >> 4;
            int stage = 0;
            uint32_t phase = 0, a_lo = a_lo0, b_lo = b_lo0;
            int64_t it = 0;
            for_each_segment<kCtas>(a, block, nblocks, [&](int, int, int kb0, int kb1) {
                mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
                tc_fence_after();
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kBK8 / kUK8; ++k) {
                        const uint32_t acc = (kb != kb0 || k != 0) ? 1u : 0u;
                        umma_i8_pair_lohi(tmem_base, a_lo + k * a_kstep, b_lo + k * b_kstep, a_hi, b_hi, idesc_hi, acc);
                        umma_i8_pair_lohi(tmem_base + BN_OUT, a_lo + k * a_kstep, b_lo + kPlaneLo + k * b_kstep, a_hi, b_hi,
                                          idesc_lo, acc);
                    }
                    umma_commit_pair(empty + stage);
                    a_lo += kStageLo;
                    b_lo += kStageLo;
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1;
                        a_lo = a_lo0;
                        b_lo = b_lo0;
                    }
                }
                umma_commit_pair(acc_full);
                ++it;
            });
        }
    } else {
        const int quad = warp & 3;
        const int half = (warp - kEpiWarp0) >> 2;  // which 128 of the tile's 256 output columns
        int64_t it = 0;
        for_each_segment<kCtas>(a, block, nblocks, [&](int rb, int cb, int kb0, int kb1) {
            const int64_t row = ((int64_t)rb * kCtas + crank) * BM + quad * 32 + lane;
            const bool park = kb0 > 0, fold = kb1 < a.kblocks;
            const uint32_t peer = (uint32_t)block + (uint32_t)(a.n_cb * kCtas);  // same tile slot, next group
            if (fold) {
                if (lane == 0) {
                    const long long t0 = clock64();
                    uint32_t ns = 100;
                    while (ld_acquire_u32(a.flags + peer) < (uint32_t)kEpiWarps) {  // see gg_body: late is legal
                        __nanosleep(ns);
                        if (ns < 8000) ns <<= 1;
                        if (clock64() - t0 > 64 * PB2_WAIT_TIMEOUT_CYCLES) {
                            printf("pb2: grad_gemm_i8 stream-K flag wait gave up after minutes (block %d)\n", block);
                            __trap();
                        }
                    }
                }
                __syncwarp();
            }
            mbar_wait(acc_full, (uint32_t)(it & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
                const int cbase = half * 128 + ch * 32;  // output column of the chunk inside the tile = TMEM column of its hi part
                uint32_t vh[32], vl[32];
                tmem_ld32(t_lane + cbase, vh);
                tmem_ld32(t_lane + BN_OUT + cbase, vl);
                tmem_ld_wait();
                if (park) {  // raw s32 accumulators, [512 / 4][128] uint4 per CTA: consecutive lanes write consecutive 16 bytes
                    float4* dh = a.parts + ((size_t)block * (kTmemCols / 4) + cbase / 4) * BM + quad * 32 + lane;
                    float4* dl = dh + (size_t)(BN_OUT / 4) * BM;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        dh[j * BM] = make_float4(__uint_as_float(vh[4 * j]), __uint_as_float(vh[4 * j + 1]),
                                                 __uint_as_float(vh[4 * j + 2]), __uint_as_float(vh[4 * j + 3]));
                        dl[j * BM] = make_float4(__uint_as_float(vl[4 * j]), __uint_as_float(vl[4 * j + 1]),
                                                 __uint_as_float(vl[4 * j + 2]), __uint_as_float(vl[4 * j + 3]));
                    }
                    continue;
                }
                if (fold) {  // integer adds: exact, order-free
                    const float4* sh = a.parts + ((size_t)peer * (kTmemCols / 4) + cbase / 4) * BM + quad * 32 + lane;
                    const float4* sl = sh + (size_t)(BN_OUT / 4) * BM;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 oh = __ldcg(sh + j * BM), ol = __ldcg(sl + j * BM);
                        vh[4 * j] += __float_as_uint(oh.x);
                        vh[4 * j + 1] += __float_as_uint(oh.y);
                        vh[4 * j + 2] += __float_as_uint(oh.z);
                        vh[4 * j + 3] += __float_as_uint(oh.w);
                        vl[4 * j] += __float_as_uint(ol.x);
                        vl[4 * j + 1] += __float_as_uint(ol.y);
                        vl[4 * j + 2] += __float_as_uint(ol.z);
                        vl[4 * j + 3] += __float_as_uint(ol.w);
                    }
                }
                if (row < a.m) {
                    float* dst = a.out + row * a.ld_out + (int64_t)cb * BN_OUT + cbase;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        // both accumulators are exact integers below 2^24 in magnitude for K <= 32768: exact in fp32
                        float4 o = make_float4(fmaf((float)(int)vh[j], 256.f, (float)(int)vl[j]) * a.alpha,
                                               fmaf((float)(int)vh[j + 1], 256.f, (float)(int)vl[j + 1]) * a.alpha,
                                               fmaf((float)(int)vh[j + 2], 256.f, (float)(int)vl[j + 2]) * a.alpha,
                                               fmaf((float)(int)vh[j + 3], 256.f, (float)(int)vl[j + 3]) * a.alpha);
                        if (a.accumulate) {
                            const float4 old = *reinterpret_cast<const float4*>(dst + j);
                            o.x += old.x;
                            o.y += old.y;
                            o.z += old.z;
                            o.w += old.w;
                        }
                        *reinterpret_cast<float4*>(dst + j) = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty), 0));
            if (park) {
                __threadfence();
                __syncwarp();
                if (lane == 0) red_release_add_u32(a.flags + block, 1u);
            }
            if (fold) {
                named_bar_sync(1, kEpiWarps * 32);
                if (warp == kEpiWarp0 && lane == 0) a.flags[peer] = 0u;
            }
            ++it;
        });
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == kTmaWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

PB2_KNOB_U32 g_mn_lbo = 8192, g_mn_sbo = 1024, g_mn_kstep = 2048;
PB2_KNOB g_units_cap = 0;

constexpr int64_t kFlagBytes = 1024;  // flags of up to 256 CTAs, then the parked accumulators
static int64_t workspace_bytes() { return kFlagBytes + (int64_t)sm_count() * BM * 512 * 4; }

// Tensor maps and kernel arguments of one product (no stream-K decision yet).
template <bool kTranspose, int BN, int kCtas>
static int prepare(const void* g, int g_fmt, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z, int z_fmt, int dim,
                   int64_t ldz, float alpha, int accumulate, float* out, int64_t ld_out, CUtensorMap* tg, CUtensorMap* tz,
                   Args* pa) {
    const int64_t m = kTranspose ? g_cols : g_rows;
    const int64_t k = kTranspose ? g_rows : g_cols;
    int rc = make_tmap_2d(tg, g, 2, (uint64_t)g_rows, (uint64_t)g_cols, (uint64_t)ld_g * 2, kTranspose ? 64 : BM, 64);
    if (rc) return rc;
    rc = make_tmap_2d(tz, z, 2, (uint64_t)k, (uint64_t)dim, (uint64_t)ldz * 2, 64, 64);
    if (rc) return rc;
    constexpr int kRowsPerUnit = BM * kCtas;  // output rows of one CTA (pair) tile
    Args& a = *pa;
    a.m = m;
    a.k = k;
    a.n_rb = (int)((m + kRowsPerUnit - 1) / kRowsPerUnit);
    a.n_cb = dim / BN;
    a.n_tiles = (int64_t)a.n_rb * a.n_cb;
    a.kblocks = (int)((k + BK - 1) / BK);
    a.alpha = alpha;
    a.accumulate = accumulate;
    a.out = out;
    a.ld_out = ld_out;
    a.mn_lbo = g_mn_lbo;
    a.mn_sbo = g_mn_sbo;
    a.mn_kstep = g_mn_kstep;
    a.idesc = make_idesc(kRowsPerUnit, BN > 256 ? 256 : BN, (uint32_t)g_fmt, (uint32_t)z_fmt, kTranspose ? kMajorMN : kMajorK, kMajorMN);
    a.n_groups = 0;
    a.flags = nullptr;
    a.parts = nullptr;
    a.z_lo_col = 0;
    return PB2_OK;
}

template <bool kTranspose, int BN, int kCtas>
static int launch(const void* g, int g_fmt, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z, int z_fmt,
                  int dim, int64_t ldz, float alpha, int accumulate, float* out, int64_t ld_out, void* workspace,
                  cudaStream_t st) {
    CUtensorMap tg, tz;
    Args a;
    int rc = prepare<kTranspose, BN, kCtas>(g, g_fmt, g_rows, g_cols, ld_g, z, z_fmt, dim, ldz, alpha, accumulate, out, ld_out,
                                            &tg, &tz, &a);
    if (rc) return rc;
    // stream-K when whole tiles would leave part of the machine idle in the last wave and every group's
    // range spans at least one full tile (so a row block is cut at most once)
    int units = sm_count() / kCtas;  // CTAs or CTA pairs the machine runs at once
    if (g_units_cap > 0 && g_units_cap < units) units = g_units_cap;  // measurement hook (pb2_debug_gg_units)
    if (workspace && a.n_cb <= units && units * kCtas <= 256) {
        const int groups = units / a.n_cb;
        const int64_t total = (int64_t)a.n_rb * a.kblocks;
        if (a.n_tiles > units && a.n_tiles % units != 0 && total / groups >= a.kblocks) {
            a.n_groups = groups;
            a.flags = static_cast<uint32_t*>(workspace);
            a.parts = reinterpret_cast<float4*>(static_cast<char*>(workspace) + kFlagBytes);
        }
    }
    auto kern = grad_gemm_kernel<kTranspose, BN, kCtas>;
    constexpr int smem = Smem<BN, kCtas>::kTotal;
    static PerDeviceOnce configured;  // per instantiation and device
    rc = ensure_dynamic_smem(configured, kern, smem, "grad_gemm");
    if (rc) return rc;
    const int n_units = a.n_groups > 0 ? a.n_groups * a.n_cb : (int)std::min<int64_t>(a.n_tiles, units);
    rc = check_cuda(launch_ex(kern, (unsigned)(n_units * kCtas), (unsigned)kThreads, (size_t)smem, st, kCtas, tg, tz, a),
                    "grad_gemm launch");
    if (rc) return rc;
    return check_launch("grad_gemm");
}


// out (=|+=) alpha * op(G) * Xhat for a one-byte G and the two-plane embedding operand (see grad_gemm_i8_kernel).
template <bool kTranspose>
static int launch_i8(const void* g, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z, int dim, int64_t ldz, float alpha,
                     int accumulate, float* out, int64_t ld_out, void* workspace, cudaStream_t st) {
    const int64_t m = kTranspose ? g_cols : g_rows;
    const int64_t k = kTranspose ? g_rows : g_cols;
    CUtensorMap tg, tz;
    int rc = make_tmap_2d(&tg, g, 1, (uint64_t)g_rows, (uint64_t)g_cols, (uint64_t)ld_g, 128, 128);
    if (rc) return rc;
    rc = make_tmap_2d(&tz, z, 1, (uint64_t)k, (uint64_t)(2 * dim), (uint64_t)ldz, 128, 128);
    if (rc) return rc;
    Args a;
    a.m = m;
    a.k = k;
    a.n_rb = (int)((m + 2 * BM - 1) / (2 * BM));
    a.n_cb = dim / i8::BN_OUT;
    a.n_tiles = (int64_t)a.n_rb * a.n_cb;
    a.kblocks = (int)((k + i8::kBK8 - 1) / i8::kBK8);
    a.alpha = alpha / i8::kScale;
    a.accumulate = accumulate;
    a.out = out;
    a.ld_out = ld_out;
    a.mn_lbo = a.mn_sbo = a.mn_kstep = a.idesc = 0;
    a.n_groups = 0;
    a.flags = nullptr;
    a.parts = nullptr;
    a.z_lo_col = dim;
    int units = sm_count() / 2;
    if (g_units_cap > 0 && g_units_cap < units) units = g_units_cap;
    if (workspace && a.n_cb <= units && units * 2 <= 256) {
        const int groups = units / a.n_cb;
        const int64_t total = (int64_t)a.n_rb * a.kblocks;
        if (a.n_tiles > units && a.n_tiles % units != 0 && total / groups >= a.kblocks) {
            a.n_groups = groups;
            a.flags = static_cast<uint32_t*>(workspace);
            a.parts = reinterpret_cast<float4*>(static_cast<char*>(workspace) + kFlagBytes);
        }
    }
    auto kern = grad_gemm_i8_kernel<kTranspose>;
    static PerDeviceOnce configured;
    rc = ensure_dynamic_smem(configured, kern, i8::kTotal, "grad_gemm_i8");
    if (rc) return rc;
    const int n_units = a.n_groups > 0 ? a.n_groups * a.n_cb : (int)std::min<int64_t>(a.n_tiles, units);
    rc = check_cuda(launch_ex(kern, (unsigned)(n_units * 2), (unsigned)kThreads, (size_t)i8::kTotal, st, 2, tg, tz, a),
                    "grad_gemm_i8 launch");
    if (rc) return rc;
    return check_launch("grad_gemm_i8");
}

// out0 = G Z0 and out1 = G^T Z1 in one grid (every tile of both products resident at once).
template <int BN>
static int launch_dual(const void* g, int g_fmt, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z0, const void* z1,
                       int z_fmt, int dim, int64_t ldz0, int64_t ldz1, float alpha, float* out0, float* out1, int64_t ld_out0,
                       int64_t ld_out1, const HingeFold& fold, cudaStream_t st) {
    CUtensorMap tg0, tz0, tg1, tz1;
    Args a0, a1;
    int rc = prepare<false, BN, 1>(g, g_fmt, g_rows, g_cols, ld_g, z0, z_fmt, dim, ldz0, alpha, 0, out0, ld_out0, &tg0, &tz0, &a0);
    if (rc) return rc;
    rc = prepare<true, BN, 1>(g, g_fmt, g_rows, g_cols, ld_g, z1, z_fmt, dim, ldz1, alpha, 0, out1, ld_out1, &tg1, &tz1, &a1);
    if (rc) return rc;
    auto kern = grad_gemm_dual_kernel<BN>;
    constexpr int smem = Smem<BN, 1>::kTotal;
    static PerDeviceOnce configured;
    rc = ensure_dynamic_smem(configured, kern, smem, "grad_gemm_dual");
    if (rc) return rc;
    const int n0 = (int)a0.n_tiles, n1 = (int)a1.n_tiles;
    const unsigned grid = (unsigned)(n0 + n1) + (fold.loss_out ? 1u : 0u);
    rc = check_cuda(launch_ex(kern, grid, (unsigned)kThreads, (size_t)smem, st, 1, tg0, tz0, tg1, tz1, a0, a1, n0, n1, fold),
                    "grad_gemm_dual launch");
    if (rc) return rc;
    return check_launch("grad_gemm_dual");
}

PB2_KNOB g_pair_mode = -1;  // measurement build (pb2_debug_gg_pair): -1 = automatic, 0 = never, 1 = whenever legal

}  // namespace gg
}  // namespace pb2

using namespace pb2;

#ifdef PB2_MEASURE
extern "C" int pb2_debug_set_mn_desc(uint32_t lbo, uint32_t sbo, uint32_t kstep) {
    gg::g_mn_lbo = lbo;
    gg::g_mn_sbo = sbo;
    gg::g_mn_kstep = kstep;
    return PB2_OK;
}

extern "C" int pb2_debug_gg_units(int cap) {
    gg::g_units_cap = cap;
    return PB2_OK;
}
extern "C" int pb2_debug_gg_pair(int mode) {
    gg::g_pair_mode = mode;
    return PB2_OK;
}
#endif

int pb2::grad_gemm_dual_fold(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z0,
                             const void* z1, int z_dtype, int dim, int64_t ldz0, int64_t ldz1, float alpha, float* out0,
                             float* out1, int64_t ld_out0, int64_t ld_out1, const HingeFold& fold, bool* folded, void* stream) {
    if (folded) *folded = false;
    if (g_rows <= 0 || g_cols <= 0) return PB2_OK;
    if (!gmat || !z0 || !z1 || !out0 || !out1) return set_error(PB2_ERR_ARG, "grad_gemm_dual: null");
    // one wave: every 128 x 64 tile of both products gets its own CTA; otherwise two ordinary launches
    const int64_t tiles = ((g_rows + gg::BM - 1) / gg::BM + (g_cols + gg::BM - 1) / gg::BM) * (dim / 64);
    const bool ok16 = (g_dtype == PB2_F16 || g_dtype == PB2_BF16) && (z_dtype == PB2_F16 || z_dtype == PB2_BF16);
    if (dim <= 0 || dim % 64 != 0 || tiles > sm_count() || !ok16 || (reinterpret_cast<uintptr_t>(out0) & 15) ||
        (reinterpret_cast<uintptr_t>(out1) & 15) || ld_out0 % 4 != 0 || ld_out1 % 4 != 0) {
        int rc = pb2_grad_gemm(gmat, g_dtype, g_rows, g_cols, ld_g, 0, z0, z_dtype, dim, ldz0, alpha, 0, out0, ld_out0, stream);
        if (rc) return rc;
        return pb2_grad_gemm(gmat, g_dtype, g_rows, g_cols, ld_g, 1, z1, z_dtype, dim, ldz1, alpha, 0, out1, ld_out1, stream);
    }
    const int gf = g_dtype == PB2_F16 ? (int)kFmtF16 : (int)kFmtBF16;
    const int zf = z_dtype == PB2_F16 ? (int)kFmtF16 : (int)kFmtBF16;
    if (folded) *folded = fold.loss_out != nullptr;
    return gg::launch_dual<64>(gmat, gf, g_rows, g_cols, ld_g, z0, z1, zf, dim, ldz0, ldz1, alpha, out0, out1, ld_out0, ld_out1,
                               fold, (cudaStream_t)stream);
}

extern "C" int pb2_grad_gemm_dual(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z0,
                                  const void* z1, int z_dtype, int dim, int64_t ldz0, int64_t ldz1, float alpha, float* out0,
                                  float* out1, int64_t ld_out0, int64_t ld_out1, void* stream) {
    return grad_gemm_dual_fold(gmat, g_dtype, g_rows, g_cols, ld_g, z0, z1, z_dtype, dim, ldz0, ldz1, alpha, out0, out1, ld_out0,
                               ld_out1, HingeFold(), nullptr, stream);
}

extern "C" int64_t pb2_grad_gemm_workspace(void) { return gg::workspace_bytes(); }

extern "C" int pb2_grad_gemm(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g,
                             int transpose, const void* z, int z_dtype, int dim, int64_t ldz, float alpha,
                             int accumulate, float* out, int64_t ld_out, void* stream) {
    return pb2_grad_gemm_ws(gmat, g_dtype, g_rows, g_cols, ld_g, transpose, z, z_dtype, dim, ldz, alpha, accumulate, out,
                            ld_out, nullptr, 0, stream);
}

extern "C" int pb2_grad_gemm_ws(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g,
                                int transpose, const void* z, int z_dtype, int dim, int64_t ldz, float alpha,
                                int accumulate, float* out, int64_t ld_out, void* workspace, int64_t workspace_bytes,
                                void* stream) {
    if (g_rows <= 0 || g_cols <= 0) return PB2_OK;
    if (workspace && (workspace_bytes < gg::workspace_bytes() || (reinterpret_cast<uintptr_t>(workspace) & 255)))
        return set_error(PB2_ERR_ARG, "grad_gemm: workspace must be 256-byte aligned and pb2_grad_gemm_workspace() bytes");
    if (!gmat || !z || !out) return set_error(PB2_ERR_ARG, "grad_gemm: null");
    if (dim <= 0 || dim % 64 != 0) return set_error(PB2_ERR_ARG, "grad_gemm: dim must be a multiple of 64");
    if ((reinterpret_cast<uintptr_t>(out) & 15) || ld_out % 4 != 0)
        return set_error(PB2_ERR_ARG, "grad_gemm: out must be 16-byte aligned with ld_out %% 4 == 0");
    if (g_dtype == PB2_U8 || z_dtype == PB2_I8_PLANES) {  // one-byte gradient matrix x two-plane embeddings: kind::i8
        if (g_dtype != PB2_U8 || z_dtype != PB2_I8_PLANES)
            return set_error(PB2_ERR_ARG, "grad_gemm: a PB2_U8 gradient matrix goes with a PB2_I8_PLANES operand (pb2_rows_quant_i8)");
        if (dim % 256 != 0) return set_error(PB2_ERR_ARG, "grad_gemm: the kind::i8 path needs dim %% 256 == 0");
        if (ldz < 2 * (int64_t)dim) return set_error(PB2_ERR_ARG, "grad_gemm: the two-plane operand is [K, 2 dim] bytes");
        if ((transpose ? g_rows : g_cols) > (1 << 15))
            return set_error(PB2_ERR_ARG, "grad_gemm: kind::i8 contraction length is limited to 32768 per call (exact fp32 join)");
        if (transpose)
            return gg::launch_i8<true>(gmat, g_rows, g_cols, ld_g, z, dim, ldz, alpha, accumulate, out, ld_out, workspace,
                                       (cudaStream_t)stream);
        return gg::launch_i8<false>(gmat, g_rows, g_cols, ld_g, z, dim, ldz, alpha, accumulate, out, ld_out, workspace,
                                    (cudaStream_t)stream);
    }
    if ((g_dtype != PB2_F16 && g_dtype != PB2_BF16) || (z_dtype != PB2_F16 && z_dtype != PB2_BF16))
        return set_error(PB2_ERR_ARG, "grad_gemm: operands must be fp16 or bf16");
    cudaStream_t st = (cudaStream_t)stream;
    // widest output tile that divides dim and still gives about one tile per SM (small batches are
    // latency bound: 1024 rows x 512 dims = 64 tiles of 128 x 64 instead of 16 of 128 x 256)
    const int64_t m_rows = transpose ? g_cols : g_rows;
    const int64_t rb = (m_rows + gg::BM - 1) / gg::BM;
    const int64_t want = std::min<int64_t>(sm_count(), 64);
    int bn = 64;
    if (dim % 256 == 0 && rb * (dim / 256) >= want) bn = 256;
    else if (dim % 128 == 0 && rb * (dim / 128) >= want) bn = 128;
    const int gf = g_dtype == PB2_F16 ? (int)kFmtF16 : (int)kFmtBF16;
    const int zf = z_dtype == PB2_F16 ? (int)kFmtF16 : (int)kFmtBF16;
#define PB2_GG(T, B, C) \
    gg::launch<T, B, C>(gmat, gf, g_rows, g_cols, ld_g, z, zf, dim, ldz, alpha, accumulate, out, ld_out, workspace, st)
    // CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles per pair) once there are enough pair tiles to fill
    // the machine: a third less operand traffic per SM than two independent 128 x 256 tiles
    // g_pair_mode: -1 automatic, 0 never, 1 pairs of 256-wide tiles, 2 pairs of 512-wide tiles
    const int64_t pair_rb = (m_rows + 2 * gg::BM - 1) / (2 * gg::BM);
    const int64_t kblocks = ((transpose ? g_rows : g_cols) + gg::BK - 1) / gg::BK;
    const bool pair = bn == 256 && (gg::g_pair_mode == 1 || (gg::g_pair_mode < 0 && pair_rb * (dim / 256) >= sm_count() / 2));
    // a pair tile spanning 512 output columns reads each G tile once chip-wide and a third fewer operand bytes
    // than two 256-wide tiles; its single TMEM stage exposes the epilogue, so only for long contractions
    const bool wide = dim % 512 == 0 && bn == 256 &&
                      (gg::g_pair_mode == 2 || (gg::g_pair_mode < 0 && pair_rb * (dim / 512) >= sm_count() / 2 && kblocks >= 128));
    if (transpose) {
        if (wide) return PB2_GG(true, 512, 2);
        if (pair) return PB2_GG(true, 256, 2);
        if (bn == 256) return PB2_GG(true, 256, 1);
        if (bn == 128) return PB2_GG(true, 128, 1);
        return PB2_GG(true, 64, 1);
    }
    if (wide) return PB2_GG(false, 512, 2);
    if (pair) return PB2_GG(false, 256, 2);
    if (bn == 256) return PB2_GG(false, 256, 1);
    if (bn == 128) return PB2_GG(false, 128, 1);
    return PB2_GG(false, 64, 1);
#undef PB2_GG
}
