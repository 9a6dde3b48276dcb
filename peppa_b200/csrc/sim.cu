// Kernels (a)/(b): the similarity matrix S = X * Y^T on the 5th-generation tensor cores with the
// consumer of S fused into the epilogue, so S never reaches HBM.
//
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of X [128 x 64] and Y [BN x 64] (bf16,
//               128-byte swizzle) into a ring of shared-memory stages, mbarrier full/empty.
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) from the
//               shared-memory descriptors into one of two TMEM accumulator stages.
//   warp 2      TMEM allocator.
//   warps 4-11  epilogue: tcgen05.ld 32 columns at a time (thread == row), apply the row/column
//               scales and the mode's reduction; the MMA of tile t+1 overlaps the epilogue of t.
//
// Persistent: gridDim = #SMs, tiles walked in bands of 8 row blocks so that concurrently
// running CTAs share X and Y tiles through L2.
//
// Modes (one template instantiation each):
//   Store   pig/util.py:9-13   cosine_matrix (the only mode that writes S)
//   Rank    pig/metrics.py:7-40  count of candidates closer than the positive
//   Hinge   pig/loss.py:41-48  symmetric margin loss + indicator counts + fp16 gradient matrix
//   LseRow  pig/loss.py:19-25  online row log-sum-exp partials
//   LseGrad MIL-NCE gradient matrix from the merged row/column statistics
#include "common.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {

constexpr int BM = 128;       // tile rows  (UMMA M)
constexpr int BK = 64;        // K elements per stage (one 128-byte swizzle row of bf16)
constexpr int UK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kEpiWarp0 = 4;  // first epilogue warp
constexpr int kEpiWarps = 8;
constexpr int kThreads = (kEpiWarp0 + kEpiWarps) * 32;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kGroupM = 8;
constexpr int kMaxColVecs = 3;

struct SimCommon {
    int64_t rows, cols;
    int kblocks;  // dim / 64
    int n_rb, n_cb;
    int64_t n_tiles;
    const float* rinv_x;  // may be null (=1)
    const float* rinv_y;  // may be null (=1)
    float scale;
};

__device__ __forceinline__ void tile_coords(int64_t t, int n_rb, int n_cb, int& rb, int& cb) {
    const int64_t per_band = (int64_t)kGroupM * n_cb;
    const int band = (int)(t / per_band);
    const int first = band * kGroupM;
    const int gsz = min(kGroupM, n_rb - first);
    const int in = (int)(t - band * per_band);
    rb = first + in % gsz;
    cb = in / gsz;
}

struct TileCtx {
    int64_t row0, col0;  // tile origin
    int64_t row;         // this thread's row
    bool row_valid;
    int cols_valid;  // number of valid columns in this tile (<= BN)
    int cb;          // column-block index
    int half;        // which column half this warp covers
};

// ------------------------------------------------------------------------------- epilogues
// Each policy: Params (POD, kernel argument), kColVecs (per-column fp32 vectors staged in smem;
// vector 0 is always rinv_y or 1), per-thread state as members.

struct StorePolicy {
    struct Params {
        float* out;
        int64_t ld;
    };
    static constexpr int kColVecs = 1;
    float ri;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void load_col(const Params&, const SimCommon& c, int64_t col, bool valid, float* v) {
        v[0] = valid ? (c.rinv_y ? c.rinv_y[col] : 1.f) : 0.f;
    }
    __device__ void tile_begin(const Params&, const SimCommon& c, const TileCtx& t) {
        ri = (t.row_valid ? (c.rinv_x ? c.rinv_x[t.row] : 1.f) : 0.f) * c.scale;
    }
    __device__ void chunk(const Params& p, const SimCommon&, const TileCtx& t, int cbase, const uint32_t (&v)[32],
                          const float* cv) {
        if (!t.row_valid) return;
        float* dst = p.out + t.row * p.ld + t.col0 + cbase;
        const int nvalid = t.cols_valid - cbase;
        if (nvalid >= 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 o;
                o.x = __uint_as_float(v[j]) * ri * cv[j];
                o.y = __uint_as_float(v[j + 1]) * ri * cv[j + 1];
                o.z = __uint_as_float(v[j + 2]) * ri * cv[j + 2];
                o.w = __uint_as_float(v[j + 3]) * ri * cv[j + 3];
                *reinterpret_cast<float4*>(dst + j) = o;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) dst[j] = __uint_as_float(v[j]) * ri * cv[j];
        }
    }
    __device__ void tile_end(const Params&, const SimCommon&, const TileCtx&) {}
    __device__ void kernel_end(const Params&, float*) {}
};

struct RankPolicy {
    struct Params {
        const float* pos_dist;
        const int64_t* pos_col;
        int64_t col_offset;
        int32_t* rank;
    };
    static constexpr int kColVecs = 1;
    float ri, pd;
    int pc;  // positive's column relative to the tile origin (may be out of range)
    int cnt;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void load_col(const Params&, const SimCommon& c, int64_t col, bool valid, float* v) {
        // an out-of-range column must never count: NaN makes every comparison false
        v[0] = valid ? (c.rinv_y ? c.rinv_y[col] : 1.f) : __int_as_float(0x7fc00000);
    }
    __device__ void tile_begin(const Params& p, const SimCommon& c, const TileCtx& t) {
        cnt = 0;
        if (t.row_valid) {
            ri = (c.rinv_x ? c.rinv_x[t.row] : 1.f) * c.scale;
            pd = p.pos_dist[t.row];
            const int64_t rel = p.pos_col[t.row] - p.col_offset - t.col0;
            pc = (rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1;
        } else {
            ri = 0.f;
            pd = -__int_as_float(0x7f800000);  // -inf: nothing is closer
            pc = -1;
        }
    }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx&, int cbase, const uint32_t (&v)[32],
                          const float* cv) {
        // fl32(1 - s) < fl32(1 - s_pos), exactly the comparison argsort resolves in pig/metrics.py:8-12
        int c = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float s = __fmul_rn(__fmul_rn(__uint_as_float(v[j]), ri), cv[j]);
            c += (__fsub_rn(1.0f, s) < pd) ? 1 : 0;  // no FMA contraction: fl32(1 - fl32(s))
        }
        const int rel = pc - cbase;
        if (rel >= 0 && rel < 32) {  // the positive itself sits in this chunk: take its vote back
            float sv = 0.f, cj = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j == rel) {
                    sv = __uint_as_float(v[j]);
                    cj = cv[j];
                }
            c -= (__fsub_rn(1.0f, __fmul_rn(__fmul_rn(sv, ri), cj)) < pd) ? 1 : 0;
        }
        cnt += c;
    }
    __device__ void tile_end(const Params& p, const SimCommon&, const TileCtx& t) {
        if (t.row_valid && cnt) atomicAdd(p.rank + t.row, cnt);
    }
    __device__ void kernel_end(const Params&, float*) {}
};

struct HingeParams {
    const float* diag_row;
    const float* diag_col;
    int64_t row_offset, col_offset;
    float margin;
    float* loss_partial;
    int32_t* row_cnt;
    int32_t* col_cnt;
    __half* gmat;
    int64_t ld_g;
    const float* pos_dist;  // kRank only: fl32(1 - diag_row)
    int32_t* rank;          // kRank only
};
constexpr int kColVecStride = 256;  // floats between column vectors in smem (= max BN)

// kRank additionally counts, per row, the columns closer than the diagonal (recall@k of the same
// similarity pass: the gallery workload shares one S pass between pig.loss and pig.metrics).
template <bool kRank>
struct HingePolicyT {
    using Params = HingeParams;
    static constexpr int kColVecs = 2;  // rinv_y, diag_col - margin
    // Instruction budget: the MMA of a 128 x 256 x 512 tile takes ~4096 cycles, i.e. ~16 issue slots per
    // element for the epilogue.  So: no per-element adds for z = margin + s - d; the indicators are
    // threshold compares (fl32(s + c) >= 0  <=>  s >= -c exactly), the loss is accumulated as
    // sum (ic + ir) * s per element and completed from the counts afterwards
    // (pb2_hinge_loss_terms: + sum_j (m - d_j) col_cnt[j] + sum_i (m - d_i) row_cnt[i]); scores use
    // packed FMUL2; column counts are byte-packed per thread and summed across the warp with REDUX.
    float ri, thr_r, loss, pd;
    int rcnt, dcol, rk;
    __device__ void kernel_begin(const Params&) { loss = 0.f; }
    __device__ static void load_col(const Params& p, const SimCommon& c, int64_t col, bool valid, float* v) {
        v[0] = valid ? (c.rinv_y ? c.rinv_y[col] : 1.f) : 0.f;
        v[1] = valid ? -(p.margin - p.diag_col[col]) : 0.f;
    }
    __device__ void tile_begin(const Params& p, const SimCommon& c, const TileCtx& t) {
        rcnt = 0;
        rk = 0;
        pd = (kRank && t.row_valid) ? p.pos_dist[t.row] : 0.f;
        if (t.row_valid) {
            ri = c.rinv_x ? c.rinv_x[t.row] : 1.f;
            thr_r = -(p.margin - p.diag_row[t.row]);
            const int64_t rel = (p.row_offset + t.row) - p.col_offset - t.col0;
            dcol = (rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1;
        } else {
            ri = 0.f;
            thr_r = 0.f;
            dcol = -1;
        }
    }
    // kSlow: chunks that contain the diagonal, out-of-range columns or out-of-range rows; those
    // elements get a hugely negative (finite) score so that no indicator fires and 0 * s stays 0.
    template <bool kSlow>
    __device__ __forceinline__ void chunk_impl(const Params& p, const TileCtx& t, int cbase, const uint32_t (&v)[32],
                                               const float* cv) {
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float4* ct4 = reinterpret_cast<const float4*>(cv + kColVecStride);
        const float2 ri2 = make_float2(ri, ri);
        const int drel = dcol - cbase;                          // diagonal position inside this chunk
        const int nvalid = t.row_valid ? t.cols_valid - cbase : 0;  // valid columns of this row's chunk
        uint32_t packed[16], pk[8];
        float l = 0.f;
        int rc = 0, rkk = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 c4 = cv4[q], t4 = ct4[q];
            // fl32(fl32(v * ri) * cj): the same two roundings as the rank kernel and pair_dot
            const float2 s01 = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), ri2),
                                          make_float2(c4.x, c4.y));
            const float2 s23 = __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), ri2),
                                          make_float2(c4.z, c4.w));
            float sv[4] = {s01.x, s01.y, s23.x, s23.y};
            const float tc[4] = {t4.x, t4.y, t4.z, t4.w};
            float g[4];
            uint32_t pkq = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (kSlow && ((4 * q + e) == drel || (4 * q + e) >= nvalid)) sv[e] = -3.0e38f;
                const bool ic = sv[e] >= tc[e];
                const bool ir = sv[e] >= thr_r;
                g[e] = (ic ? 1.f : 0.f) + (ir ? 1.f : 0.f);
                l = fmaf(g[e], sv[e], l);
                rc += ir ? 1 : 0;
                if (kRank) rkk += (__fsub_rn(1.0f, sv[e]) < pd) ? 1 : 0;
                pkq += ic ? (1u << (8 * e)) : 0u;
            }
            pk[q] = pkq;
            const __half2 h01 = __floats2half2_rn(g[0], g[1]), h23 = __floats2half2_rn(g[2], g[3]);
            packed[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
            packed[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        }
        loss += l;
        rcnt += rc;
        if (kRank) rk += rkk;
        // column counts: byte-packed (<= 32 per byte), summed over the warp's 32 rows with REDUX
        const int lane = lane_id();
        uint32_t mine = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t tot = __reduce_add_sync(0xffffffffu, pk[q]);
            if ((lane >> 2) == q) mine = tot;
        }
        const int ccnt = (int)((mine >> ((lane & 3) * 8)) & 0xffu);
        if (ccnt) atomicAdd(p.col_cnt + t.col0 + cbase + lane, ccnt);  // ccnt == 0 for out-of-range columns
        if (p.gmat && t.row_valid) {
            uint4* dst = reinterpret_cast<uint4*>(p.gmat + t.row * p.ld_g + t.col0 + cbase);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
        }
    }
    __device__ void chunk(const Params& p, const SimCommon&, const TileCtx& t, int cbase, const uint32_t (&v)[32],
                          const float* cv) {
        if (cbase >= t.cols_valid) return;  // warp-uniform
        const int drel = dcol - cbase;
        const bool slow = (drel >= 0 && drel < 32) || !t.row_valid || (t.cols_valid - cbase) < 32;
        if (__any_sync(0xffffffffu, slow)) chunk_impl<true>(p, t, cbase, v, cv);
        else chunk_impl<false>(p, t, cbase, v, cv);
    }
    __device__ void tile_end(const Params& p, const SimCommon&, const TileCtx& t) {
        if (t.row_valid && rcnt) atomicAdd(p.row_cnt + t.row, rcnt);
        if (kRank && t.row_valid && rk) atomicAdd(p.rank + t.row, rk);
    }
    __device__ void kernel_end(const Params& p, float* red) {
        // fixed-order reduction over the 256 epilogue threads -> one deterministic partial per CTA
        const int e = threadIdx.x - kEpiWarp0 * 32;
        const float w = warp_sum(loss);
        if ((e & 31) == 0) red[e >> 5] = w;
        named_bar_sync(2, kEpiThreads);
        if (e == 0) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < kEpiWarps; ++i) s += red[i];
            p.loss_partial[blockIdx.x] = s;
        }
    }
};

struct LseRowPolicy {
    struct Params {
        float* part_max;
        float* part_sum;
    };
    static constexpr int kColVecs = 1;
    float ri, m, s;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void load_col(const Params&, const SimCommon& c, int64_t col, bool valid, float* v) {
        v[0] = valid ? (c.rinv_y ? c.rinv_y[col] : 1.f) : 0.f;
    }
    __device__ void tile_begin(const Params&, const SimCommon& c, const TileCtx& t) {
        // work in the log2 domain: t = s_ij * log2(e)
        ri = (t.row_valid ? (c.rinv_x ? c.rinv_x[t.row] : 1.f) : 0.f) * c.scale * 1.4426950408889634f;
        m = -__int_as_float(0x7f800000);
        s = 0.f;
    }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx& t, int cbase, const uint32_t (&v)[32],
                          const float* cv) {
        const int nvalid = t.cols_valid - cbase;
        if (nvalid <= 0) return;
        float x[32];
        float cm = -__int_as_float(0x7f800000);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            x[j] = (j < nvalid) ? __uint_as_float(v[j]) * ri * cv[j] : -__int_as_float(0x7f800000);
            cm = fmaxf(cm, x[j]);
        }
        const float mn = fmaxf(m, cm);
        // mn == -inf only if every logit so far is -inf; keep the state untouched then
        if (mn > -__int_as_float(0x7f800000)) {
            float acc = s * exp2f(m - mn);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += exp2f(x[j] - mn);
            s = acc;
            m = mn;
        }
    }
    __device__ void tile_end(const Params& p, const SimCommon& c, const TileCtx& t) {
        if (!t.row_valid) return;
        const int64_t slot = ((int64_t)t.cb * 2 + t.half) * c.rows + t.row;
        p.part_max[slot] = m;
        p.part_sum[slot] = s;
    }
    __device__ void kernel_end(const Params&, float*) {}
};

struct LseGradPolicy {
    struct Params {
        const float* den_row;
        const float* den_col;
        __half* gmat;
        int64_t ld_g;
    };
    static constexpr int kColVecs = 2;  // rinv_y, 13 - den_col * log2e
    float ri, drow;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void load_col(const Params& p, const SimCommon& c, int64_t col, bool valid, float* v) {
        v[0] = valid ? (c.rinv_y ? c.rinv_y[col] : 1.f) : 0.f;
        v[1] = valid ? (13.0f - p.den_col[col] * 1.4426950408889634f) : -__int_as_float(0x7f800000);
    }
    __device__ void tile_begin(const Params& p, const SimCommon& c, const TileCtx& t) {
        ri = (t.row_valid ? (c.rinv_x ? c.rinv_x[t.row] : 1.f) : 0.f) * c.scale * 1.4426950408889634f;
        drow = t.row_valid ? (13.0f - p.den_row[t.row] * 1.4426950408889634f) : -__int_as_float(0x7f800000);
    }
    __device__ void chunk(const Params& p, const SimCommon&, const TileCtx& t, int cbase, const uint32_t (&v)[32],
                          const float* cv) {
        if (cbase >= t.cols_valid || !t.row_valid) return;
        const float* cd = cv + kColVecStride;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]) * ri * cv[j];
            const float g = exp2f(x + drow) + exp2f(x + cd[j]);
            const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn(g));
            if (j & 1) packed[j >> 1] |= h << 16;
            else packed[j >> 1] = h;
        }
        uint4* dst = reinterpret_cast<uint4*>(p.gmat + t.row * p.ld_g + t.col0 + cbase);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
    }
    __device__ void tile_end(const Params&, const SimCommon&, const TileCtx&) {}
    __device__ void kernel_end(const Params&, float*) {}
};

// ---------------------------------------------------------------------------------- kernel
template <int BN>
struct SimSmem {
    static constexpr int kStageBytes = (BM + BN) * BK * 2;
    static constexpr int kStages = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int kTileBytes = kStages * kStageBytes;
    static constexpr int kColVecBytes = 2 * kMaxColVecs * 256 * 4;  // [acc stage][vec][256]
    static constexpr int kBarBytes = 256;
    static constexpr int kTotal = 1024 /*align slack*/ + kTileBytes + kColVecBytes + kBarBytes;
};

template <class Policy, int BN>
__global__ void __launch_bounds__(kThreads, 1)
    sim_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, const SimCommon c,
               const typename Policy::Params p) {
    using L = SimSmem<BN>;
    // 128-byte-swizzled TMA/UMMA tiles need 1024-byte alignment; the kernel has no static shared
    // memory, so the dynamic segment starts at the (aligned) base of the CTA's shared window.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    float* colvec = reinterpret_cast<float*>(smem + L::kTileBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kColVecBytes);
    uint64_t* full = bars;                    // [kStages]
    uint64_t* empty = bars + L::kStages;      // [kStages]
    uint64_t* acc_full = bars + 2 * L::kStages;   // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [kEpiWarps]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_y);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, kEpiWarps);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
                int rb, cb;
                tile_coords(t, c.n_rb, c.n_cb, rb, cb);
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + stage * L::kStageBytes;
                    uint8_t* sy = sx + BM * BK * 2;
                    mbar_arrive_expect_tx(full + stage, L::kStageBytes);
                    tma_load_2d(sx, &tm_x, full + stage, kb * BK, rb * BM, kEvictNormal);
                    tma_load_2d(sy, &tm_y, full + stage, kb * BK, cb * BN, kEvictNormal);
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ====================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BM, BN, kFmtBF16, kFmtBF16, kMajorK, kMajorK);
            int stage = 0;
            uint32_t phase = 0;
            int64_t it = 0;
            for (int64_t t = blockIdx.x; t < c.n_tiles; t += gridDim.x, ++it) {
                const int as = (int)(it & 1);
                mbar_wait(acc_empty + as, (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
                    const uint32_t sx = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t sy = sx + BM * BK * 2;
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint64_t da = make_smem_desc(sx + k * UK * 2, 16, 1024);
                        const uint64_t db = make_smem_desc(sy + k * UK * 2, 16, 1024);
                        umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty + stage);  // stage reusable once these MMAs have read it
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(acc_full + as);  // accumulator complete -> epilogue
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ======================================================================== epilogue
        const int e = threadIdx.x - kEpiWarp0 * 32;  // 0..255
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
        const int half = (warp - kEpiWarp0) >> 2;    // column half
        constexpr int kChunksPerHalf = BN / 64;      // 32-column chunks per half (BN=64 -> 1)
        Policy pol;
        pol.kernel_begin(p);
        int64_t it = 0;
        for (int64_t t = blockIdx.x; t < c.n_tiles; t += gridDim.x, ++it) {
            const int as = (int)(it & 1);
            int rb, cb;
            tile_coords(t, c.n_rb, c.n_cb, rb, cb);
            TileCtx ctx;
            ctx.row0 = (int64_t)rb * BM;
            ctx.col0 = (int64_t)cb * BN;
            ctx.row = ctx.row0 + quad * 32 + lane;
            ctx.row_valid = ctx.row < c.rows;
            ctx.cols_valid = (int)min((int64_t)BN, c.cols - ctx.col0);
            ctx.cb = cb;
            ctx.half = half;
            float* cv = colvec + as * (kMaxColVecs * 256);
            for (int col = e; col < BN; col += kEpiThreads) {
                float tmp[kMaxColVecs];
                Policy::load_col(p, c, ctx.col0 + col, col < ctx.cols_valid, tmp);
#pragma unroll
                for (int k = 0; k < Policy::kColVecs; ++k) cv[k * 256 + col] = tmp[k];
            }
            pol.tile_begin(p, c, ctx);
            named_bar_sync(1, kEpiThreads);  // column vectors visible; previous user of this buffer done
            mbar_wait(acc_full + as, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
#pragma unroll 1
            for (int ch = 0; ch < kChunksPerHalf; ++ch) {
                const int cbase = (half * kChunksPerHalf + ch) * 32;
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(t_lane + cbase, v);
                tmem_ld_wait();
                pol.chunk(p, c, ctx, cbase, v, cv + cbase);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + as);
            pol.tile_end(p, c, ctx);
        }
        pol.kernel_end(p, red);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// ------------------------------------------------------------------------------------ host
static int pick_bn(int64_t rows, int64_t cols) {
    // widest tile that still yields at least ~one tile per SM; small problems are latency bound
    const int64_t sms = sm_count();
    const int64_t rb = (rows + BM - 1) / BM;
    for (int bn : {256, 128}) {
        if (rb * ((cols + bn - 1) / bn) >= sms) return bn;
    }
    return 64;
}

template <class Policy, int BN>
static int launch_sim(const void* x, const void* y, int64_t rows, int64_t cols, int dim, int64_t ldx, int64_t ldy,
                      const float* rinv_x, const float* rinv_y, float scale, const typename Policy::Params& pp,
                      cudaStream_t st, const char* what) {
    CUtensorMap tx, ty;
    int rc = make_tmap_2d(&tx, x, 2, (uint64_t)rows, (uint64_t)dim, (uint64_t)ldx * 2, BM, BK);
    if (rc) return rc;
    rc = make_tmap_2d(&ty, y, 2, (uint64_t)cols, (uint64_t)dim, (uint64_t)ldy * 2, BN, BK);
    if (rc) return rc;
    SimCommon c;
    c.rows = rows;
    c.cols = cols;
    c.kblocks = dim / BK;
    c.n_rb = (int)((rows + BM - 1) / BM);
    c.n_cb = (int)((cols + BN - 1) / BN);
    c.n_tiles = (int64_t)c.n_rb * c.n_cb;
    c.rinv_x = rinv_x;
    c.rinv_y = rinv_y;
    c.scale = scale;
    auto kern = sim_kernel<Policy, BN>;
    constexpr int smem = SimSmem<BN>::kTotal;
    static bool configured = false;  // per instantiation
    if (!configured) {
        rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), what);
        if (rc) return rc;
        configured = true;
    }
    const int grid = (int)std::min<int64_t>(c.n_tiles, pb2_sim_grid());
    kern<<<grid, kThreads, smem, st>>>(tx, ty, c, pp);
    return check_launch(what);
}

template <class Policy>
static int dispatch_sim(const void* x, const void* y, int64_t rows, int64_t cols, int dim, int64_t ldx, int64_t ldy,
                        const float* rinv_x, const float* rinv_y, float scale, const typename Policy::Params& pp,
                        void* stream, const char* what, int force_bn = 0) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (dim <= 0 || dim % BK != 0) return set_error(PB2_ERR_ARG, "%s: dim must be a positive multiple of 64", what);
    if (!x || !y) return set_error(PB2_ERR_ARG, "%s: null operand", what);
    if (rows > 0x7fffffffll * BM / 2 || cols > 0x7fffffffll) return set_error(PB2_ERR_ARG, "%s: too large", what);
    cudaStream_t st = (cudaStream_t)stream;
    const int bn = force_bn ? force_bn : pick_bn(rows, cols);
    switch (bn) {
        case 256:
            return launch_sim<Policy, 256>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, st, what);
        case 128:
            return launch_sim<Policy, 128>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, st, what);
        default:
            return launch_sim<Policy, 64>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, st, what);
    }
}

static int g_force_bn = 0;  // test hook (pb2_debug_force_bn)

}  // namespace pb2

using namespace pb2;

extern "C" int pb2_sim_grid(void) { return sm_count(); }
extern "C" int pb2_debug_force_bn(int bn) {
    g_force_bn = (bn == 64 || bn == 128 || bn == 256) ? bn : 0;
    return PB2_OK;
}

extern "C" int pb2_sim_matrix(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                              int64_t cols, int dim, int64_t ldx, int64_t ldy, float scale, float* out,
                              int64_t ld_out, void* stream) {
    if (!out && rows > 0 && cols > 0) return set_error(PB2_ERR_ARG, "sim_matrix: null output");
    StorePolicy::Params pp{out, ld_out};
    return dispatch_sim<StorePolicy>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                     "sim_matrix", g_force_bn);
}

extern "C" int pb2_sim_rank(const void* q, const void* g, const float* rinv_q, const float* rinv_g,
                            const float* pos_dist, const int64_t* pos_col, int64_t rows, int64_t cols,
                            int64_t col_offset, int dim, int64_t ldq, int64_t ldg, int32_t* rank, void* stream) {
    if (rows > 0 && cols > 0 && (!pos_dist || !pos_col || !rank)) return set_error(PB2_ERR_ARG, "sim_rank: null");
    RankPolicy::Params pp{pos_dist, pos_col, col_offset, rank};
    return dispatch_sim<RankPolicy>(q, g, rows, cols, dim, ldq, ldg, rinv_q, rinv_g, 1.0f, pp, stream, "sim_rank",
                                    g_force_bn);
}

extern "C" int pb2_sim_hinge(const void* x, const void* y, const float* rinv_x, const float* rinv_y,
                             const float* diag_row, const float* diag_col, int64_t rows, int64_t cols,
                             int64_t row_offset, int64_t col_offset, int dim, int64_t ldx, int64_t ldy, float margin,
                             float* loss_partial, int n_partials, int32_t* row_cnt, int32_t* col_cnt, void* gmat,
                             int64_t ld_g, const float* pos_dist, int32_t* rank, void* stream) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (!diag_row || !diag_col || !loss_partial || !row_cnt || !col_cnt)
        return set_error(PB2_ERR_ARG, "sim_hinge: null");
    if (n_partials < pb2_sim_grid()) return set_error(PB2_ERR_ARG, "sim_hinge: loss_partial too small");
    if (gmat && (ld_g % 8 != 0 || ld_g < ((cols + 31) / 32) * 32 || (reinterpret_cast<uintptr_t>(gmat) & 15)))
        return set_error(PB2_ERR_ARG, "sim_hinge: gmat needs 16-byte alignment and ld_g >= round_up(cols, 32)");
    int rc = check_cuda(cudaMemsetAsync(loss_partial, 0, sizeof(float) * n_partials, (cudaStream_t)stream),
                        "sim_hinge memset");
    if (rc) return rc;
    if ((pos_dist == nullptr) != (rank == nullptr))
        return set_error(PB2_ERR_ARG, "sim_hinge: pos_dist and rank go together");
    HingeParams pp{diag_row, diag_col, row_offset, col_offset,      margin, loss_partial,
                   row_cnt,  col_cnt,  (__half*)gmat, ld_g, pos_dist, rank};
    if (rank)
        return dispatch_sim<HingePolicyT<true>>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream,
                                                "sim_hinge+rank", g_force_bn);
    return dispatch_sim<HingePolicyT<false>>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream,
                                             "sim_hinge", g_force_bn);
}

// LSE partial layout is fixed to the 128-column tile so the caller can size buffers up front.
extern "C" int pb2_sim_lse_parts(int64_t cols) { return (int)((cols + 127) / 128) * 2; }

extern "C" int pb2_sim_lse_rows(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                                int64_t cols, int dim, int64_t ldx, int64_t ldy, float scale, float* part_max,
                                float* part_sum, void* stream) {
    if (rows > 0 && cols > 0 && (!part_max || !part_sum)) return set_error(PB2_ERR_ARG, "sim_lse_rows: null");
    LseRowPolicy::Params pp{part_max, part_sum};
    return dispatch_sim<LseRowPolicy>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                      "sim_lse_rows", 128);
}

extern "C" int pb2_sim_lse_grad(const void* x, const void* y, const float* rinv_x, const float* rinv_y,
                                const float* den_row, const float* den_col, int64_t rows, int64_t cols, int dim,
                                int64_t ldx, int64_t ldy, float scale, void* gmat, int64_t ld_g, void* stream) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (!den_row || !den_col || !gmat) return set_error(PB2_ERR_ARG, "sim_lse_grad: null");
    if (ld_g % 8 != 0 || ld_g < ((cols + 31) / 32) * 32 || (reinterpret_cast<uintptr_t>(gmat) & 15))
        return set_error(PB2_ERR_ARG, "sim_lse_grad: gmat needs 16-byte alignment and ld_g >= round_up(cols, 32)");
    LseGradPolicy::Params pp{den_row, den_col, (__half*)gmat, ld_g};
    return dispatch_sim<LseGradPolicy>(x, y, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                       "sim_lse_grad", g_force_bn);
}
