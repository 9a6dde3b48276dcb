// Kernels (a)/(b): the similarity matrix S = X * Y^T on the 5th-generation tensor cores with the
// consumer of S fused into the epilogue, so S never reaches HBM.
//
//   warps 0..4G-1   epilogue (G groups of 4 warps, one warp per TMEM lane quadrant and column group):
//                   tcgen05.ld 32 columns at a time (thread == row; with G <= 2 the load of chunk c+1 is in
//                   flight while chunk c is processed), row/column scales, the mode's reduction; the MMA of
//                   tile t+1 overlaps the epilogue of tile t (two TMEM accumulator stages).
//   warp 4G         vector loader: stages the next tile's per-column / per-row epilogue operands in shared
//                   memory and hands them over with an mbarrier.
//   warp 4G+1       MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16, kind::f16) from the
//                   shared-memory descriptors into one of two TMEM accumulator stages.
//   warp 4G+2       TMA producer: cp.async.bulk.tensor tiles of X [128 x 64] and Y [BN x 64] (bf16, 128-byte
//                   swizzle) into a ring of shared-memory stages, mbarrier full/empty; TMEM alloc/dealloc.
//
// Persistent: gridDim = #SMs, tiles walked in bands of 8 row blocks so that concurrently
// running CTAs share X and Y tiles through L2.
//
// Modes (one template instantiation each):
//   Store   pig/util.py:9-13   cosine_matrix (the only mode that writes S)
//   Rank    pig/metrics.py:7-40  count of candidates closer than the positive
//   Hinge   pig/loss.py:41-48  symmetric margin loss + indicator counts + fp16 gradient matrix
//           (optionally also the rank counts of the diagonal: loss and recall@k from one pass)
//   LseRow  pig/loss.py:19-25  online row log-sum-exp partials
//   LseGrad MIL-NCE gradient matrix from the merged row/column statistics
//   Diag    paired scores s_kk through the same tensor-core arithmetic (only the diagonal tiles)
// Gradient-matrix and score tiles leave the SM through per-warp swizzled staging buffers and TMA stores.
#include "common.cuh"
#include "host_util.h"
#include "peppa_b200.h"

#include <type_traits>

namespace pb2 {

constexpr int BM = 128;       // tile rows  (UMMA M)
constexpr int BK = 64;        // K elements per stage (one 128-byte swizzle row of bf16)
constexpr int UK = 16;        // K per tcgen05.mma for 16-bit inputs
// Warp roles.  The SM's warp arbiter prefers the HIGHEST warp id among eligible warps, so the warps on the
// critical path of the tensor pipeline get the highest ids: epilogue warps 0 .. 4G-1, then the vector
// loader (4G), the MMA issuer (4G+1) and the TMA producer (4G+2, also TMEM alloc/dealloc).
constexpr int kEpiWarp0 = 0;
constexpr int kAuxWarps = 3;
constexpr int kMaxRowVecs = 4;
constexpr int kMaxEpiWarps = 16;
// Epilogue warps come in groups of 4 (one per TMEM lane quadrant); G groups split the BN columns of a
// tile between them.  More groups = more warps per scheduler to hide the epilogue's latencies.
__host__ __device__ constexpr int sim_threads(int groups) { return (4 * groups + kAuxWarps) * 32; }
#ifndef PB2_GROUP_M
#define PB2_GROUP_M 8  // row blocks per band of the tile walk (measurement builds: tools/build_variant.sh X -DPB2_GROUP_M=16)
#endif
constexpr int kGroupM = PB2_GROUP_M;
constexpr int kMaxColVecs = 3;
constexpr int kColVecStride = 256;          // floats between column vectors in smem (= max BN)
constexpr int kOutSlabBytes = 32 * 64 * 2;  // one warp's [32 rows x 64 fp16] staging slab (4 KiB)
#define PB2_NAN __int_as_float(0x7fc00000)
#define PB2_INF __int_as_float(0x7f800000)
constexpr float kMasked = -3.0e38f;  // score of an excluded element: no indicator fires, 0 * s stays 0

struct SimCommon {
    int64_t rows, cols;
    int kblocks;  // dim / 64
    int n_rb, n_cb;
    int64_t n_tiles;
    const float* rinv_x;  // may be null (=1)
    const float* rinv_y;  // may be null (=1)
    float scale;
    int diag_only;  // visit only tiles (rb, rb): paired scores (BN == BM)
    uint32_t fmt;   // operand format of both X and Y: kFmtBF16 or kFmtF16 (kind::f16 takes either natively)
    // programmatic dependent launch: X and Y were complete before the predecessor passed ITS griddepcontrol.wait
    // (host_util.h: OperandsReadyScope), so the TMA / MMA warps start on them while the predecessor still runs
    int early_operands;
};

__device__ __forceinline__ void tile_coords(int64_t t, int n_rb, int n_cb, int& rb, int& cb) {
    if (n_cb == 0) {  // diag_only
        rb = cb = (int)t;
        return;
    }
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
    int half;        // which column group of the tile this warp covers
    int warp_cols;   // columns of the tile one epilogue warp covers (BN / G)
    int quad;        // TMEM lane quadrant of this warp (rows quad*32 .. +31 of the tile)
};

// Per-warp output staging for gradient-matrix tiles: two 4 KiB slabs, 128-byte swizzled, each
// written by the warp's 32 threads (one 128-byte row per thread) and drained by a TMA store.
// PB2_G_STORE_HINT (measurement builds): 1 = output tiles leave with an evict-first L2 hint, 2 = and operand loads
// evict-last.  Back to back on a 32768^2 block the hint is worth 2.5 % (hinge pass 1.249 -> 1.219 ms sustained: the
// 2 GiB block no longer evicts the 64 MB of operands the next launch re-reads); inside the real gallery step, where
// the gradient GEMMs stream 4 GiB through L2 between two hinge passes anyway, it is neutral (131072-clip step, two
// builds alternating in one call: 47.3 ms without, 47.6 ms with; tools/ab_gallery.py), so the default stays plain.
#ifndef PB2_G_STORE_HINT
#define PB2_G_STORE_HINT 0
#endif
#if PB2_G_STORE_HINT >= 2
#define PB2_OPERAND_POLICY kEvictLast
#else
#define PB2_OPERAND_POLICY kEvictNormal
#endif
struct OutStage {
    uint8_t* buf;             // this warp's nbuf * kOutSlabBytes
    const CUtensorMap* tmap;  // gradient matrix [rows, cols] fp16, box [32 x 64]
    uint32_t slab;            // running slab counter of this warp
    uint32_t mask;            // nbuf - 1 (1: double buffered, 0: one slab per tile and warp)
#ifdef PB2_MEASURE
    int skip;                 // measurement build only (pb2_debug_force_bn bits 16..18): 1 no TMA store, 2 no STS, 4 no fence
#else
    static constexpr int skip = 0;
#endif
    // 16 fp16 (two uint4) of this thread's row, chunk parity cp (0: columns 0-31, 1: columns 32-63)
    __device__ __forceinline__ void write(int lane, int cp, const uint32_t (&packed)[16]) {
        if (skip & 2) return;
        uint8_t* row = buf + (slab & mask) * kOutSlabBytes + lane * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c16 = (cp * 4 + k) ^ (lane & 7);  // 128B swizzle: 16-byte chunk index ^ (row % 8)
            *reinterpret_cast<uint4*>(row + c16 * 16) =
                make_uint4(packed[4 * k], packed[4 * k + 1], packed[4 * k + 2], packed[4 * k + 3]);
        }
    }
    // 32 bytes (two uint4) of this thread's row: chunk ch % 4 of a slab of 128 one-byte columns
    __device__ __forceinline__ void write_u8(int lane, int c4, const uint32_t (&w)[8]) {
        if (skip & 2) return;
        uint8_t* row = buf + (slab & mask) * kOutSlabBytes + lane * 128;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int c16 = (c4 * 2 + k) ^ (lane & 7);
            *reinterpret_cast<uint4*>(row + c16 * 16) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
        }
    }
    // 32 fp32 of this thread's row = one whole 128-byte swizzled row (slab = one 32-column chunk)
    __device__ __forceinline__ void write_f32(int lane, const float (&o)[32]) {
        uint8_t* row = buf + (slab & mask) * kOutSlabBytes + lane * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            *reinterpret_cast<float4*>(row + ((k ^ (lane & 7)) * 16)) =
                make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
    __device__ __forceinline__ void begin_slab(int lane) {  // the slab buffer about to be reused must be drained
        if (lane == 0) {
            if (mask) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
        }
        __syncwarp();
    }
    __device__ __forceinline__ void end_slab(int lane, int32_t col, int32_t row) {
        if (!(skip & 4)) fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
#if PB2_G_STORE_HINT
            if (!(skip & 1)) tma_store_2d_hint(tmap, buf + (slab & mask) * kOutSlabBytes, col, row, kEvictFirst);
#else
            if (!(skip & 1)) tma_store_2d(tmap, buf + (slab & mask) * kOutSlabBytes, col, row);
#endif
            tma_store_commit();
        }
        ++slab;
    }
    __device__ __forceinline__ void finish(int lane) {
        if (lane == 0) tma_store_wait_all<0>();
        __syncwarp();
    }
    // One 32-column chunk (16 packed fp16 pairs per thread) of this warp's share of a gradient-matrix tile: two
    // chunks fill a 64-column slab, which is fenced and handed to a TMA store while the next slab is written.
    // Measured and dropped (32768^2 block, 200 launches back to back, same board): staging both slabs of a 256-wide
    // tile and synchronising once per tile 1.27 -> 1.33 ms (the drain wait then follows the stores immediately; it
    // needs a second pair of slabs); draining the slab with LDS + whole-line STG instead of the proxy fence + TMA
    // store 1.28 -> 1.44 ms.
    __device__ __forceinline__ void stage(int lane, int ch, const uint32_t (&packed)[16], int32_t col_of_chunk, int32_t row) {
        if ((ch & 1) == 0) begin_slab(lane);
        write(lane, ch & 1, packed);
        if (ch & 1) end_slab(lane, col_of_chunk - 32, row);
    }
    // One-byte gradient matrix: four 32-column chunks (8 words of 4 bytes per thread each) fill a slab of 128 columns.
    __device__ __forceinline__ void stage_u8(int lane, int ch, const uint32_t (&w)[8], int32_t col_of_chunk, int32_t row) {
        if ((ch & 3) == 0) begin_slab(lane);
        write_u8(lane, ch & 3, w);
        if ((ch & 3) == 3) end_slab(lane, col_of_chunk - 96, row);
    }
};

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// [a >= b] as an exact 1.0f / 0.0f from ONE ALU-pipe instruction (FSET.BF): the sub-partition issues one
// FMA-pipe and one ALU-pipe instruction every other cycle each (tools/ubench/pipes.cu), so the epilogue
// splits its indicators between fma_sat (FMA pipe) and this.
__device__ __forceinline__ float fset_ge(float a, float b) {
    float d;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Largest float below a finite x (nextafterf(x, -inf)) without control flow: the vector-loader warp must
// keep every global load of a tile in flight at once, and libdevice's nextafterf puts branches between them
// (measured: twelve serialised round trips per tile, the loader -- not the MMA or the epilogue -- set the
// tile period).  NaN stays NaN; -inf stays -inf.
__device__ __forceinline__ float pred_f32(float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t down = (b & 0x7fffffffu) == 0u ? 0x80000001u : ((b >> 31) ? b + 1u : b - 1u);
    return (x != x || x == -PB2_INF) ? x : __uint_as_float(down);
}
// s = fl32(fl32(v * ri) * cj) for two adjacent columns (packed FMUL2; same two roundings everywhere)
__device__ __forceinline__ float2 score2(uint32_t v0, uint32_t v1, float2 ri2, float c0, float c1) {
    return __fmul2_rn(__fmul2_rn(make_float2(__uint_as_float(v0), __uint_as_float(v1)), ri2), make_float2(c0, c1));
}

// Vector-loader contract.  Each policy splits its per-column / per-row epilogue operands into
//   fetch_*: global loads ONLY (predicated, no arithmetic on loaded values) into raw[] registers, and
//   make_*:  the arithmetic that turns raw[] into the staged fp32 vectors.
// The loader warp calls every fetch of a tile before the first make, so all of a tile's loads are in flight
// together: one global round trip per tile instead of one per 32 columns.
__device__ __forceinline__ uint32_t ldu(const float* p, int64_t i, bool pred, uint32_t dflt) {
    return (pred && p) ? __float_as_uint(p[i]) : dflt;
}
constexpr uint32_t kOneBits = 0x3f800000u;
constexpr int kMaxColRaw = 2, kMaxRowRaw = 4;

// ------------------------------------------------------------------------------- epilogues
// Each policy: Params (POD, kernel argument), kColVecs (per-column fp32 vectors staged in smem),
// kStoresG (needs the OutStage), per-thread state as members.

struct StorePolicy {
    static constexpr bool kByteG = false;
    struct Params {
        float* out;
        int64_t ld;
    };
    static constexpr int kColVecs = 1;
    static constexpr bool kStoresG = false;
    static constexpr bool kStoresF32 = true;  // S tiles leave through the OutStage (fp32 boxes [32 x 32])
    float ri;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params&, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
    }
    static constexpr int kRowVecs = 1;
    __device__ static void fetch_row(const Params&, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
    }
    __device__ static void make_row(const Params&, const SimCommon& c, int64_t, bool valid, const uint32_t* raw, float* v) {
        v[0] = (valid ? __uint_as_float(raw[0]) : 0.f) * c.scale;
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx&, const float* rv) { ri = rv[0]; }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage& os) {
        if (cbase >= t.cols_valid) return;  // warp-uniform; TMA clips partially valid boxes itself
        const int lane = threadIdx.x & 31;
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = __fmul_rn(__fmul_rn(__uint_as_float(v[j]), ri), cv[j]);
        os.begin_slab(lane);
        os.write_f32(lane, o);
        os.end_slab(lane, (int32_t)(t.col0 + cbase), (int32_t)(t.row0 + t.quad * 32));
    }
    __device__ void tile_end(const Params&, const SimCommon&, const TileCtx&) {}
    __device__ void kernel_end(const Params&, float*, int) {}
};

struct RankPolicy {
    static constexpr bool kByteG = false;
    static constexpr bool kStoresF32 = false;
    struct Params {
        const float* pos_thr;  // rank_threshold(fl32(1 - s_pos)) per row (pb2_sim_diag / pb2_pair_dot)
        const int64_t* pos_col;
        int64_t col_offset;
        int32_t* rank;
    };
    static constexpr int kColVecs = 1;
    static constexpr bool kStoresG = false;
    float ri, thr;
    int pc;  // positive's column relative to the tile origin (may be out of range)
    int cnt;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params&, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        // an out-of-range column must never count: NaN makes every comparison false
        v[0] = valid ? __uint_as_float(raw[0]) : PB2_NAN;
    }
    static constexpr int kRowVecs = 3;  // rinv_x * scale, rank threshold, positive's column (int bits)
    __device__ static void fetch_row(const Params& p, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
        raw[1] = ldu(p.pos_thr, row, valid, 0u);
        const long long pc = valid ? p.pos_col[row] : -1ll;
        raw[2] = (uint32_t)(unsigned long long)pc;
        raw[3] = (uint32_t)((unsigned long long)pc >> 32);
    }
    __device__ static void make_row(const Params& p, const SimCommon& c, int64_t, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) * c.scale : 0.f;
        v[1] = valid ? __uint_as_float(raw[1]) : PB2_INF;  // s >= thr  <=>  fl32(1 - s) < fl32(1 - s_pos)
        const int64_t pc = (int64_t)(((unsigned long long)raw[3] << 32) | raw[2]);
        const int64_t rel = valid ? pc - p.col_offset : -1;
        v[2] = __int_as_float((rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1);
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx& t, const float* rv) {
        cnt = 0;
        ri = rv[0];
        thr = rv[128];
        const int g = __float_as_int(rv[256]);
        const int64_t rel = (int64_t)g - t.col0;
        pc = (g >= 0 && rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1;
    }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx&, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage&) {
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float2 ri2 = make_float2(ri, ri);
        const int rel = pc - cbase;  // the positive itself never counts
        int c = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 c4 = cv4[q];
            const float2 s01 = score2(v[4 * q], v[4 * q + 1], ri2, c4.x, c4.y);
            const float2 s23 = score2(v[4 * q + 2], v[4 * q + 3], ri2, c4.z, c4.w);
            c += (s01.x >= thr) ? 1 : 0;
            c += (s01.y >= thr) ? 1 : 0;
            c += (s23.x >= thr) ? 1 : 0;
            c += (s23.y >= thr) ? 1 : 0;
        }
        if (rel >= 0 && rel < 32) {  // take the positive's own vote back
            float sv = 0.f, cj = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j == rel) {
                    sv = __uint_as_float(v[j]);
                    cj = cv[j];
                }
            c -= (__fmul_rn(__fmul_rn(sv, ri), cj) >= thr) ? 1 : 0;
        }
        cnt += c;
    }
    __device__ void tile_end(const Params& p, const SimCommon&, const TileCtx& t) {
        if (t.row_valid && cnt) atomicAdd(p.rank + t.row, cnt);
    }
    __device__ void kernel_end(const Params&, float*, int) {}
};

// Paired scores s_k = <x_k, y_k> computed BY THE SAME tensor-core arithmetic as the full passes (only
// the diagonal tiles are visited): a gallery row that duplicates the positive then scores exactly
// like it, as in the reference where both come out of one GEMM, and "strictly closer" stays strict.
struct DiagPolicy {
    static constexpr bool kByteG = false;
    static constexpr bool kStoresF32 = false;
    struct Params {
        float* out;   // s_k            (may be null)
        float* dist;  // fl32(1 - s_k)  (may be null)
        float* thr;   // rank_threshold(fl32(1 - s_k)) (may be null)
    };
    static constexpr int kColVecs = 1;
    static constexpr bool kStoresG = false;
    float ri;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params&, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
    }
    static constexpr int kRowVecs = 1;
    __device__ static void fetch_row(const Params&, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
    }
    __device__ static void make_row(const Params&, const SimCommon& c, int64_t, bool valid, const uint32_t* raw, float* v) {
        v[0] = (valid ? __uint_as_float(raw[0]) : 0.f) * c.scale;
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx&, const float* rv) { ri = rv[0]; }
    __device__ void chunk(const Params& p, const SimCommon&, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage&) {
        // tile (rb, rb) with BN == BM: row quad*32 + lane pairs with column quad*32 + lane
        if (cbase != t.quad * 32 || !t.row_valid) return;
        const int lane = lane_id();
        float sv = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j == lane) sv = __uint_as_float(v[j]);
        const float s = __fmul_rn(__fmul_rn(sv, ri), cv[lane]);
        const float d = __fsub_rn(1.0f, s);
        if (p.out) p.out[t.row] = s;
        if (p.dist) p.dist[t.row] = d;
        if (p.thr) p.thr[t.row] = rank_threshold(d);
    }
    __device__ void tile_end(const Params&, const SimCommon&, const TileCtx&) {}
    __device__ void kernel_end(const Params&, float*, int) {}
};

struct HingeParams {
    const float* diag_row;
    const float* diag_col;
    int64_t row_offset, col_offset;
    float margin;
    float* loss_partial;
    int32_t* row_cnt;
    int32_t* col_cnt;
    int has_gmat;
    const float* pos_thr;   // kRank only: rank threshold of the diagonal score
    int32_t* rank;          // kRank only
};

// kRank additionally counts, per row, the columns closer than the diagonal (recall@k of the same
// similarity pass: the gallery workload shares one S pass between pig.loss and pig.metrics).
//
// Instruction budget: the MMA of a 128 x 256 x 512 tile takes ~4096 cycles = ~16 issue slots per
// element, split over an FMA pipe and an ALU pipe that each accept one warp instruction every
// other cycle.  So per element:
//   FMA pipe  score (2 packed FMUL2 per pair), the two hinge indicators as floats with a saturating
//             FFMA ( sat((s - pred(thr)) * 2^120) is exactly [s >= thr] ), g = ic + ir, loss += g * s,
//             row count += ir (packed FADD2 / FFMA2);
//   ALU pipe  column-count compare + byte-packed add, rank compare + add, fp16 pack.
// There is no per-element z = margin + s - d: fl32(s + c) >= 0  <=>  s >= -c exactly, and
// sum relu(z) = sum g*s + sum_j (m - d_j) col_cnt[j] + sum_i (m - d_i) row_cnt[i] is completed from the
// counts afterwards (pb2_hinge_loss_terms).  Column counts are summed over the warp's rows with REDUX.
#ifndef PB2_HINGE_KO
#define PB2_HINGE_KO 0  // knock-out timing builds (wrong results): 1 no column counts, 2 no loss sum, 4 no row counts
#endif
#ifndef PB2_HINGE_PIPES
#define PB2_HINGE_PIPES 0  // measurement builds: bit 0 column indicator on the ALU pipe, bits 1 / 2 row / rank indicator on the FMA pipe
#endif
constexpr float kBig = 1.329227995784916e36f;  // 2^120
// kByteG: the gradient matrix leaves as ONE BYTE per entry (its values are exactly {0, 1, 2}) for the kind::i8
// gradient GEMMs (gradgemm.cu): half the bytes written here and read there twice.
template <bool kRank, bool kByteG_ = false>
struct HingePolicyT {
    static constexpr bool kByteG = kByteG_;
    static constexpr bool kStoresF32 = false;
    using Params = HingeParams;
    static constexpr int kColVecs = 2;  // rinv_y, -pred(thr_c) * 2^120 with thr_c = diag_col - margin
    static constexpr bool kStoresG = true;
    float ri, thr_r, thr_k, cr_r, cr_k;
    float2 loss2, rc2, rk2;
    int dcol;
    __device__ void kernel_begin(const Params&) { loss2 = make_float2(0.f, 0.f); }
    __device__ static void fetch_col(const Params& p, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
        raw[1] = ldu(p.diag_col, col, valid, 0u);
    }
    __device__ static void make_col(const Params& p, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
#if PB2_HINGE_PIPES & 1
        v[1] = valid ? __uint_as_float(raw[1]) - p.margin : PB2_INF;
#else
        v[1] = valid ? -(pred_f32(__uint_as_float(raw[1]) - p.margin) * kBig) : -PB2_INF;
#endif
    }
    static constexpr int kRowVecs = 4;  // rinv_x, thr_r = diag_row - margin, pos_thr, diagonal column
    __device__ static void fetch_row(const Params& p, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
        raw[1] = ldu(p.diag_row, row, valid, 0u);
        raw[2] = ldu(p.pos_thr, row, kRank && valid, 0u);
    }
    __device__ static void make_row(const Params& p, const SimCommon&, int64_t row, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
        // per-row thresholds are compared on the ALU pipe (fset_ge); an invalid row never fires
        v[1] = valid ? __uint_as_float(raw[1]) - p.margin : PB2_INF;
        v[2] = (kRank && valid) ? __uint_as_float(raw[2]) : PB2_INF;
        const int64_t rel = valid ? (p.row_offset + row) - p.col_offset : -1;
        v[3] = __int_as_float((rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1);
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx& t, const float* rv) {
        rc2 = make_float2(0.f, 0.f);
        rk2 = make_float2(0.f, 0.f);
        ri = rv[0];
        thr_r = rv[128];
        thr_k = rv[256];
#if PB2_HINGE_PIPES & 6
        cr_r = -(pred_f32(thr_r) * kBig);
        cr_k = -(pred_f32(thr_k) * kBig);
#endif
        const int g = __float_as_int(rv[384]);
        const int64_t rel = (int64_t)g - t.col0;
        dcol = (g >= 0 && rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1;
    }
    // kSlow: chunks containing the diagonal, out-of-range columns or out-of-range rows; those elements
    // get the masked score.
    template <bool kSlow>
    __device__ __forceinline__ void chunk_impl(const Params& p, const TileCtx& t, int ch, int cbase,
                                               const uint32_t (&v)[32], const float* cv, OutStage& os) {
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float4* cb4 = reinterpret_cast<const float4*>(cv + kColVecStride);
        const float2 ri2 = make_float2(ri, ri);
        const int drel = dcol - cbase;                              // diagonal position inside this chunk
        const int nvalid = t.row_valid ? t.cols_valid - cbase : 0;  // valid columns of this row's chunk
        const int lane = threadIdx.x & 31;
        uint32_t packed[kByteG ? 8 : 16];
        uint32_t mine = 0;
        // independent accumulators: no serial dependency chain longer than 8 per chunk
        float2 la = make_float2(0.f, 0.f), lb = make_float2(0.f, 0.f);
        float2 rka = make_float2(0.f, 0.f), rkb = make_float2(0.f, 0.f);
        float2 rca = make_float2(0.f, 0.f), rcb = make_float2(0.f, 0.f);
        // software pipelining: the three column-vector loads of group q+1 and the REDUX of group q are in
        // flight while group q's arithmetic issues (their latencies were the top stall reasons in ncu)
        float4 c4 = cv4[0], b4 = cb4[0];
        uint32_t tot[8], pk_prev = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float4 nc4 = c4, nb4 = b4;
            if (q < 7) {
                nc4 = cv4[q + 1];
                nb4 = cb4[q + 1];
            }
            float2 s01 = score2(v[4 * q], v[4 * q + 1], ri2, c4.x, c4.y);
            float2 s23 = score2(v[4 * q + 2], v[4 * q + 3], ri2, c4.z, c4.w);
            if (kSlow) {
                if (4 * q + 0 == drel || 4 * q + 0 >= nvalid) s01.x = kMasked;
                if (4 * q + 1 == drel || 4 * q + 1 >= nvalid) s01.y = kMasked;
                if (4 * q + 2 == drel || 4 * q + 2 >= nvalid) s23.x = kMasked;
                if (4 * q + 3 == drel || 4 * q + 3 >= nvalid) s23.y = kMasked;
            }
            // FMA pipe: indicators as exact 0/1 floats
#if PB2_HINGE_PIPES & 1
            const float2 ic01 = make_float2(fset_ge(s01.x, b4.x), fset_ge(s01.y, b4.y));
            const float2 ic23 = make_float2(fset_ge(s23.x, b4.z), fset_ge(s23.y, b4.w));
#else
            const float2 ic01 = make_float2(fma_sat(s01.x, kBig, b4.x), fma_sat(s01.y, kBig, b4.y));
            const float2 ic23 = make_float2(fma_sat(s23.x, kBig, b4.z), fma_sat(s23.y, kBig, b4.w));
#endif
#if PB2_HINGE_PIPES & 2
            const float2 ir01 = make_float2(fma_sat(s01.x, kBig, cr_r), fma_sat(s01.y, kBig, cr_r));
            const float2 ir23 = make_float2(fma_sat(s23.x, kBig, cr_r), fma_sat(s23.y, kBig, cr_r));
#else
            const float2 ir01 = make_float2(fset_ge(s01.x, thr_r), fset_ge(s01.y, thr_r));  // ALU pipe
            const float2 ir23 = make_float2(fset_ge(s23.x, thr_r), fset_ge(s23.y, thr_r));
#endif
            const float2 g01 = __fadd2_rn(ic01, ir01), g23 = __fadd2_rn(ic23, ir23);
#if !(PB2_HINGE_KO & 2)
            la = __ffma2_rn(g01, s01, la);
            lb = __ffma2_rn(g23, s23, lb);
#endif
#if !(PB2_HINGE_KO & 4)
            rca = __fadd2_rn(rca, ir01);
            rcb = __fadd2_rn(rcb, ir23);
#endif
            if (kRank) {
#if PB2_HINGE_PIPES & 4
                rka = __fadd2_rn(rka, make_float2(fma_sat(s01.x, kBig, cr_k), fma_sat(s01.y, kBig, cr_k)));
                rkb = __fadd2_rn(rkb, make_float2(fma_sat(s23.x, kBig, cr_k), fma_sat(s23.y, kBig, cr_k)));
#else
                rka = __fadd2_rn(rka, make_float2(fset_ge(s01.x, thr_k), fset_ge(s01.y, thr_k)));
                rkb = __fadd2_rn(rkb, make_float2(fset_ge(s23.x, thr_k), fset_ge(s23.y, thr_k)));
#endif
            }
            // column counts: the four 0/1 indicators packed into 6-bit fields of one exact fp32 integer
            // (ic0 + 64 ic1 + 4096 ic2 + 262144 ic3 < 2^19), converted once and summed over the warp's
            // 32 rows with REDUX (<= 32 per field); lanes 4q..4q+3 keep group q
            // the conversion of group q is in flight while group q-1 is reduced (F2I and REDUX latencies overlap the
            // next group's arithmetic); the totals are warp-uniform and picked per lane after the loop
#if PB2_HINGE_KO & 1
            if (q > 0) tot[q - 1] = 0;
#else
            const uint32_t pkq = __float2uint_rz(fmaf(fmaf(ic23.y, 64.f, ic23.x), 4096.f, fmaf(ic01.y, 64.f, ic01.x)));
            if (q > 0) tot[q - 1] = __reduce_add_sync(0xffffffffu, pk_prev);
            pk_prev = pkq;
#endif
            if constexpr (kByteG) {
                // four entries in {0, 1, 2} -> four bytes: t = g0 + 256 g1 and u = g2 + 256 g3 are exact small integers;
                // adding 2^23 parks them in the low mantissa bits, one PRMT joins the two 16-bit halves
                const float2 tu = __fadd2_rn(make_float2(fmaf(g01.y, 256.f, g01.x), fmaf(g23.y, 256.f, g23.x)),
                                             make_float2(8388608.f, 8388608.f));
                packed[q] = __byte_perm(__float_as_uint(tu.x), __float_as_uint(tu.y), 0x5410);
            } else {
                const __half2 h01 = __float22half2_rn(g01), h23 = __float22half2_rn(g23);
                packed[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
                packed[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            }
            c4 = nc4;
            b4 = nb4;
        }
#if PB2_HINGE_KO & 1
        tot[7] = pk_prev;
#else
        tot[7] = __reduce_add_sync(0xffffffffu, pk_prev);
#endif
        {   // lane L keeps group L >> 2: a three-level select on the lane's bits 2..4 (7 SEL, no per-group predicates)
            const bool b0 = lane & 4, b1 = lane & 8, b2 = lane & 16;
            const uint32_t a0 = b0 ? tot[1] : tot[0], a1 = b0 ? tot[3] : tot[2];
            const uint32_t a2 = b0 ? tot[5] : tot[4], a3 = b0 ? tot[7] : tot[6];
            const uint32_t e0 = b1 ? a1 : a0, e1 = b1 ? a3 : a2;
            mine = b2 ? e1 : e0;
        }
        loss2 = __fadd2_rn(loss2, __fadd2_rn(la, lb));
        rc2 = __fadd2_rn(rc2, __fadd2_rn(rca, rcb));
        if (kRank) rk2 = __fadd2_rn(rk2, __fadd2_rn(rka, rkb));
        const int ccnt = (int)((mine >> ((lane & 3) * 6)) & 0x3fu);
        if (ccnt) atomicAdd(p.col_cnt + t.col0 + cbase + lane, ccnt);  // 0 for out-of-range columns
        if (p.has_gmat) {
            if constexpr (kByteG) os.stage_u8(lane, ch, packed, (int32_t)(t.col0 + cbase), (int32_t)(t.row0 + t.quad * 32));
            else os.stage(lane, ch, packed, (int32_t)(t.col0 + cbase), (int32_t)(t.row0 + t.quad * 32));
        }
    }
    __device__ void chunk(const Params& p, const SimCommon&, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage& os) {
        const int drel = dcol - cbase;
        const bool slow = (drel >= 0 && drel < 32) || !t.row_valid || (t.cols_valid - cbase) < 32;
        if (__any_sync(0xffffffffu, slow)) chunk_impl<true>(p, t, ch, cbase, v, cv, os);
        else chunk_impl<false>(p, t, ch, cbase, v, cv, os);
    }
    __device__ void tile_end(const Params& p, const SimCommon&, const TileCtx& t) {
        const int rcnt = (int)(rc2.x + rc2.y);  // exact: small integers in fp32
        if (t.row_valid && rcnt) atomicAdd(p.row_cnt + t.row, rcnt);
        const int rk = (int)(rk2.x + rk2.y);
        if (kRank && t.row_valid && rk) atomicAdd(p.rank + t.row, rk);
    }
    __device__ void kernel_end(const Params& p, float* red, int n_epi_warps) {
        // fixed-order reduction over the epilogue threads -> one deterministic partial per CTA
        const int e = threadIdx.x - kEpiWarp0 * 32;
        const float w = warp_sum(loss2.x + loss2.y);
        if ((e & 31) == 0) red[e >> 5] = w;
        named_bar_sync(2, n_epi_warps * 32);
        if (e == 0) {
            float s = 0.f;
            for (int i = 0; i < n_epi_warps; ++i) s += red[i];
            p.loss_partial[blockIdx.x] = s;
        }
    }
};

struct LseRowPolicy {
    static constexpr bool kByteG = false;
    static constexpr bool kStoresF32 = false;
    struct Params {
        float* part_max;
        float* part_sum;
    };
    static constexpr int kColVecs = 1;
    static constexpr bool kStoresG = false;
    float ri, m, s;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params&, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
    }
    static constexpr int kRowVecs = 1;
    __device__ static void fetch_row(const Params&, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
    }
    __device__ static void make_row(const Params&, const SimCommon& c, int64_t, bool valid, const uint32_t* raw, float* v) {
        // work in the log2 domain: t = s_ij * log2(e)
        v[0] = (valid ? __uint_as_float(raw[0]) : 0.f) * c.scale * 1.4426950408889634f;
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx&, const float* rv) {
        ri = rv[0];
        m = -PB2_INF;
        s = 0.f;
    }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage&) {
        const int nvalid = t.cols_valid - cbase;
        if (nvalid <= 0) return;
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float2 ri2 = make_float2(ri, ri);
        float2 x[16];
        // logits of the chunk (packed FMUL2) and their maximum as a pairwise tree
        float2 mx = make_float2(-PB2_INF, -PB2_INF);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 c4 = cv4[q];
            x[2 * q] = score2(v[4 * q], v[4 * q + 1], ri2, c4.x, c4.y);
            x[2 * q + 1] = score2(v[4 * q + 2], v[4 * q + 3], ri2, c4.z, c4.w);
            if (nvalid < 32) {  // warp-uniform: only the last column tile has a ragged chunk
                if (4 * q + 0 >= nvalid) x[2 * q].x = -PB2_INF;
                if (4 * q + 1 >= nvalid) x[2 * q].y = -PB2_INF;
                if (4 * q + 2 >= nvalid) x[2 * q + 1].x = -PB2_INF;
                if (4 * q + 3 >= nvalid) x[2 * q + 1].y = -PB2_INF;
            }
            mx.x = fmaxf(mx.x, fmaxf(x[2 * q].x, x[2 * q + 1].x));
            mx.y = fmaxf(mx.y, fmaxf(x[2 * q].y, x[2 * q + 1].y));
        }
        const float mn = fmaxf(m, fmaxf(mx.x, mx.y));
        // mn == -inf only if every logit so far is -inf; keep the state untouched then
        if (mn > -PB2_INF) {
            const float2 neg = make_float2(-mn, -mn);
            float a0 = s * ex2_approx(m - mn), a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const float2 d0 = __fadd2_rn(x[k], neg), d1 = __fadd2_rn(x[k + 1], neg);
                a0 += ex2_approx(d0.x);
                a1 += ex2_approx(d0.y);
                a2 += ex2_approx(d1.x);
                a3 += ex2_approx(d1.y);
            }
            s = (a0 + a1) + (a2 + a3);
            m = mn;
        }
    }
    __device__ void tile_end(const Params& p, const SimCommon& c, const TileCtx& t) {
        if (!t.row_valid) return;
        // partial layout: two slots per 128 columns (pb2_sim_lse_parts).  A warp of a 128-wide tile covers 64
        // columns = one slot; a warp of a 256-wide tile covers 128 columns = a slot pair (second one empty).
        if (t.warp_cols == 64) {
            const int64_t slot = ((int64_t)t.cb * 2 + t.half) * c.rows + t.row;
            p.part_max[slot] = m;
            p.part_sum[slot] = s;
        } else if (((int64_t)t.cb * 2 + t.half) * 128 < c.cols) {  // a 128-column unit past the last column has no slots
            const int64_t slot = ((int64_t)t.cb * 4 + t.half * 2) * c.rows + t.row;
            p.part_max[slot] = m;
            p.part_sum[slot] = s;
            p.part_max[slot + c.rows] = -PB2_INF;
            p.part_sum[slot + c.rows] = 0.f;
        }
    }
    __device__ void kernel_end(const Params&, float*, int) {}
};

// Row AND column log-sum-exp statistics from ONE pass over S (north-star kernel (a)), for logits with a known bound:
// with |scale * <x_i, y_j>| <= bound, e_ij = 2^(t_ij - M) (t = logit * log2 e, M = bound * log2 e) never overflows and
// keeps full relative precision while 2 M stays far below the fp32 exponent range (the host refuses M > 60), so the
// row sums and the column sums are plain sums of the SAME exponentials: one ex2 per element instead of two passes with
// running maxima.  A thread owns a row, so its row sum is a register; the column sums of a 32 x 32 chunk go through
// the warp's staging slab (one swizzled 128-byte row per thread, read back eight rows x four columns per thread,
// two shuffles) and leave as one fp32 per (32-row group, column): fixed order, no atomics, bit-reproducible.
// Partials: row sums in the layout of pb2_sim_lse_parts (per 128 columns), column sums [4 * row blocks, cols];
// both are merged by pb2_lse_merge_const.
// kRank: the same pass also counts, per row, the columns whose COSINE beats the positive's (recall@k of the gallery,
// pig/metrics.py:23-40) -- the loss statistics use the caller's logits (raw dot products over a temperature for
// pig.loss.MILNCELoss), the ranking re-normalises like pig.util.cosine_matrix, so the count takes its own 1 / ||row||
// vectors and scores every element a second time from the same accumulator: s = fl32(fl32(acc * ra_i) * rv_j), the
// very expression (and bits) of RankPolicy, compared with the threshold of pb2_sim_diag.
struct LseBothParams {
    float* row_part_sum;  // [pb2_sim_lse_parts(cols), rows]
    float* col_part_sum;  // [pb2_sim_lse_col_parts(rows), cols]
    float shift;          // M: log2-domain upper bound of the logits
    // kRank only
    const float* rank_rinv_x;  // 1 / ||x_i|| of the ranking (may be null = 1)
    const float* rank_rinv_y;  // 1 / ||y_j||
    const float* pos_thr;      // rank_threshold(fl32(1 - s_pos)) per row
    int64_t row_offset, col_offset;  // the positive of row i is global column row_offset + i
    int32_t* rank;
};
template <bool kRank>
struct LseBothPolicyT {
    static constexpr bool kByteG = false;
    static constexpr bool kStoresF32 = false;
    static constexpr bool kStoresG = false;
    static constexpr bool kUsesStage = true;
    using Params = LseBothParams;
    static constexpr int kColVecs = kRank ? 2 : 1;  // rinv_y of the logits; rinv_y of the ranking
    float ri, s;
    float rri, thr_k;
    float2 rk2;
    int dcol;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params& p, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
        if (kRank) raw[1] = ldu(p.rank_rinv_y, col, valid, kOneBits);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
        if (kRank) v[1] = valid ? __uint_as_float(raw[1]) : PB2_NAN;  // an out-of-range column never counts
    }
    static constexpr int kRowVecs = kRank ? 4 : 1;  // logit factor; ranking 1 / ||x||, rank threshold, positive's column
    __device__ static void fetch_row(const Params& p, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
        if (kRank) {
            raw[1] = ldu(p.rank_rinv_x, row, valid, kOneBits);
            raw[2] = ldu(p.pos_thr, row, valid, 0u);
        }
    }
    __device__ static void make_row(const Params& p, const SimCommon& c, int64_t row, bool valid, const uint32_t* raw, float* v) {
        v[0] = (valid ? __uint_as_float(raw[0]) : 0.f) * c.scale * 1.4426950408889634f;  // log2 domain
        if (kRank) {
            v[1] = valid ? __uint_as_float(raw[1]) : 0.f;
            v[2] = valid ? __uint_as_float(raw[2]) : PB2_INF;  // an out-of-range row never counts
            const int64_t rel = valid ? (p.row_offset + row) - p.col_offset : -1;
            v[3] = __int_as_float((rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1);
        }
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx& t, const float* rv) {
        ri = rv[0];
        s = 0.f;
        if (kRank) {
            rri = rv[128];
            thr_k = rv[256];
            rk2 = make_float2(0.f, 0.f);
            const int g = __float_as_int(rv[384]);
            const int64_t rel = (int64_t)g - t.col0;
            dcol = (g >= 0 && rel >= 0 && rel < 0x7fffffff) ? (int)rel : -1;
        }
    }
    __device__ void chunk(const Params& p, const SimCommon& c, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage& os) {
        const int nvalid = t.cols_valid - cbase;
        if (nvalid <= 0) return;  // warp-uniform
        // kSlow: ragged last column tile, or rows past the end in this warp (those elements add nothing); with kRank
        // also a chunk that holds some row's positive (it never counts against itself)
        bool slow = nvalid < 32 || !t.row_valid;
        if (kRank) slow = slow || (dcol - cbase >= 0 && dcol - cbase < 32);
        if (__any_sync(0xffffffffu, slow)) chunk_impl<true>(p, c, t, cbase, v, cv, os);
        else chunk_impl<false>(p, c, t, cbase, v, cv, os);
    }
    template <bool kSlow>
    __device__ __forceinline__ void chunk_impl(const Params& p, const SimCommon& c, const TileCtx& t, int cbase,
                                               const uint32_t (&v)[32], const float* cv, OutStage& os) {
        const int nvalid = t.cols_valid - cbase;
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float2 ri2 = make_float2(ri, ri);
        const float2 neg = make_float2(-p.shift, -p.shift);
        float e[32];
        float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
        [[maybe_unused]] const float4* ck4 = reinterpret_cast<const float4*>(cv + kColVecStride);
        [[maybe_unused]] const float2 rri2 = make_float2(rri, rri);
        [[maybe_unused]] const int drel = dcol - cbase;
        [[maybe_unused]] float2 rka = make_float2(0.f, 0.f), rkb = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 c4 = cv4[q];
            const float2 d0 = __fadd2_rn(score2(v[4 * q], v[4 * q + 1], ri2, c4.x, c4.y), neg);
            const float2 d1 = __fadd2_rn(score2(v[4 * q + 2], v[4 * q + 3], ri2, c4.z, c4.w), neg);
            if constexpr (kRank) {  // cosine of the same accumulator entries against the positive's threshold (ALU pipe)
                const float4 k4 = ck4[q];
                const float2 r01 = score2(v[4 * q], v[4 * q + 1], rri2, k4.x, k4.y);
                const float2 r23 = score2(v[4 * q + 2], v[4 * q + 3], rri2, k4.z, k4.w);
                float2 i01 = make_float2(fset_ge(r01.x, thr_k), fset_ge(r01.y, thr_k));
                float2 i23 = make_float2(fset_ge(r23.x, thr_k), fset_ge(r23.y, thr_k));
                if (kSlow) {
                    if (4 * q + 0 == drel) i01.x = 0.f;
                    if (4 * q + 1 == drel) i01.y = 0.f;
                    if (4 * q + 2 == drel) i23.x = 0.f;
                    if (4 * q + 3 == drel) i23.y = 0.f;
                }
                rka = __fadd2_rn(rka, i01);
                rkb = __fadd2_rn(rkb, i23);
            }
            e[4 * q + 0] = ex2_approx(d0.x);
            e[4 * q + 1] = ex2_approx(d0.y);
            e[4 * q + 2] = ex2_approx(d1.x);
            e[4 * q + 3] = ex2_approx(d1.y);
            if (kSlow) {
                if (4 * q + 0 >= nvalid || !t.row_valid) e[4 * q + 0] = 0.f;
                if (4 * q + 1 >= nvalid || !t.row_valid) e[4 * q + 1] = 0.f;
                if (4 * q + 2 >= nvalid || !t.row_valid) e[4 * q + 2] = 0.f;
                if (4 * q + 3 >= nvalid || !t.row_valid) e[4 * q + 3] = 0.f;
            }
            a01 = __fadd2_rn(a01, make_float2(e[4 * q + 0], e[4 * q + 1]));
            a23 = __fadd2_rn(a23, make_float2(e[4 * q + 2], e[4 * q + 3]));
        }
        s += (a01.x + a01.y) + (a23.x + a23.y);
        if (kRank) rk2 = __fadd2_rn(rk2, __fadd2_rn(rka, rkb));
        // column sums over this warp's 32 rows: transpose through the slab
        const int lane = lane_id();
        os.write_f32(lane, e);
        __syncwarp();
        const uint8_t* sl = os.buf + (os.slab & os.mask) * kOutSlabBytes;
        const int c16 = lane & 7, rg = lane >> 3;
        float2 axy = make_float2(0.f, 0.f), azw = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // rows rg*8 + k (row % 8 == k), columns 4*c16 .. 4*c16+3
            const float4 w = *reinterpret_cast<const float4*>(sl + (rg * 8 + k) * 128 + ((c16 ^ k) * 16));
            axy = __fadd2_rn(axy, make_float2(w.x, w.y));
            azw = __fadd2_rn(azw, make_float2(w.z, w.w));
        }
        float4 acc = make_float4(axy.x, axy.y, azw.x, azw.y);
#pragma unroll
        for (int off = 8; off <= 16; off <<= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
            acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
            acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
        }
        const float mine = rg == 0 ? acc.x : (rg == 1 ? acc.y : (rg == 2 ? acc.z : acc.w));
        const int col = c16 * 4 + rg;  // the 32 lanes cover the chunk's 32 columns once
        if (!kSlow || col < nvalid) {
            const int64_t part = (t.row0 >> 7) * 4 + t.quad;
            p.col_part_sum[part * c.cols + t.col0 + cbase + col] = mine;
        }
        ++os.slab;  // the next chunk stages into the other slab while stragglers still read this one
    }
    __device__ void tile_end(const Params& p, const SimCommon& c, const TileCtx& t) {
        if (!t.row_valid) return;
        if (kRank) {
            const int rk = (int)(rk2.x + rk2.y);  // exact: small integers in fp32
            if (rk) atomicAdd(p.rank + t.row, rk);
        }
        // row-sum partials: two slots per 128 columns like LseRowPolicy (pb2_sim_lse_parts)
        if (t.warp_cols == 64) {
            p.row_part_sum[((int64_t)t.cb * 2 + t.half) * c.rows + t.row] = s;
        } else if (((int64_t)t.cb * 2 + t.half) * 128 < c.cols) {
            const int64_t slot = ((int64_t)t.cb * 4 + t.half * 2) * c.rows + t.row;
            p.row_part_sum[slot] = s;
            p.row_part_sum[slot + c.rows] = 0.f;
        }
    }
    __device__ void kernel_end(const Params&, float*, int) {}
};

struct LseGradPolicy {
    static constexpr bool kByteG = false;
    static constexpr bool kStoresF32 = false;
    struct Params {
        const float* den_row;
        const float* den_col;
    };
    static constexpr int kColVecs = 2;  // rinv_y, 13 - den_col * log2e
    static constexpr bool kStoresG = true;
    float ri, drow;
    __device__ void kernel_begin(const Params&) {}
    __device__ static void fetch_col(const Params& p, const SimCommon& c, int64_t col, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_y, col, valid, kOneBits);
        raw[1] = ldu(p.den_col, col, valid, 0u);
    }
    __device__ static void make_col(const Params&, const SimCommon&, bool valid, const uint32_t* raw, float* v) {
        v[0] = valid ? __uint_as_float(raw[0]) : 0.f;
        v[1] = valid ? (13.0f - __uint_as_float(raw[1]) * 1.4426950408889634f) : -PB2_INF;
    }
    static constexpr int kRowVecs = 2;
    __device__ static void fetch_row(const Params& p, const SimCommon& c, int64_t row, bool valid, uint32_t* raw) {
        raw[0] = ldu(c.rinv_x, row, valid, kOneBits);
        raw[1] = ldu(p.den_row, row, valid, 0u);
    }
    __device__ static void make_row(const Params&, const SimCommon& c, int64_t, bool valid, const uint32_t* raw, float* v) {
        v[0] = (valid ? __uint_as_float(raw[0]) : 0.f) * c.scale * 1.4426950408889634f;
        v[1] = valid ? (13.0f - __uint_as_float(raw[1]) * 1.4426950408889634f) : -PB2_INF;
    }
    __device__ void tile_begin(const Params&, const SimCommon&, const TileCtx&, const float* rv) {
        ri = rv[0];
        drow = rv[128];
    }
    __device__ void chunk(const Params&, const SimCommon&, const TileCtx& t, int ch, int cbase, const uint32_t (&v)[32],
                          const float* cv, OutStage& os) {
        const float4* cv4 = reinterpret_cast<const float4*>(cv);
        const float4* cd4 = reinterpret_cast<const float4*>(cv + kColVecStride);
        const float2 ri2 = make_float2(ri, ri), dr2 = make_float2(drow, drow);
        uint32_t packed[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 c4 = cv4[q], d4 = cd4[q];
            const float2 x01 = score2(v[4 * q], v[4 * q + 1], ri2, c4.x, c4.y);
            const float2 x23 = score2(v[4 * q + 2], v[4 * q + 3], ri2, c4.z, c4.w);
            const float2 r01 = __fadd2_rn(x01, dr2), r23 = __fadd2_rn(x23, dr2);
            const float2 k01 = __fadd2_rn(x01, make_float2(d4.x, d4.y)), k23 = __fadd2_rn(x23, make_float2(d4.z, d4.w));
            const float2 g01 = __fadd2_rn(make_float2(ex2_approx(r01.x), ex2_approx(r01.y)),
                                          make_float2(ex2_approx(k01.x), ex2_approx(k01.y)));
            const float2 g23 = __fadd2_rn(make_float2(ex2_approx(r23.x), ex2_approx(r23.y)),
                                          make_float2(ex2_approx(k23.x), ex2_approx(k23.y)));
            const __half2 h01 = __float22half2_rn(g01), h23 = __float22half2_rn(g23);
            packed[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
            packed[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
        }
        const int lane = lane_id();
        os.stage(lane, ch, packed, (int32_t)(t.col0 + cbase), (int32_t)(t.row0 + t.quad * 32));
    }
    __device__ void tile_end(const Params&, const SimCommon&, const TileCtx&) {}
    __device__ void kernel_end(const Params&, float*, int) {}
};

// ---------------------------------------------------------------------------------- kernel
// kCtas: 1 = independent CTAs; 2 = CTA pairs sharing one MMA (cta_group::2); 3 = clusters of 2 whose CTAs take
// the tiles (rb, cb) and (rb + 1, cb), run independent MMAs and share the Y tile: each loads half of it and
// multicasts that half to both (a third fewer L2 lookups, no coupling through the accumulators); 4 = CTA pairs whose X
// strip stays resident in shared memory across a run of column tiles (measurement builds; see dispatch_sim).
// does the policy need the per-warp staging slabs (gradient-matrix / fp32 output tiles, or the column-sum transpose)?
template <class P, class = void>
struct uses_stage : std::integral_constant<bool, P::kStoresG || P::kStoresF32> {};
template <class P>
struct uses_stage<P, typename std::enable_if<P::kUsesStage>::type> : std::true_type {};

template <int BN, int G, bool kOut, int kCtas>
struct SimSmem {
    // kCtas == 4: CTA pairs whose X strip (128 rows x dim <= 512: eight k-block slots of 16 KiB) STAYS in shared memory
    // while the pair walks a run of column tiles; only Y streams through the ring
    static constexpr bool kResX = kCtas == 4;
    static constexpr int kXSlots = kResX ? 8 : 0;
    static constexpr int kXBytes = BM * BK * 2;
    static constexpr int kStageBytes = kResX ? (BN / 2) * BK * 2 : (BM + (kCtas == 2 ? BN / 2 : BN)) * BK * 2;  // a pair CTA stages half of Y
    // shared memory not spent on staging goes to the TMA pipeline: bytes in flight bound the MMA rate
    static constexpr int kOutBufs = (G >= 3 || kResX) ? 1 : 2;  // G >= 3: one 64-column slab per warp and tile
    static constexpr int kOutBytes = kOut ? 4 * G * kOutBufs * kOutSlabBytes : 0;
    static constexpr int kColVecBytes = 2 * kMaxColVecs * kColVecStride * 4 + 2 * kMaxRowVecs * BM * 4;  // col + row vectors
    static constexpr int kBarBytes = 512;
    static constexpr int kBudget = 227 * 1024 - kOutBytes - kColVecBytes - kBarBytes - kXSlots * kXBytes;
    static constexpr int kFit = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
#ifdef PB2_STAGE_CAP  // measurement builds: how much of the tile period is TMA bytes in flight?
    static constexpr int kStages = kFit > PB2_STAGE_CAP ? PB2_STAGE_CAP : kFit;
#else
    static constexpr int kStages = kFit;
#endif
    static constexpr int kTileBytes = kXSlots * kXBytes + kStages * kStageBytes;
    static constexpr int kTotal = kTileBytes + kOutBytes + kColVecBytes + kBarBytes;
    static_assert(kStages >= 2, "not enough shared memory for a pipeline");
};

// kCtas == 2: CTA pairs (cluster of 2, tcgen05 cta_group::2).  One MMA of M = 256 covers the 128 rows of
// each CTA; each CTA stages its own X rows and HALF of the tile's Y rows, accumulates S[its 128 rows, BN]
// in its own TMEM and runs its own loader and epilogue.  A third less operand traffic (L2 -> smem and
// smem -> tensor core) per SM than two independent 128 x BN tiles; see common.cuh for the protocol.
template <class Policy, int BN, int G, int kCtas>
__global__ void __launch_bounds__(sim_threads(G), 1)
    sim_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
               const __grid_constant__ CUtensorMap tm_out, const SimCommon c, const typename Policy::Params p) {
    using L = SimSmem<BN, G, uses_stage<Policy>::value, kCtas>;
    constexpr int kEpiWarps = 4 * G;
    constexpr int kEpiThreads = kEpiWarps * 32;
    constexpr int kTmemCols = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);  // power of two >= 2 * BN
    // 128-byte-swizzled TMA/UMMA tiles need 1024-byte alignment; the kernel has no static shared
    // memory, so the dynamic segment starts at the (aligned) base of the CTA's shared window.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* out_stage = smem + L::kTileBytes;
    float* colvec = reinterpret_cast<float*>(smem + L::kTileBytes + L::kOutBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kTileBytes + L::kOutBytes + L::kColVecBytes);
    uint64_t* full = bars;                       // [kStages]
    uint64_t* empty = bars + L::kStages;         // [kStages]
    uint64_t* acc_full = bars + 2 * L::kStages;  // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint64_t* vec_full = acc_empty + 2;          // [2]
    uint64_t* vec_empty = vec_full + 2;          // [2] the epilogue is done with a vector buffer (CTA-local)
    uint64_t* x_full = vec_empty + 2;            // [kXSlots] resident X strip (kCtas == 4)
    uint64_t* x_empty = x_full + L::kXSlots;     // [kXSlots]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_empty + L::kXSlots);
    float* rowvec = colvec + 2 * kMaxColVecs * kColVecStride;  // [acc stage][vec][128]
    float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [kMaxEpiWarps]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr int kLoaderWarp = 4 * G, kMmaWarp = 4 * G + 1, kTmaWarp = 4 * G + 2;
    constexpr int kCluster = kCtas >= 2 ? 2 : 1;    // CTAs per cluster
    constexpr bool kSharedMma = kCtas == 2 || kCtas == 4;  // one M = 256 MMA per pair
    constexpr bool kMcast = kCtas == 3;             // independent MMAs, Y tile multicast
    constexpr bool kResX = L::kResX;                // pairs with a resident X strip
    const uint32_t crank = kCluster == 2 ? cluster_ctarank() : 0u;  // pairs: 0 = leader (issues the MMAs)
    const int64_t unit0 = blockIdx.x / kCluster, n_units = gridDim.x / kCluster;  // tiles are dealt to CTAs / clusters
    // Tile walk.  Default: tile unit0, unit0 + n_units, ... of the banded order (concurrent CTAs share operands in L2).
    // Resident X: each pair takes ONE contiguous range of the row-block-major order, so that consecutive tiles keep
    // their row block (the X strip is loaded two or three times per launch instead of once per tile).
    const int64_t t_first = kResX ? (unit0 * c.n_tiles) / n_units : unit0;
    const int64_t t_last = kResX ? ((unit0 + 1) * c.n_tiles) / n_units : c.n_tiles;
    const int64_t t_step = kResX ? 1 : n_units;
    auto coords = [&](int64_t t, int& rb, int& cb) {
        if (kResX) {
            rb = (int)(t / c.n_cb);
            cb = (int)(t - (int64_t)rb * c.n_cb);
        } else {
            tile_coords(t, c.n_rb, c.n_cb, rb, cb);
        }
    };

    if (warp == kTmaWarp && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_y);
        if (Policy::kStoresG || Policy::kStoresF32) tma_prefetch_desc(&tm_out);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kMcast ? 2 : 1);  // multicast: a stage is refilled once BOTH CTAs' MMAs have read it
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(acc_full + a, 1);
            mbar_init(acc_empty + a, kEpiWarps * (kSharedMma ? 2 : 1));  // a pair leader's collects both CTAs' epilogues
            mbar_init(vec_full + a, 1);
            mbar_init(vec_empty + a, kEpiWarps);
        }
        for (int s = 0; s < L::kXSlots; ++s) {
            mbar_init(x_full + s, 1);
            mbar_init(x_empty + s, 1);
        }
        fence_mbar_init();
    }
    if (warp == kTmaWarp) {
        if (kSharedMma) tmem_alloc_pair(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
    }
    pdl_launch_dependents();
    tc_fence_before();
    if (kCluster == 2) cluster_sync_all();  // the peer's barriers exist before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Everything above overlapped the previous kernel's tail; global memory from here on.  With early operands (the
    // launch-bound training step only: 128-wide tiles of the fp16-G hinge pass; every other instantiation keeps the
    // plain wait and its code) the producer and MMA warps skip the wait: they only read X / Y (complete by the
    // caller's promise), shared memory and TMEM, so the main loop of the first tile runs beside the predecessor
    // (pb2_hinge_step: hinge_prep, whose 1/||row|| and diagonal only the loader and epilogue warps read -- behind
    // their own wait).
    constexpr bool kEarlyOk = std::is_same<Policy, HingePolicyT<false, false>>::value && BN == 128 && kCtas == 1;
    if constexpr (kEarlyOk) {
        if (!(c.early_operands && (warp == kTmaWarp || warp == kMmaWarp))) pdl_wait();
    } else {
        pdl_wait();
    }

    if (warp == kTmaWarp) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            [[maybe_unused]] uint32_t x_gen = 0;  // resident X: strips loaded so far
            for (int64_t t = t_first; t < t_last; t += t_step) {
                int rb, cb;
                coords(t, rb, cb);
                const int xrow = (rb * kCluster + (int)crank) * BM, yrow = cb * BN + (int)crank * (BN / kCluster);
                [[maybe_unused]] const bool new_x = kResX && (t == t_first || cb == 0);
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    if constexpr (kResX) {
                        if (new_x) {  // slot kb is free once the previous strip's last tile has consumed it
                            mbar_wait(x_empty + kb, (x_gen & 1u) ^ 1u);
                            if (crank == 0) mbar_arrive_expect_tx(x_full + kb, 2 * L::kXBytes);
                            tma_load_2d_pair(smem + kb * L::kXBytes, &tm_x, mapa_u32(smem_u32(x_full + kb), 0), kb * BK, xrow,
                                             PB2_OPERAND_POLICY);
                        }
                        mbar_wait(empty + stage, phase ^ 1);
                        if (crank == 0) mbar_arrive_expect_tx(full + stage, 2 * L::kStageBytes);
                        tma_load_2d_pair(smem + L::kXSlots * L::kXBytes + stage * L::kStageBytes, &tm_y,
                                         mapa_u32(smem_u32(full + stage), 0), kb * BK, yrow, PB2_OPERAND_POLICY);
                        if (++stage == L::kStages) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_wait(empty + stage, phase ^ 1);
                    uint8_t* sx = smem + stage * L::kStageBytes;
                    uint8_t* sy = sx + BM * BK * 2;
                    if (kCtas == 1) {
                        mbar_arrive_expect_tx(full + stage, L::kStageBytes);
                        tma_load_2d(sx, &tm_x, full + stage, kb * BK, xrow, PB2_OPERAND_POLICY);
                        tma_load_2d(sy, &tm_y, full + stage, kb * BK, yrow, PB2_OPERAND_POLICY);
                    } else if (kMcast) {  // own X tile; this CTA's half of the Y tile lands in both CTAs
                        mbar_arrive_expect_tx(full + stage, L::kStageBytes);  // X + both halves of Y arrive here
                        tma_load_2d(sx, &tm_x, full + stage, kb * BK, xrow, PB2_OPERAND_POLICY);
                        tma_load_2d_mcast(sy + crank * (BN / 2) * BK * 2, &tm_y, full + stage, kb * BK, yrow, (uint16_t)3,
                                          PB2_OPERAND_POLICY);
                    } else {  // both CTAs' bytes are counted on the leader's barrier
                        if (crank == 0) mbar_arrive_expect_tx(full + stage, L::kStageBytes * kCtas);
                        const uint32_t lbar = mapa_u32(smem_u32(full + stage), 0);
                        tma_load_2d_pair(sx, &tm_x, lbar, kb * BK, xrow, PB2_OPERAND_POLICY);
                        tma_load_2d_pair(sy, &tm_y, lbar, kb * BK, yrow, PB2_OPERAND_POLICY);
                    }
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (kResX && new_x) ++x_gen;
            }
        }
    } else if (warp == kMmaWarp) {
        // ====================================================================== MMA issuer
        if (lane == 0 && (crank == 0 || !kSharedMma)) {
            const uint32_t idesc = make_idesc(BM * (kSharedMma ? 2 : 1), BN, c.fmt, c.fmt, kMajorK, kMajorK);
            // K-major 128B-swizzled operands: descriptor = {start >> 4 | LBO 16 B, SBO 1024 B | version | swizzle}.
            // Only the start address changes, linearly: one running low word, adds instead of rebuilds.
            const uint64_t d0 = make_smem_desc(smem_u32(smem), 16, 1024);
            const uint32_t desc_hi = (uint32_t)(d0 >> 32), lo0 = (uint32_t)d0;
            constexpr uint32_t kStageLo = L::kStageBytes >> 4, kYLo = (BM * BK * 2) >> 4, kKLo = (UK * 2) >> 4;
            [[maybe_unused]] constexpr uint32_t kXLo = L::kXBytes >> 4, kRingLo = (L::kXSlots * L::kXBytes) >> 4;
            int stage = 0;
            uint32_t phase = 0, lo = lo0 + (kResX ? kRingLo : 0u);
            [[maybe_unused]] uint32_t x_gen = 0;
            int64_t it = 0;
            for (int64_t t = t_first; t < t_last; t += t_step, ++it) {
                const int as = (int)(it & 1);
                [[maybe_unused]] bool new_x = false, last_x = false;
                if constexpr (kResX) {
                    const int64_t cb = t % c.n_cb;
                    new_x = t == t_first || cb == 0;
                    last_x = t + 1 == t_last || cb + 1 == c.n_cb;
                }
                mbar_wait(acc_empty + as, (uint32_t)((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < c.kblocks; ++kb) {
                    if constexpr (kResX) {
                        if (new_x) mbar_wait(x_full + kb, x_gen & 1u);
                    }
                    mbar_wait(full + stage, phase);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        if constexpr (kResX)
                            umma_f16_pair_lohi(d_tmem, lo0 + kb * kXLo + k * kKLo, lo + k * kKLo, desc_hi, desc_hi, idesc, acc);
                        else if (kSharedMma) umma_f16_pair_lohi(d_tmem, lo + k * kKLo, lo + kYLo + k * kKLo, desc_hi, desc_hi, idesc, acc);
                        else umma_f16_lohi(d_tmem, lo + k * kKLo, lo + kYLo + k * kKLo, desc_hi, idesc, acc);
                    }
                    // stage reusable (in both CTAs of a cluster) once these MMAs have read it
                    if (kSharedMma) umma_commit_pair(empty + stage);
                    else if (kMcast) umma_commit_mcast(empty + stage, (uint16_t)3);
                    else umma_commit(empty + stage);
                    if constexpr (kResX) {
                        if (last_x) umma_commit_pair(x_empty + kb);  // the strip's last tile: slot kb may take the next strip
                    }
                    lo += kStageLo;
                    if (++stage == L::kStages) {
                        stage = 0;
                        phase ^= 1;
                        lo = lo0 + (kResX ? kRingLo : 0u);
                    }
                }
                if (kResX && last_x) ++x_gen;
                if (kSharedMma) umma_commit_pair(acc_full + as);  // accumulator complete -> both epilogues
                else umma_commit(acc_full + as);
            }
        }
    } else if (warp == kLoaderWarp) {
        // ============================ vector loader: per-column and per-row epilogue operands -> smem
        int64_t it = 0;
        for (int64_t t = t_first; t < t_last; t += t_step, ++it) {
            const int as = (int)(it & 1);
            int rb, cb;
            coords(t, rb, cb);
            const int64_t row0 = ((int64_t)rb * kCluster + crank) * BM, col0 = (int64_t)cb * BN;
            mbar_wait(vec_empty + as, (uint32_t)((it >> 1) & 1) ^ 1);  // the epilogue is done with this buffer
            float* cv = colvec + as * (kMaxColVecs * kColVecStride);
            float* rv = rowvec + as * (kMaxRowVecs * BM);
            // phase 1: every global load of the tile (fetch_* = loads only), phase 2: arithmetic + smem stores
            constexpr int kColIters = (BN + 31) / 32, kRowIters = BM / 32;
            uint32_t rc[kColIters][kMaxColRaw] = {}, rr[kRowIters][kMaxRowRaw] = {};
#pragma unroll
            for (int i = 0; i < kColIters; ++i) {
                const int col = lane + 32 * i;
                Policy::fetch_col(p, c, col0 + col, col < BN && col0 + col < c.cols, rc[i]);
            }
#pragma unroll
            for (int i = 0; i < kRowIters; ++i) {
                const int r = lane + 32 * i;
                Policy::fetch_row(p, c, row0 + r, row0 + r < c.rows, rr[i]);
            }
            // keep the two phases apart: neither the front end nor ptxas may sink a load to its use.  The memory
            // clobbers alone do not stop NVVM from hoisting the first ARITHMETIC on a loaded value above the barrier;
            // ptxas may then allocate one register for several loads, and their round trips serialise (seen in a
            // variant of the MIL-NCE gradient policy: seven loads in a row through one register, 1.42 instead of
            // 1.05 ms per block; the shipped policies had two or three).  So every raw value is re-defined by a
            // volatile move behind the barrier: all loads of a tile are in flight before the first use.
            asm volatile("" ::: "memory");
            __syncwarp();
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = 0; i < kColIters; ++i)
#pragma unroll
                for (int k = 0; k < kMaxColRaw; ++k) asm volatile("mov.b32 %0, %0;" : "+r"(rc[i][k]));
#pragma unroll
            for (int i = 0; i < kRowIters; ++i)
#pragma unroll
                for (int k = 0; k < kMaxRowRaw; ++k) asm volatile("mov.b32 %0, %0;" : "+r"(rr[i][k]));
#pragma unroll
            for (int i = 0; i < kColIters; ++i) {
                const int col = lane + 32 * i;
                float tc[kMaxColVecs];
                Policy::make_col(p, c, col < BN && col0 + col < c.cols, rc[i], tc);
                if (col < BN) {
#pragma unroll
                    for (int k = 0; k < Policy::kColVecs; ++k) cv[k * kColVecStride + col] = tc[k];
                }
            }
#pragma unroll
            for (int i = 0; i < kRowIters; ++i) {
                const int r = lane + 32 * i;
                float tr[kMaxRowVecs];
                Policy::make_row(p, c, row0 + r, row0 + r < c.rows, rr[i], tr);
#pragma unroll
                for (int k = 0; k < Policy::kRowVecs; ++k) rv[k * BM + r] = tr[k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(vec_full + as);
        }
    } else {
        // ======================================================================== epilogue
        const int quad = warp & 3;                   // TMEM lane quadrant this warp may read
        const int half = (warp - kEpiWarp0) >> 2;    // column group of this warp
        static_assert(BN % (32 * G) == 0, "column groups are whole 32-column chunks");
        constexpr int kChunks = BN / (32 * G);       // 32-column chunks per warp
        static_assert(!Policy::kStoresG || kChunks % 2 == 0, "gradient-matrix slabs are 64 columns wide");
        static_assert(!Policy::kByteG || kChunks % 4 == 0, "one-byte gradient-matrix slabs are 128 columns wide");
        // with >= 3 warps per scheduler the TMEM load latency is hidden by the other warps; with 2 the
        // next chunk's load is kept in flight in a second register buffer
#ifdef PB2_NO_PINGPONG  // measurement builds: one TMEM register buffer (32 fewer live registers)
        constexpr bool kPingPong = false;
#else
        constexpr bool kPingPong = (G <= 2) && (kChunks > 1);
#endif
        Policy pol;
        pol.kernel_begin(p);
        OutStage os;
        os.buf = out_stage + (warp - kEpiWarp0) * L::kOutBufs * kOutSlabBytes;
        os.tmap = &tm_out;
        os.slab = 0;
        os.mask = L::kOutBufs - 1;
#ifdef PB2_MEASURE
        os.skip = c.diag_only < 0 ? -c.diag_only : 0;
#endif
        int64_t it = 0;
        for (int64_t t = t_first; t < t_last; t += t_step, ++it) {
            const int as = (int)(it & 1);
            int rb, cb;
            coords(t, rb, cb);
            TileCtx ctx;
            ctx.row0 = ((int64_t)rb * kCluster + crank) * BM;
            ctx.col0 = (int64_t)cb * BN;
            ctx.row = ctx.row0 + quad * 32 + lane;
            ctx.row_valid = ctx.row < c.rows;
            ctx.cols_valid = (int)min((int64_t)BN, c.cols - ctx.col0);
            ctx.cb = cb;
            ctx.half = half;
            ctx.warp_cols = kChunks * 32;
            ctx.quad = quad;
            const float* cv = colvec + as * (kMaxColVecs * kColVecStride);
            mbar_wait(vec_full + as, (uint32_t)((it >> 1) & 1));  // operands staged by the loader warp
            pol.tile_begin(p, c, ctx, rowvec + as * (kMaxRowVecs * BM) + quad * 32 + lane);
            mbar_wait(acc_full + as, (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN);
            const int c0 = half * kChunks * 32;  // first column of this warp's half
            // two register buffers: the TMEM load of chunk ch+1 is in flight while chunk ch is processed
            uint32_t va[32];
            __syncwarp();
            tmem_ld32(t_lane + c0, va);
            if constexpr (!kPingPong) {
#pragma unroll 1
                for (int ch = 0; ch < kChunks; ++ch) {
                    tmem_ld_wait();
                    pol.chunk(p, c, ctx, ch, c0 + ch * 32, va, cv + c0 + ch * 32, os);
                    __syncwarp();
                    if (ch + 1 < kChunks) tmem_ld32(t_lane + c0 + (ch + 1) * 32, va);
                }
            } else {
                uint32_t vb[32];
#pragma unroll 1
                for (int ch = 0; ch < kChunks; ch += 2) {
                    tmem_ld_wait();
                    __syncwarp();
                    tmem_ld32(t_lane + c0 + (ch + 1) * 32, vb);
                    pol.chunk(p, c, ctx, ch, c0 + ch * 32, va, cv + c0 + ch * 32, os);
                    tmem_ld_wait();
                    __syncwarp();
                    if (ch + 2 < kChunks) tmem_ld32(t_lane + c0 + (ch + 2) * 32, va);
                    pol.chunk(p, c, ctx, ch + 1, c0 + (ch + 1) * 32, vb, cv + c0 + (ch + 1) * 32, os);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(vec_empty + as);  // the loader may restage this buffer
                if (kSharedMma) mbar_arrive_cluster(mapa_u32(smem_u32(acc_empty + as), 0));  // the leader's MMA warp
                else mbar_arrive(acc_empty + as);
            }
            pol.tile_end(p, c, ctx);
        }
        if (Policy::kStoresG || Policy::kStoresF32) os.finish(lane);
        pol.kernel_end(p, red, kEpiWarps);
    }
    tc_fence_before();
    if (kCluster == 2) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer still signals it
    else __syncthreads();
    if (warp == kTmaWarp) {
        tc_fence_after();
        if (kSharedMma) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------ host
PB2_KNOB g_skip_store = 0;  // measurement build: pb2_debug_force_bn(bn | 0x10000)

static int pick_bn(int64_t rows, int64_t cols, bool stores_g) {
    // widest tile that still yields at least ~one tile per SM; small problems are latency bound.
    // (128 x 192 tiles with 12 epilogue warps exist for the gradient-matrix producers -- force_bn 192 --
    // but measure ~4 % slower than 128 x 256 with 8: the fused pass runs at the board's power cap, where
    // time tracks energy, not issue-slot occupancy)
    const int64_t sms = sm_count();
    const int64_t rb = (rows + BM - 1) / BM;
    if (stores_g) return rb * ((cols + 255) / 256) >= sms ? 256 : 128;
    if (rb * ((cols + 255) / 256) >= sms) return 256;
    if (rb * ((cols + 127) / 128) >= sms) return 128;
    return 64;
}

struct OutMatrix {  // optional fp16 gradient matrix drained by TMA stores
    void* ptr = nullptr;
    int64_t ld = 0;
};

template <class Policy, int BN, int G, int kCtas>
static int launch_sim(const void* x, const void* y, int dtype, int64_t rows, int64_t cols, int dim, int64_t ldx, int64_t ldy,
                      const float* rinv_x, const float* rinv_y, float scale, const typename Policy::Params& pp,
                      const OutMatrix& om, cudaStream_t st, const char* what) {
    CUtensorMap tx, ty, to;
    int rc = make_tmap_2d(&tx, x, 2, (uint64_t)rows, (uint64_t)dim, (uint64_t)ldx * 2, BM, BK);
    if (rc) return rc;
    constexpr int kCluster = kCtas >= 2 ? 2 : 1;
    rc = make_tmap_2d(&ty, y, 2, (uint64_t)cols, (uint64_t)dim, (uint64_t)ldy * 2, BN / kCluster, BK);
    if (rc) return rc;
    if (Policy::kStoresG && om.ptr) {
        if (Policy::kByteG) rc = make_tmap_2d(&to, om.ptr, 1, (uint64_t)rows, (uint64_t)cols, (uint64_t)om.ld, 32, 128);
        else rc = make_tmap_2d(&to, om.ptr, 2, (uint64_t)rows, (uint64_t)cols, (uint64_t)om.ld * 2, 32, 64);
        if (rc) return rc;
    } else if (Policy::kStoresF32 && om.ptr) {
        rc = make_tmap_2d(&to, om.ptr, 4, (uint64_t)rows, (uint64_t)cols, (uint64_t)om.ld * 4, 32, 32);
        if (rc) return rc;
    } else {
        to = tx;  // never dereferenced
    }
    SimCommon c;
    c.rows = rows;
    c.cols = cols;
    c.kblocks = dim / BK;
    c.n_rb = (int)((rows + BM * kCluster - 1) / (BM * kCluster));  // row blocks of a CTA (cluster) tile
    c.n_cb = (int)((cols + BN - 1) / BN);
    c.n_tiles = (int64_t)c.n_rb * c.n_cb;
    c.diag_only = -g_skip_store;
    if (std::is_same<Policy, DiagPolicy>::value) {  // paired rows: only the diagonal tiles
        c.diag_only = 1;
        c.n_cb = 0;
        c.n_tiles = c.n_rb;
    }
    c.rinv_x = rinv_x;
    c.rinv_y = rinv_y;
    c.scale = scale;
    c.fmt = dtype == PB2_F16 ? (uint32_t)kFmtF16 : (uint32_t)kFmtBF16;
    c.early_operands = g_operands_ready_depth > 0 ? 1 : 0;
    auto kern = sim_kernel<Policy, BN, G, kCtas>;
    constexpr int smem = SimSmem<BN, G, uses_stage<Policy>::value, kCtas>::kTotal;
    static PerDeviceOnce configured;  // per instantiation and device
    rc = ensure_dynamic_smem(configured, kern, smem, what);
    if (rc) return rc;
    const int grid = (int)std::min<int64_t>(c.n_tiles, pb2_sim_grid() / kCluster) * kCluster;
    rc = check_cuda(launch_ex(kern, (unsigned)grid, (unsigned)sim_threads(G), (size_t)smem, st, kCluster, tx, ty, to, c, pp), what);
    if (rc) return rc;
    return check_launch(what);
}

PB2_KNOB g_force_bn = 0;  // measurement build: pb2_debug_force_bn
// pb2_debug_sim_pair: 1 = CTA pairs whenever the tile is 256 wide, 2 = multicast clusters, 0 = independent CTAs, -1 = default = 1 when
// the problem has at least a tile per SM.
// CTA pairs (one M = 256 tcgen05.mma per two SMs; each CTA stages its X rows and half of the Y tile) move a third less
// operand data per SM (L2 -> smem -> tensor core).  Through round 2 they measured 5 % SLOWER than independent CTAs and
// stayed an option -- until the cause turned out to be the hand-back of the accumulator, not the coupling of the two
// epilogues: `mbarrier.arrive.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR, which every epilogue
// warp paid once per tile while its count atomics and TMA stores were still in flight.  With the default-semantics
// arrive (common.cuh: mbar_arrive_cluster; the TMEM hand-back is ordered by tcgen05.wait::ld + tcgen05.fence) pairs
// win for every policy, back to back on a 32768^2 block on one board (tools/ab_hinge.py, profiles/r2_ab_sim_pairs.txt):
// rank 0.855 -> 0.795 ms (1383 TF/s), one-pass log-sum-exp 1.14 -> 1.06 ms, hinge + rank + one-byte G 1.155 -> 1.13 ms;
// inside the steps: 65536-clip MIL-NCE 16.8-18.4 -> 15.5-16.1 ms, 2^20 hinge gallery 2687-2697 -> 2676-2681 ms (that
// step sits at the board's power cap, where time tracks energy).  2 = clusters of 2 with independent MMAs and a
// multicast Y tile: within noise of independent CTAs; measurement builds only.
PB2_KNOB g_sim_pair = -1;

template <class Policy>
static int dispatch_sim(const void* x, const void* y, int dtype, int64_t rows, int64_t cols, int dim, int64_t ldx, int64_t ldy,
                        const float* rinv_x, const float* rinv_y, float scale, const typename Policy::Params& pp,
                        void* stream, const char* what, int force_bn = 0, const OutMatrix& om = OutMatrix()) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (dim <= 0 || dim % BK != 0) return set_error(PB2_ERR_ARG, "%s: dim must be a positive multiple of 64", what);
    if (!x || !y) return set_error(PB2_ERR_ARG, "%s: null operand", what);
    if (dtype != PB2_BF16 && dtype != PB2_F16)
        return set_error(PB2_ERR_ARG, "%s: tensor-core operands are bf16 or fp16 (fp32 rows go through pb2_split_f16)", what);
    if (rows > 0x7fffffffll * BM / 2 || cols > 0x7fffffffll) return set_error(PB2_ERR_ARG, "%s: too large", what);
    cudaStream_t st = (cudaStream_t)stream;
    int bn = force_bn ? force_bn : pick_bn(rows, cols, Policy::kStoresG);
    if (Policy::kStoresG && bn == 64) bn = 128;
    if (!Policy::kStoresG && bn == 192) bn = 256;
#define PB2_SIM(B, GG, CC) \
    launch_sim<Policy, B, GG, CC>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, om, st, what)
    // pairs by default once every SM has a tile of its own (below that the passes are latency bound)
    const bool many = ((rows + BM - 1) / BM) * ((cols + 255) / 256) >= sm_count();
    const int cluster_mode = g_sim_pair >= 0 ? g_sim_pair : (many ? 1 : 0);
    const bool pair = bn == 256 && cluster_mode == 1;
#ifdef PB2_MEASURE
    const bool mcast = bn == 256 && cluster_mode == 2;
#endif
#ifdef PB2_MEASURE
    // 3 = CTA pairs with a resident X strip (dim <= 512; built for the one-byte-G hinge pass and the rank pass): the pair
    // takes a contiguous run of column tiles of ONE row block and keeps that block's X strip in shared memory, so only Y
    // streams (half of the pass's L2 -> SM operand traffic).  Measured, modes alternating on one board: hinge + rank +
    // one-byte G on a 32768^2 block alone 1.15 -> 1.11 ms sustained (-3 %), the rank pass 0.80 -> 0.82 ms (+3 %: its ring
    // is five 16 KiB stages of Y instead of four of X + Y) -- and INSIDE the 262144-clip gallery step 164.9-165.2 ->
    // 163.8-164.4 ms on one box, 162.8-162.9 -> 162.8-163.1 ms on the next: nothing, the step sits at the power cap
    // (profiles/r2b_ab_gallery_resx.txt).  By the rule of section 4.1 (a change is kept when the step that contains it
    // moves) plain pairs stay the product; the variant stays bit-identical and selectable here.
    const bool resx = bn == 256 && dim / BK <= 8 && g_sim_pair == 3;
#endif
    if constexpr (std::is_same<Policy, DiagPolicy>::value) {
        return PB2_SIM(128, 2, 1);
    } else if constexpr (Policy::kByteG) {
#ifdef PB2_MEASURE  // tools/ab_gallery_pair.py
        if (mcast) return PB2_SIM(256, 2, 3);
#endif
#ifdef PB2_MEASURE
        if (resx) return PB2_SIM(256, 2, 4);
#endif
        if (pair) return PB2_SIM(256, 2, 2);
        return PB2_SIM(256, 2, 1);  // a warp's 128 columns are one slab of bytes: 256-wide tiles only
    } else if constexpr (Policy::kStoresG) {
        if (bn == 192) return PB2_SIM(192, 3, 1);
#ifdef PB2_MEASURE
        if (mcast) return PB2_SIM(256, 2, 3);
#endif
        if (pair) return PB2_SIM(256, 2, 2);
#ifdef PB2_HINGE_G4  // measurement builds: four column groups (16 epilogue warps, 64 columns each) on 256-wide tiles
        if (bn == 256) return PB2_SIM(256, 4, 1);
#else
        if (bn == 256) return PB2_SIM(256, 2, 1);
#endif
        return PB2_SIM(128, 2, 1);
    } else {
#ifdef PB2_MEASURE
        if (mcast) return PB2_SIM(256, 2, 3);
        if constexpr (std::is_same<Policy, RankPolicy>::value) {
            if (resx) return PB2_SIM(256, 2, 4);
        }
        if (cluster_mode == 3) return PB2_SIM(256, 2, 2);  // policies without a resident-X build: plain pairs
#endif
        if (pair) return PB2_SIM(256, 2, 2);
        if (bn == 256) return PB2_SIM(256, 2, 1);
        if (bn == 128) return PB2_SIM(128, 2, 1);
        return PB2_SIM(64, 2, 1);
    }
#undef PB2_SIM
}

}  // namespace pb2

using namespace pb2;

extern "C" int pb2_sim_grid(void) { return sm_count(); }
#ifdef PB2_MEASURE
extern "C" int pb2_debug_sim_pair(int mode) {
    g_sim_pair = mode;
    return PB2_OK;
}
extern "C" int pb2_debug_force_bn(int bn) {
    g_skip_store = (bn >> 16) & 7;
    bn &= 0xffff;
    g_force_bn = (bn == 64 || bn == 128 || bn == 192 || bn == 256) ? bn : 0;
    return PB2_OK;
}
#endif

extern "C" int pb2_sim_matrix(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                              int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float* out,
                              int64_t ld_out, void* stream) {
    if (!out && rows > 0 && cols > 0) return set_error(PB2_ERR_ARG, "sim_matrix: null output");
    if (rows > 0 && cols > 0 && ((reinterpret_cast<uintptr_t>(out) & 15) || ld_out % 4 != 0 || ld_out < cols))
        return set_error(PB2_ERR_ARG, "sim_matrix: out must be 16-byte aligned with ld_out %% 4 == 0, ld_out >= cols");
    StorePolicy::Params pp{out, ld_out};
    OutMatrix om;
    om.ptr = out;
    om.ld = ld_out;
    return dispatch_sim<StorePolicy>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                     "sim_matrix", g_force_bn, om);
}

extern "C" int pb2_sim_diag(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t n, int dim, int dtype,
                            int64_t ldx, int64_t ldy, float* out, float* dist_out, float* thr_out, void* stream) {
    DiagPolicy::Params pp{out, dist_out, thr_out};
    return dispatch_sim<DiagPolicy>(x, y, dtype, n, n, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream, "sim_diag", 128);
}

extern "C" int pb2_sim_rank(const void* q, const void* g, const float* rinv_q, const float* rinv_g,
                            const float* pos_thr, const int64_t* pos_col, int64_t rows, int64_t cols,
                            int64_t col_offset, int dim, int dtype, int64_t ldq, int64_t ldg, int32_t* rank, void* stream) {
    if (rows > 0 && cols > 0 && (!pos_thr || !pos_col || !rank)) return set_error(PB2_ERR_ARG, "sim_rank: null");
    RankPolicy::Params pp{pos_thr, pos_col, col_offset, rank};
    return dispatch_sim<RankPolicy>(q, g, dtype, rows, cols, dim, ldq, ldg, rinv_q, rinv_g, 1.0f, pp, stream, "sim_rank",
                                    g_force_bn);
}

static int check_gmat(const void* gmat, int64_t ld_g, int64_t cols, const char* what, int elem_bytes = 2) {
    if (gmat && ((ld_g * elem_bytes) % 16 != 0 || ld_g < cols || (reinterpret_cast<uintptr_t>(gmat) & 15)))
        return set_error(PB2_ERR_ARG, "%s: gmat needs 16-byte alignment, a 16-byte multiple row pitch and ld_g >= cols", what);
    return PB2_OK;
}

extern "C" int pb2_sim_hinge(const void* x, const void* y, const float* rinv_x, const float* rinv_y,
                             const float* diag_row, const float* diag_col, int64_t rows, int64_t cols,
                             int64_t row_offset, int64_t col_offset, int dim, int dtype, int64_t ldx, int64_t ldy, float margin,
                             float* loss_partial, int n_partials, int32_t* row_cnt, int32_t* col_cnt, void* gmat,
                             int g_dtype, int64_t ld_g, const float* pos_thr, int32_t* rank, void* stream) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (!diag_row || !diag_col || !loss_partial || !row_cnt || !col_cnt)
        return set_error(PB2_ERR_ARG, "sim_hinge: null");
    const bool prezeroed = n_partials < 0;  // pb2_hinge_prep already cleared the partial buffer
    if (prezeroed) n_partials = -n_partials;
    if (n_partials < pb2_sim_grid()) return set_error(PB2_ERR_ARG, "sim_hinge: loss_partial too small");
    if (gmat && g_dtype != PB2_F16 && g_dtype != PB2_U8)
        return set_error(PB2_ERR_ARG, "sim_hinge: the gradient matrix is PB2_F16 or PB2_U8");
    const bool byte_g = gmat && g_dtype == PB2_U8;
    int rc = check_gmat(gmat, ld_g, cols, "sim_hinge", byte_g ? 1 : 2);
    if (rc) return rc;
    if ((pos_thr == nullptr) != (rank == nullptr))
        return set_error(PB2_ERR_ARG, "sim_hinge: pos_thr and rank go together");
    if (!prezeroed) {
        rc = check_cuda(cudaMemsetAsync(loss_partial, 0, sizeof(float) * n_partials, (cudaStream_t)stream),
                        "sim_hinge memset");
        if (rc) return rc;
    }
    HingeParams pp{diag_row, diag_col, row_offset, col_offset, margin,   loss_partial,
                   row_cnt,  col_cnt,  gmat ? 1 : 0, pos_thr, rank};
    OutMatrix om;
    om.ptr = gmat;
    om.ld = ld_g;
    if (byte_g) {
        if (rank)
            return dispatch_sim<HingePolicyT<true, true>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp,
                                                          stream, "sim_hinge+rank (u8 G)", 256, om);
        return dispatch_sim<HingePolicyT<false, true>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream,
                                                       "sim_hinge (u8 G)", 256, om);
    }
    if (rank)
        return dispatch_sim<HingePolicyT<true>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream,
                                                "sim_hinge+rank", g_force_bn, om);
    return dispatch_sim<HingePolicyT<false>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, 1.0f, pp, stream,
                                             "sim_hinge", g_force_bn, om);
}

extern "C" int pb2_sim_lse_col_parts(int64_t rows) { return (int)((rows + BM - 1) / BM) * 4; }

extern "C" int pb2_sim_lse_both(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                                int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float bound,
                                float* row_part_sum, float* col_part_sum, void* stream) {
    if (rows > 0 && cols > 0 && (!row_part_sum || !col_part_sum)) return set_error(PB2_ERR_ARG, "sim_lse_both: null");
    const float shift = bound * 1.4426950408889634f;
    if (!(bound >= 0.f) || !(shift <= 60.f))
        return set_error(PB2_ERR_ARG, "sim_lse_both: needs 0 <= bound and bound * log2(e) <= 60 (use the two-pass path)");
    LseBothParams pp{row_part_sum, col_part_sum, shift, nullptr, nullptr, nullptr, 0, 0, nullptr};
    const int bn = g_force_bn == 256 || g_force_bn == 128 ? g_force_bn : (pick_bn(rows, cols, false) == 256 ? 256 : 128);
    return dispatch_sim<LseBothPolicyT<false>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                               "sim_lse_both", bn);
}

// The same pass with the ranking of pb2_sim_rank fused in (north-star kernels (a) + (b) from one S pass): `rank[i]` +=
// the number of columns j != row_offset + i - col_offset whose cosine fl32(fl32(<x_i, y_j> * rank_rinv_x[i]) *
// rank_rinv_y[j]) reaches pos_thr[i] (pb2_sim_diag's threshold).  Counts are bit-identical to pb2_sim_rank's.
extern "C" int pb2_sim_lse_both_rank(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                                     int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float bound,
                                     float* row_part_sum, float* col_part_sum, const float* rank_rinv_x,
                                     const float* rank_rinv_y, const float* pos_thr, int64_t row_offset, int64_t col_offset,
                                     int32_t* rank, void* stream) {
    if (rows > 0 && cols > 0 && (!row_part_sum || !col_part_sum || !pos_thr || !rank))
        return set_error(PB2_ERR_ARG, "sim_lse_both_rank: null");
    const float shift = bound * 1.4426950408889634f;
    if (!(bound >= 0.f) || !(shift <= 60.f))
        return set_error(PB2_ERR_ARG, "sim_lse_both_rank: needs 0 <= bound and bound * log2(e) <= 60 (use the two-pass path)");
    LseBothParams pp{row_part_sum, col_part_sum, shift, rank_rinv_x, rank_rinv_y, pos_thr, row_offset, col_offset, rank};
    const int bn = g_force_bn == 256 || g_force_bn == 128 ? g_force_bn : (pick_bn(rows, cols, false) == 256 ? 256 : 128);
    return dispatch_sim<LseBothPolicyT<true>>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                              "sim_lse_both_rank", bn);
}

// LSE partial layout is fixed to the 128-column tile so the caller can size buffers up front.
extern "C" int pb2_sim_lse_parts(int64_t cols) { return (int)((cols + 127) / 128) * 2; }

extern "C" int pb2_sim_lse_rows(const void* x, const void* y, const float* rinv_x, const float* rinv_y, int64_t rows,
                                int64_t cols, int dim, int dtype, int64_t ldx, int64_t ldy, float scale, float* part_max,
                                float* part_sum, void* stream) {
    if (rows > 0 && cols > 0 && (!part_max || !part_sum)) return set_error(PB2_ERR_ARG, "sim_lse_rows: null");
    LseRowPolicy::Params pp{part_max, part_sum};
    // 256-wide tiles once they fill the machine (the partial layout is the 128-column one either way)
    const int bn = g_force_bn == 256 || g_force_bn == 128 ? g_force_bn : (pick_bn(rows, cols, false) == 256 ? 256 : 128);
    return dispatch_sim<LseRowPolicy>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                      "sim_lse_rows", bn);
}

extern "C" int pb2_sim_lse_grad(const void* x, const void* y, const float* rinv_x, const float* rinv_y,
                                const float* den_row, const float* den_col, int64_t rows, int64_t cols, int dim, int dtype,
                                int64_t ldx, int64_t ldy, float scale, void* gmat, int64_t ld_g, void* stream) {
    if (rows <= 0 || cols <= 0) return PB2_OK;
    if (!den_row || !den_col || !gmat) return set_error(PB2_ERR_ARG, "sim_lse_grad: null");
    int rc = check_gmat(gmat, ld_g, cols, "sim_lse_grad");
    if (rc) return rc;
    LseGradPolicy::Params pp{den_row, den_col};
    OutMatrix om;
    om.ptr = gmat;
    om.ld = ld_g;
    return dispatch_sim<LseGradPolicy>(x, y, dtype, rows, cols, dim, ldx, ldy, rinv_x, rinv_y, scale, pp, stream,
                                       "sim_lse_grad", g_force_bn, om);
}
