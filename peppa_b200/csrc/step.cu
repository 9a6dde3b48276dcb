// The TripletLoss training step (pig/models.py:262 -> pig/loss.py:33-48 + autograd) as C calls that chain their kernel
// launches with programmatic dependent launch on the caller's stream.  At batch ~1k the step is launch bound: one call
// keeps the host side to a single FFI crossing and the scratch in caller-provided buffers.
//
//   pb2_hinge_step      forward AND gradients in one call, four launches:
//                         hinge_prep -> sim_hinge (tcgen05, fused loss / counts / gradient matrix) -> grad_gemm_dual (both
//                         backward products in one grid) -> hinge_finish2 (Jacobians + scalar loss)
//   pb2_hinge_forward   the forward half, three launches: hinge_prep -> sim_hinge -> grad_gemm_dual with the scalar loss
//                         folded by a spare CTA of the product grid; what the backward needs (both products, 1/||row||,
//                         the indicator counts) stays in a caller-owned STATE buffer; without one (state == NULL) it is
//                         the loss alone: hinge_prep -> sim_hinge without a gradient matrix -> the fold
//   pb2_hinge_backward  the backward half, one launch: hinge_finish2 over the state with autograd's grad_output applied
//                         in fp32 before the rounding to the gradients' dtype
// Forward + backward are four launches like pb2_hinge_step, against its five with the pb2_scale_pair a deferred
// grad_output needs, and no fp32 gradient round trip; both routes give the same bits.
#include <cuda_fp16.h>

#include "host_util.h"
#include "peppa_b200.h"

namespace {
constexpr int64_t kAlign = 256;
int64_t up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
// what outlives the forward: read by hinge_finish2
struct StateLayout {
    int64_t rinv_v, rinv_a, row_cnt, col_cnt, pv, pa, total;
};
// scratch of the forward
struct Layout {
    int64_t diag, part, vh, ah, g, vx, ax, sv, sa, total, ld_g;
    int n_part;
};
StateLayout state_layout(int64_t n, int dim) {
    StateLayout S;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        const int64_t at = o;
        o += up(bytes);
        return at;
    };
    S.rinv_v = take(n * 4);
    S.rinv_a = take(n * 4);
    S.row_cnt = take(n * 4);
    S.col_cnt = take(n * 4);
    S.pv = take(n * dim * 4);
    S.pa = take(n * dim * 4);
    S.total = o;
    return S;
}
Layout layout(int64_t n, int dim, int in_dtype) {
    Layout L;
    L.n_part = pb2_sim_grid();
    L.ld_g = (n + 63) / 64 * 64;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        const int64_t at = o;
        o += up(bytes);
        return at;
    };
    L.diag = take(n * 4);
    L.part = take((int64_t)L.n_part * 4);
    L.vh = take(n * dim * 2);
    L.ah = take(n * dim * 2);
    L.g = take(n * L.ld_g * 2);
    // fp32 inputs: split-fp16 tensor-core operands [n, 3 dim] and their per-row power-of-two scales (pb2_split_f16)
    L.vx = in_dtype == PB2_F32 ? take(n * 3 * dim * 2) : -1;
    L.ax = in_dtype == PB2_F32 ? take(n * 3 * dim * 2) : -1;
    L.sv = in_dtype == PB2_F32 ? take(n * 4) : -1;
    L.sa = in_dtype == PB2_F32 ? take(n * 4) : -1;
    L.total = o;
    return L;
}
PB2_KNOB g_step_stages = 4;  // measurement build (pb2_debug_step_stages): stop the step after its first k kernels

struct State {
    float *rinv_v, *rinv_a, *pv, *pa;
    int32_t *row_cnt, *col_cnt;
};
State state_at(void* state, int64_t n, int dim) {
    const StateLayout S = state_layout(n, dim);
    char* s = static_cast<char*>(state);
    State st;
    st.rinv_v = reinterpret_cast<float*>(s + S.rinv_v);
    st.rinv_a = reinterpret_cast<float*>(s + S.rinv_a);
    st.row_cnt = reinterpret_cast<int32_t*>(s + S.row_cnt);
    st.col_cnt = reinterpret_cast<int32_t*>(s + S.col_cnt);
    st.pv = reinterpret_cast<float*>(s + S.pv);
    st.pa = reinterpret_cast<float*>(s + S.pa);
    return st;
}

int check_common(const char* what, const void* v, const void* a, int in_dtype, int64_t n, int dim) {
    using pb2::set_error;
    if (n <= 0) return set_error(PB2_ERR_ARG, "%s: empty batch", what);
    if (!v || !a) return set_error(PB2_ERR_ARG, "%s: null", what);
    if (dim <= 0 || dim % 64 != 0) return set_error(PB2_ERR_ARG, "%s: dim must be a positive multiple of 64", what);
    if (in_dtype != PB2_BF16 && in_dtype != PB2_F16 && in_dtype != PB2_F32)
        return set_error(PB2_ERR_ARG, "%s: inputs are bf16, fp16 or fp32 rows", what);
    return PB2_OK;
}

// hinge_prep -> sim_hinge -> both gradient products.  fold.loss_out != nullptr: the scalar loss is folded beside the
// products (a spare CTA of their grid, or a launch of its own when they do not fit one grid); the caller's PdlScope is
// alive.  loss_only: no gradient matrix and no products, the fold is the last launch.  *complete: false when a
// measurement build truncated the step (pb2_debug_step_stages).
int forward_launches(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float margin,
                     char* w, const Layout& L, const State& st, const float* rinv_v_in, const float* rinv_a_in,
                     pb2::HingeFold fold, bool fold_in_forward, bool loss_only, bool* complete, void* stream) {
    *complete = false;
    float* diag = reinterpret_cast<float*>(w + L.diag);
    float* part = reinterpret_cast<float*>(w + L.part);
    void* vh = w + L.vh;
    void* ah = w + L.ah;
    void* g = w + L.g;
    // tensor-core operands: bf16 / fp16 rows as they are; fp32 rows as the split-fp16 pair of their normalised,
    // power-of-two-scaled values (contraction length 3 dim), with the scales in place of 1/||row|| in the epilogue
    const bool split = in_dtype == PB2_F32;
    void* vx = split ? w + L.vx : nullptr;
    void* ax = split ? w + L.ax : nullptr;
    float* sv = split ? reinterpret_cast<float*>(w + L.sv) : nullptr;
    float* sa = split ? reinterpret_cast<float*>(w + L.sa) : nullptr;
    int rc = pb2_hinge_prep(v, a, in_dtype, n, dim, ldv, lda, st.rinv_v, st.rinv_a, diag, vh, ah, st.row_cnt, st.col_cnt, part,
                            L.n_part, vx, ax, sv, sa, rinv_v_in, rinv_a_in, stream);
    if (rc || g_step_stages < 2) return rc;
    // bf16 / fp16 rows are the caller's own tensors: complete before hinge_prep (which waits BEFORE it triggers its
    // dependents) got past its wait, so the similarity pass streams them through the tensor cores while hinge_prep
    // still computes the norms and the diagonal its epilogue needs.  The split operands of fp32 rows are hinge_prep's
    // own output: no early start there.
    {
        pb2::OperandsReadyScope early(!split);
        rc = pb2_sim_hinge(split ? vx : v, split ? ax : a, split ? sv : st.rinv_v, split ? sa : st.rinv_a, diag, diag, n, n, 0, 0,
                           split ? 3 * dim : dim, split ? PB2_F16 : in_dtype, split ? 3 * (int64_t)dim : ldv,
                           split ? 3 * (int64_t)dim : lda, margin, part, -L.n_part, st.row_cnt, st.col_cnt,
                           loss_only ? nullptr : g, PB2_F16, L.ld_g, nullptr, nullptr, stream);
    }
    if (rc || g_step_stages < 3) return rc;
    fold.loss_partial = part;
    fold.n_partials = L.n_part;
    fold.diag = diag;
    if (!fold_in_forward) fold.loss_out = nullptr;
    if (loss_only) {  // no gradient matrix, no products: the fold is the third and last launch
        *complete = true;
        return pb2::hinge_fold(fold, stream);
    }
    // dV partials = G A^, dA partials = G^T V^: one launch when all their tiles fit the machine at once
    bool folded = false;
    rc = pb2::grad_gemm_dual_fold(g, PB2_F16, n, n, L.ld_g, ah, vh, PB2_F16, dim, dim, dim, 1.0f, st.pv, st.pa, dim, dim, fold,
                                  &folded, stream);
    if (rc) return rc;
    if (fold.loss_out && !folded) {
        rc = pb2::hinge_fold(fold, stream);
        if (rc) return rc;
    }
    *complete = g_step_stages >= 4;
    return PB2_OK;
}

pb2::HingeFold fold_of(const State& st, int64_t n, float margin, float* loss_out) {
    pb2::HingeFold f;
    f.row_cnt = st.row_cnt;
    f.col_cnt = st.col_cnt;
    f.rinv_v = st.rinv_v;
    f.rinv_a = st.rinv_a;
    f.n = n;
    f.margin = margin;
    f.coef = 1.0f / ((float)n * (float)n);
    f.loss_out = loss_out;
    return f;
}
}  // namespace

#ifdef PB2_MEASURE
// Timing only (tools/timeline_train1024.py: the cost of each kernel INSIDE the chained step): 1 = hinge_prep, 2 = + the
// similarity pass, 3 = + the gradient products, 4 = the whole step.  Truncated steps leave loss / gradients unwritten.
extern "C" int pb2_debug_step_stages(int k) {
    g_step_stages = k >= 1 && k <= 4 ? k : 4;
    return PB2_OK;
}
#endif

extern "C" int64_t pb2_hinge_step_workspace(int64_t n, int dim, int in_dtype) {
    return n > 0 && dim > 0 ? layout(n, dim, in_dtype).total + state_layout(n, dim).total : 0;
}

extern "C" int64_t pb2_hinge_forward_workspace(int64_t n, int dim, int in_dtype) {
    return n > 0 && dim > 0 ? layout(n, dim, in_dtype).total : 0;
}

extern "C" int64_t pb2_hinge_state_bytes(int64_t n, int dim) { return n > 0 && dim > 0 ? state_layout(n, dim).total : 0; }

extern "C" int pb2_hinge_step(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float margin,
                              void* workspace, int64_t workspace_bytes, float* loss_out, void* d_v, void* d_a,
                              int out_dtype, const float* rinv_v_in, const float* rinv_a_in, void* stream) {
    using pb2::set_error;
    int rc = check_common("hinge_step", v, a, in_dtype, n, dim);
    if (rc) return rc;
    if (!workspace || !loss_out || !d_v || !d_a) return set_error(PB2_ERR_ARG, "hinge_step: null");
    if ((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) != 0)
        return set_error(PB2_ERR_ARG, "hinge_step: workspace must be 256-byte aligned");
    const Layout L = layout(n, dim, in_dtype);
    if (workspace_bytes < L.total + state_layout(n, dim).total) return set_error(PB2_ERR_ARG, "hinge_step: workspace too small");
    char* w = static_cast<char*>(workspace);
    const State st = state_at(w + L.total, n, dim);  // the state of this one-call form lives behind the scratch
    // programmatic dependent launch between the four kernels: each one's prologue (barrier init, TMEM
    // allocation, descriptor prefetch) overlaps its predecessor's tail
    pb2::PdlScope pdl;
    pb2::HingeFold fold = fold_of(st, n, margin, loss_out);
    bool complete = false;
    rc = forward_launches(v, a, in_dtype, n, dim, ldv, lda, margin, w, L, st, rinv_v_in, rinv_a_in, fold, false, false, &complete, stream);
    if (rc || !complete) return rc;
    fold.loss_partial = reinterpret_cast<float*>(w + L.part);
    fold.n_partials = L.n_part;
    fold.diag = reinterpret_cast<float*>(w + L.diag);
    return pb2::hinge_finish2_ex(st.pv, st.pa, v, a, in_dtype, n, dim, ldv, lda, st.rinv_v, st.rinv_a, st.row_cnt, st.col_cnt,
                                 fold.coef, nullptr, fold, d_v, d_a, out_dtype, stream);
}

extern "C" int pb2_hinge_forward(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda,
                                 float margin, void* workspace, int64_t workspace_bytes, void* state, int64_t state_bytes,
                                 float* loss_out, const float* rinv_v_in, const float* rinv_a_in, void* stream) {
    using pb2::set_error;
    int rc = check_common("hinge_forward", v, a, in_dtype, n, dim);
    if (rc) return rc;
    if (!workspace || !loss_out) return set_error(PB2_ERR_ARG, "hinge_forward: null");
    if (((reinterpret_cast<uintptr_t>(workspace) | reinterpret_cast<uintptr_t>(state)) & (kAlign - 1)) != 0)
        return set_error(PB2_ERR_ARG, "hinge_forward: workspace and state must be 256-byte aligned");
    const Layout L = layout(n, dim, in_dtype);
    const StateLayout S = state_layout(n, dim);
    // state == NULL: the loss alone (no gradient matrix, no products); 1/||row|| and the counts then live behind the
    // scratch, in a workspace of pb2_hinge_step_workspace bytes
    const bool loss_only = state == nullptr;
    if (workspace_bytes < L.total + (loss_only ? S.total : 0)) return set_error(PB2_ERR_ARG, "hinge_forward: workspace too small");
    if (!loss_only && state_bytes < S.total) return set_error(PB2_ERR_ARG, "hinge_forward: state too small");
    char* w = static_cast<char*>(workspace);
    const State st = state_at(loss_only ? static_cast<void*>(w + L.total) : state, n, dim);
    pb2::PdlScope pdl;
    bool complete = false;
    return forward_launches(v, a, in_dtype, n, dim, ldv, lda, margin, w, L, st, rinv_v_in, rinv_a_in, fold_of(st, n, margin, loss_out),
                            true, loss_only, &complete, stream);
}

extern "C" int pb2_hinge_backward(const void* state, int64_t state_bytes, const void* v, const void* a, int in_dtype, int64_t n,
                                  int dim, int64_t ldv, int64_t lda, const float* grad_out, void* d_v, void* d_a, int out_dtype,
                                  void* stream) {
    using pb2::set_error;
    int rc = check_common("hinge_backward", v, a, in_dtype, n, dim);
    if (rc) return rc;
    if (!state || !d_v || !d_a) return set_error(PB2_ERR_ARG, "hinge_backward: null");
    if ((reinterpret_cast<uintptr_t>(state) & (kAlign - 1)) != 0)
        return set_error(PB2_ERR_ARG, "hinge_backward: state must be 256-byte aligned");
    if (state_bytes < state_layout(n, dim).total) return set_error(PB2_ERR_ARG, "hinge_backward: state too small");
    const State st = state_at(const_cast<void*>(state), n, dim);
    // launched with programmatic stream serialization: scheduled while its predecessor (autograd's ones_like fill) drains
    pb2::PdlScope pdl;
    return pb2::hinge_finish2_ex(st.pv, st.pa, v, a, in_dtype, n, dim, ldv, lda, st.rinv_v, st.rinv_a, st.row_cnt, st.col_cnt,
                                 1.0f / ((float)n * (float)n), grad_out, pb2::HingeFold(), d_v, d_a, out_dtype, stream);
}
