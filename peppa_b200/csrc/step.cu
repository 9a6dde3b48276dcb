// pb2_hinge_step: the whole TripletLoss forward + gradients of one training step (pig/models.py:262 ->
// pig/loss.py:33-48 + autograd) as ONE C call = four kernel launches on the caller's stream:
//   hinge_prep -> sim_hinge (tcgen05, fused loss/counts/gradient matrix) -> grad_gemm_dual (both backward
//   products in one grid) -> hinge_finish2, chained with programmatic dependent launch.
// At batch ~1k the step is launch bound; one call keeps the host side to a single FFI crossing and the
// scratch in one caller-provided workspace.
#include <cuda_fp16.h>

#include "host_util.h"
#include "peppa_b200.h"

namespace {
constexpr int64_t kAlign = 256;
int64_t up(int64_t x) { return (x + kAlign - 1) / kAlign * kAlign; }
struct Layout {
    int64_t rinv_v, rinv_a, diag, part, row_cnt, col_cnt, vh, ah, g, pv, pa, vx, ax, sv, sa, total, ld_g;
    int n_part;
};
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
    L.rinv_v = take(n * 4);
    L.rinv_a = take(n * 4);
    L.diag = take(n * 4);
    L.part = take((int64_t)L.n_part * 4);
    L.row_cnt = take(n * 4);
    L.col_cnt = take(n * 4);
    L.vh = take(n * dim * 2);
    L.ah = take(n * dim * 2);
    L.g = take(n * L.ld_g * 2);
    L.pv = take(n * dim * 4);
    L.pa = take(n * dim * 4);
    // fp32 inputs: split-fp16 tensor-core operands [n, 3 dim] and their per-row power-of-two scales (pb2_split_f16)
    L.vx = in_dtype == PB2_F32 ? take(n * 3 * dim * 2) : -1;
    L.ax = in_dtype == PB2_F32 ? take(n * 3 * dim * 2) : -1;
    L.sv = in_dtype == PB2_F32 ? take(n * 4) : -1;
    L.sa = in_dtype == PB2_F32 ? take(n * 4) : -1;
    L.total = o;
    return L;
}
PB2_KNOB g_step_stages = 4;  // measurement build (pb2_debug_step_stages): stop the step after its first k kernels
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
    return n > 0 && dim > 0 ? layout(n, dim, in_dtype).total : 0;
}

extern "C" int pb2_hinge_step(const void* v, const void* a, int in_dtype, int64_t n, int dim, int64_t ldv, int64_t lda, float margin,
                              void* workspace, int64_t workspace_bytes, float* loss_out, void* d_v, void* d_a,
                              int out_dtype, const float* rinv_v_in, const float* rinv_a_in, void* stream) {
    using pb2::set_error;
    if (n <= 0) return set_error(PB2_ERR_ARG, "hinge_step: empty batch");
    if (!v || !a || !workspace || !loss_out || !d_v || !d_a) return set_error(PB2_ERR_ARG, "hinge_step: null");
    if (dim <= 0 || dim % 64 != 0) return set_error(PB2_ERR_ARG, "hinge_step: dim must be a positive multiple of 64");
    if ((reinterpret_cast<uintptr_t>(workspace) & (kAlign - 1)) != 0)
        return set_error(PB2_ERR_ARG, "hinge_step: workspace must be 256-byte aligned");
    if (in_dtype != PB2_BF16 && in_dtype != PB2_F16 && in_dtype != PB2_F32)
        return set_error(PB2_ERR_ARG, "hinge_step: inputs are bf16, fp16 or fp32 rows");
    const Layout L = layout(n, dim, in_dtype);
    if (workspace_bytes < L.total) return set_error(PB2_ERR_ARG, "hinge_step: workspace too small");
    char* w = static_cast<char*>(workspace);
    float* rinv_v = reinterpret_cast<float*>(w + L.rinv_v);
    float* rinv_a = reinterpret_cast<float*>(w + L.rinv_a);
    float* diag = reinterpret_cast<float*>(w + L.diag);
    float* part = reinterpret_cast<float*>(w + L.part);
    int32_t* row_cnt = reinterpret_cast<int32_t*>(w + L.row_cnt);
    int32_t* col_cnt = reinterpret_cast<int32_t*>(w + L.col_cnt);
    void* vh = w + L.vh;
    void* ah = w + L.ah;
    void* g = w + L.g;
    float* pv = reinterpret_cast<float*>(w + L.pv);
    float* pa = reinterpret_cast<float*>(w + L.pa);
    // programmatic dependent launch between the four kernels: each one's prologue (barrier init, TMEM
    // allocation, descriptor prefetch) overlaps its predecessor's tail
    pb2::PdlScope pdl;
    // tensor-core operands: bf16 / fp16 rows as they are; fp32 rows as the split-fp16 pair of their normalised,
    // power-of-two-scaled values (contraction length 3 dim), with the scales in place of 1/||row|| in the epilogue
    const bool split = in_dtype == PB2_F32;
    void* vx = split ? w + L.vx : nullptr;
    void* ax = split ? w + L.ax : nullptr;
    float* sv = split ? reinterpret_cast<float*>(w + L.sv) : nullptr;
    float* sa = split ? reinterpret_cast<float*>(w + L.sa) : nullptr;
    int rc = pb2_hinge_prep(v, a, in_dtype, n, dim, ldv, lda, rinv_v, rinv_a, diag, vh, ah, row_cnt, col_cnt, part, L.n_part,
                            vx, ax, sv, sa, rinv_v_in, rinv_a_in, stream);
    if (rc || g_step_stages < 2) return rc;
    // bf16 / fp16 rows are the caller's own tensors: complete before hinge_prep (which waits BEFORE it triggers its
    // dependents) got past its wait, so the similarity pass streams them through the tensor cores while hinge_prep
    // still computes the norms and the diagonal its epilogue needs.  The split operands of fp32 rows are hinge_prep's
    // own output: no early start there.
    {
        pb2::OperandsReadyScope early(!split);
        rc = pb2_sim_hinge(split ? vx : v, split ? ax : a, split ? sv : rinv_v, split ? sa : rinv_a, diag, diag, n, n, 0, 0,
                           split ? 3 * dim : dim, split ? PB2_F16 : in_dtype, split ? 3 * (int64_t)dim : ldv,
                           split ? 3 * (int64_t)dim : lda, margin, part,
                           -L.n_part, row_cnt, col_cnt, g, PB2_F16, L.ld_g, nullptr, nullptr, stream);
    }
    if (rc || g_step_stages < 3) return rc;
    // dV partials = G A^, dA partials = G^T V^: one launch when all their tiles fit the machine at once
    rc = pb2_grad_gemm_dual(g, PB2_F16, n, n, L.ld_g, ah, vh, PB2_F16, dim, dim, dim, 1.0f, pv, pa, dim, dim, stream);
    if (rc || g_step_stages < 4) return rc;
    return pb2_hinge_finish2(pv, pa, v, a, in_dtype, n, dim, ldv, lda, rinv_v, rinv_a, diag, row_cnt, col_cnt, part, L.n_part, margin,
                             1.0f / ((float)n * (float)n), loss_out, d_v, d_a, out_dtype, stream);
}
