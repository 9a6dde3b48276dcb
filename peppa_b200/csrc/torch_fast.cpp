// Thin torch extension over the C ABI for the batch-~1k training step (BASELINE config 2; the call the reference
// trains with: pig/models.py:262 -> pig/loss.py:33-39 + autograd).  That step is HOST bound: its kernels take ~28 us,
// the Python glue of a torch.autograd.Function (ctypes marshalling of 17 arguments, tensor allocations, the engine's
// round trip through Python for backward) ~140 us.  This file is the same glue in C++: one autograd node whose forward is
// ONE call of pb2_hinge_forward (three launches) and whose backward is ONE call of pb2_hinge_backward (one launch).  No
// kernels here and no arithmetic: everything numeric is behind include/peppa_b200.h.  peppa_b200/loss.py takes this path when the inputs
// need no conversion (CUDA, 2-D, one dtype of bf16 / fp16 / fp32, contiguous rows, D % 64 == 0, N <= 32768) and the
// ctypes path otherwise; both end in the same entry points.
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>

#include <mutex>
#include <vector>

#include "peppa_b200.h"

namespace {

int dtype_code(at::ScalarType t) {
    switch (t) {
        case at::kBFloat16: return PB2_BF16;
        case at::kHalf: return PB2_F16;
        case at::kFloat: return PB2_F32;
        default: TORCH_CHECK(false, "peppa_b200: embeddings are bf16, fp16 or fp32");
    }
}

[[noreturn]] void fail(const char* what, int rc) {
    const char* msg = pb2_last_error();
    TORCH_CHECK(false, "peppa_b200 ", what, " failed (status ", rc, "): ", msg ? msg : "");
}

// scratch of the fused step, reused across steps ON ONE STREAM (reuse is stream-ordered); a workspace handed to a graph
// capture belongs to that graph and is not cached -- the rules of ops.hinge_step
struct WsEntry {
    int device;
    void* stream;
    int64_t n, d;
    int code;
    at::Tensor buf;
};
std::mutex g_ws_mutex;
std::vector<WsEntry> g_ws;

at::Tensor step_workspace(const at::Tensor& like, void* stream, int64_t n, int64_t d, int code, bool capturing) {
    const int dev = like.get_device();
    if (!capturing) {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        for (const auto& e : g_ws)
            if (e.device == dev && e.stream == stream && e.n == n && e.d == d && e.code == code) return e.buf;
    }
    const int64_t bytes = pb2_hinge_step_workspace(n, (int)d, code);
    at::Tensor buf = at::empty({bytes}, like.options().dtype(at::kByte));
    if (!capturing) {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        if (g_ws.size() > 8) g_ws.clear();
        g_ws.push_back({dev, stream, n, d, code, buf});
    }
    return buf;
}

class HingeStepFn : public torch::autograd::Function<HingeStepFn> {
 public:
    static at::Tensor forward(torch::autograd::AutogradContext* ctx, const at::Tensor& V, const at::Tensor& A, double margin,
                              const c10::optional<at::Tensor>& rinv_v, const c10::optional<at::Tensor>& rinv_a) {
        TORCH_CHECK(V.dim() == 2 && A.dim() == 2 && V.sizes() == A.sizes(), "TripletLoss expects V [N, D] and A [N, D]");
        TORCH_CHECK(V.is_cuda() && A.is_cuda() && V.get_device() == A.get_device() && V.scalar_type() == A.scalar_type(),
                    "peppa_b200 fast path: one CUDA device, one dtype");
        const int64_t n = V.size(0), d = V.size(1);
        TORCH_CHECK(V.stride(1) == 1 && A.stride(1) == 1 && d % 64 == 0 && n <= 32768, "peppa_b200 fast path: layout");
        const int code = dtype_code(V.scalar_type());
        const c10::cuda::CUDAGuard guard(V.device());
        const auto stream = c10::cuda::getCurrentCUDAStream(V.get_device());
        const bool capturing = c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None;
        at::Tensor ws = step_workspace(V, (void*)stream.stream(), n, d, code, capturing);
        // what the backward needs (both gradient products, 1/||row||, the indicator counts) stays in a buffer of this
        // call's own; grad_output is applied there in fp32 before the rounding (an AMP GradScaler's 65536 must reach an
        // fp16 gradient of ~1e-7 before the rounding does)
        at::Tensor state = at::empty({pb2_hinge_state_bytes(n, (int)d)}, V.options().dtype(at::kByte));
        at::Tensor loss = at::empty({}, V.options().dtype(at::kFloat));
        const float* rv = rinv_v.has_value() && rinv_v->defined() ? rinv_v->data_ptr<float>() : nullptr;
        const float* ra = rinv_a.has_value() && rinv_a->defined() ? rinv_a->data_ptr<float>() : nullptr;
        const int rc = pb2_hinge_forward(V.data_ptr(), A.data_ptr(), code, n, (int)d, V.stride(0), A.stride(0), (float)margin,
                                         ws.data_ptr(), ws.numel(), state.data_ptr(), state.numel(), loss.data_ptr<float>(), rv,
                                         ra, (void*)stream.stream());
        if (rc != PB2_OK) fail("hinge_forward", rc);
        ctx->save_for_backward({state, V, A});
        ctx->saved_data["code"] = (int64_t)code;
        return loss;
    }

    static torch::autograd::variable_list backward(torch::autograd::AutogradContext* ctx,
                                                   torch::autograd::variable_list grad_outputs) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor &state = saved[0], &V = saved[1], &A = saved[2];
        const int code = (int)ctx->saved_data["code"].toInt();
        const int64_t n = V.size(0), d = V.size(1);
        at::Tensor go = grad_outputs[0];
        if (!go.is_cuda() || go.get_device() != state.get_device() || go.scalar_type() != at::kFloat)
            go = go.to(state.options().dtype(at::kFloat));
        go = go.contiguous();
        const c10::cuda::CUDAGuard guard(state.device());
        const auto stream = c10::cuda::getCurrentCUDAStream(state.get_device());
        // two fresh contiguous tensors in the inputs' dtype: AccumulateGrad takes them without a copy
        at::Tensor g0 = at::empty({n, d}, V.options()), g1 = at::empty({n, d}, V.options());
        const int rc = pb2_hinge_backward(state.data_ptr(), state.numel(), V.data_ptr(), A.data_ptr(), code, n, (int)d, V.stride(0),
                                          A.stride(0), go.data_ptr<float>(), g0.data_ptr(), g1.data_ptr(), code,
                                          (void*)stream.stream());
        if (rc != PB2_OK) fail("hinge_backward", rc);
        return {g0, g1, at::Tensor(), at::Tensor(), at::Tensor()};
    }
};

at::Tensor triplet_loss(const at::Tensor& V, const at::Tensor& A, double margin, const c10::optional<at::Tensor>& rinv_v,
                        const c10::optional<at::Tensor>& rinv_a) {
    return HingeStepFn::apply(V, A, margin, rinv_v, rinv_a);
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "peppa_b200: C++ autograd glue over the C ABI for the launch-bound training step";
    m.def("triplet_loss", &triplet_loss, "TripletLoss forward (+ autograd node) through pb2_hinge_forward / pb2_hinge_backward",
          py::arg("V"), py::arg("A"), py::arg("margin"), py::arg("rinv_v") = py::none(), py::arg("rinv_a") = py::none());
}
