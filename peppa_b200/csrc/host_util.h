// Host-side helpers shared by the C-ABI translation units: error reporting, device queries,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <utility>

namespace pb2 {

int set_error(int code, const char* fmt, ...);      // records a thread-local message, returns code
int check_launch(const char* what);                 // cudaGetLastError -> PB2_OK / PB2_ERR_CUDA
int check_cuda(cudaError_t e, const char* what);
int sm_count();                                     // SMs of the current device (cached per device)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a property of (kernel, DEVICE): a process that launches on
// cuda:0 and then on cuda:1 has to opt in on both.  One bit per device, lock-free; a lost race only repeats the
// (idempotent) attribute call.  Devices >= 64 set the attribute on every launch.
struct PerDeviceOnce {
    std::atomic<uint64_t> bits{0};
    static int device() {
        int dev = 0;
        return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
    }
    bool done(int dev) const { return dev >= 0 && dev < 64 && ((bits.load(std::memory_order_acquire) >> dev) & 1u); }
    void mark(int dev) {
        if (dev >= 0 && dev < 64) bits.fetch_or(1ull << dev, std::memory_order_release);
    }
};
// opt a kernel in to `smem` bytes of dynamic shared memory on the current device (once per device and call site)
template <class K>
inline int ensure_dynamic_smem(PerDeviceOnce& once, K kern, int smem, const char* what) {
    const int dev = PerDeviceOnce::device();
    if (once.done(dev)) return 0;
    const int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), what);
    if (rc == 0) once.mark(dev);
    return rc;
}

// The scalar hinge loss of a training step from what the similarity pass left behind (device code: fold.cuh):
// loss = coef * (sum of the CTA partials + sum_i (margin - diag_i) (row_cnt_i + col_cnt_i)), NaN if any 1/||row|| is
// not finite (the reference's 0/0).  Runs in 256 threads of ONE block: block 0 of hinge_finish2, or a spare CTA of the
// dual gradient-product grid (pb2_hinge_forward).  loss_out == nullptr: no fold.
struct HingeFold {
    const float* loss_partial = nullptr;
    int n_partials = 0;
    const float* diag = nullptr;
    const int32_t* row_cnt = nullptr;
    const int32_t* col_cnt = nullptr;
    const float* rinv_v = nullptr;
    const float* rinv_a = nullptr;
    int64_t n = 0;
    float margin = 0.f, coef = 0.f;
    float* loss_out = nullptr;
};
// pb2_grad_gemm_dual with the loss fold in a CTA of its own at the end of the grid (gradgemm.cu).  *folded says whether
// the launch took the fold (it does whenever both products fit one grid); otherwise the caller launches hinge_fold.
int grad_gemm_dual_fold(const void* gmat, int g_dtype, int64_t g_rows, int64_t g_cols, int64_t ld_g, const void* z0, const void* z1,
                        int z_dtype, int dim, int64_t ldz0, int64_t ldz1, float alpha, float* out0, float* out1, int64_t ld_out0,
                        int64_t ld_out1, const HingeFold& fold, bool* folded, void* stream);
int hinge_fold(const HingeFold& fold, void* stream);  // the fold as a launch of its own (rowstats.cu)
// pb2_hinge_finish2 with an optional device-side factor (autograd's grad_output, applied in fp32 before the rounding)
// and an optional loss fold (fold.loss_out != nullptr: block 0 of the grid) (rowstats.cu)
int hinge_finish2_ex(const float* p_v, const float* p_a, const void* v, const void* a, int dtype, int64_t n, int dim, int64_t ldv,
                     int64_t lda, const float* rinv_v, const float* rinv_a, const int32_t* row_cnt, const int32_t* col_cnt,
                     float coef, const float* coef_dev, const HingeFold& fold, void* d_v, void* d_a, int out_dtype, void* stream);

// 2-D row-major tensor map with 128-byte swizzle: inner dimension `cols` (contiguous), outer
// `rows`, row pitch `ld_bytes`; box = box_cols x box_rows elements; out-of-bounds reads give 0.
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_bytes, uint32_t box_rows, uint32_t box_cols);

// Measurement-only switches (tile-shape / cluster-variant selectors, wrong-result knock-outs) exist only in the
// measurement build (-DPB2_MEASURE -> libpeppa_b200_measure.so, used by tools/ and the variant tests).  In the
// product library they are compile-time constants: no pb2_debug_* export, no mutable global state.
#ifdef PB2_MEASURE
#define PB2_KNOB static int
#define PB2_KNOB_U32 static uint32_t
#else
#define PB2_KNOB static constexpr int
#define PB2_KNOB_U32 static constexpr uint32_t
#endif

// Launch helper: optional thread-block cluster, and -- while a PdlScope is alive on this thread -- the
// programmatic-stream-serialization attribute (programmatic dependent launch): the kernel may start while its
// predecessor in the stream drains, runs its prologue (barrier init, TMEM allocation, descriptor prefetch) and
// blocks in griddepcontrol.wait (pdl_wait() in common.cuh) until the predecessor's memory is visible.  Every
// kernel launched through here calls pdl_wait() before it touches global memory.
extern thread_local int g_pdl_depth;
struct PdlScope {
    PdlScope() { ++g_pdl_depth; }
    ~PdlScope() { --g_pdl_depth; }
};
// While an OperandsReadyScope is alive on this thread, the similarity pass launched through here may fetch its X / Y
// operand tiles BEFORE griddepcontrol.wait: the caller promises that they were complete before the pass's predecessor
// in the stream passed its own wait (pb2_hinge_step: the raw input rows, with hinge_prep waiting before it triggers).
extern thread_local int g_operands_ready_depth;
struct OperandsReadyScope {
    explicit OperandsReadyScope(bool on) : on_(on) { g_operands_ready_depth += on_; }
    ~OperandsReadyScope() { g_operands_ready_depth -= on_; }
    OperandsReadyScope(const OperandsReadyScope&) = delete;
    OperandsReadyScope& operator=(const OperandsReadyScope&) = delete;
   private:
    int on_;
};

template <class... KArgs, class... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, int cluster,
                             Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = (unsigned)cluster;
        attr[n].val.clusterDim.y = 1;
        attr[n].val.clusterDim.z = 1;
        ++n;
    }
    if (g_pdl_depth > 0) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

}  // namespace pb2
