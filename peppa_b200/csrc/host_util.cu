#include "host_util.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <mutex>

#include "peppa_b200.h"

namespace pb2 {

static thread_local char g_err[512] = "";
thread_local int g_pdl_depth = 0;
thread_local int g_operands_ready_depth = 0;

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
const char* last_error() { return g_err; }

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PB2_OK;
    return set_error(PB2_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
static std::atomic<long long> g_launches{0};
int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return check_cuda(cudaGetLastError(), what);
}

int sm_count() {
    static int cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

// A training step at batch ~1k is launch bound (pig/models.py:262: the call the reference actually trains with), and
// every tcgen05 launch needs 2-4 tensor maps whose encoding is a driver call of ~1 us each on buffers that are the
// same step after step (the caller's batch tensors, the cached step workspace).  Encoded maps are therefore kept in a
// small per-thread direct-mapped cache keyed by everything the encoding depends on; a map is a pure function of its
// key, so a hit is always valid (no state of the pointed-to memory is captured).
struct TmapKey {
    const void* base;
    uint64_t rows, cols, ld_bytes;
    uint32_t box_rows, box_cols;
    int elem_bytes;
    bool operator==(const TmapKey& o) const {
        return base == o.base && rows == o.rows && cols == o.cols && ld_bytes == o.ld_bytes && box_rows == o.box_rows &&
               box_cols == o.box_cols && elem_bytes == o.elem_bytes;
    }
};
struct TmapSlot {
    TmapKey key;
    CUtensorMap map;
    bool valid = false;
};
constexpr int kTmapSlots = 64;

int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols, uint64_t ld_bytes,
                 uint32_t box_rows, uint32_t box_cols) {
    auto fn = encode_fn();
    if (!fn) return set_error(PB2_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld_bytes & 15))
        return set_error(PB2_ERR_ARG, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    static thread_local TmapSlot cache[kTmapSlots];
    const TmapKey key{base, rows, cols, ld_bytes, box_rows, box_cols, elem_bytes};
    uint64_t h = reinterpret_cast<uintptr_t>(base) >> 4;
    h ^= rows * 0x9E3779B97F4A7C15ull;
    h ^= (cols + ((uint64_t)box_rows << 20) + ((uint64_t)box_cols << 40)) * 0xC2B2AE3D27D4EB4Full;
    TmapSlot& slot = cache[(h ^ (h >> 29)) % kTmapSlots];
    if (slot.valid && slot.key == key) {
        *map = slot.map;
        return PB2_OK;
    }
    // the data type only matters for out-of-bounds fill and interleaving, neither of which is used: 16-bit data of
    // either format travels as BFLOAT16, bytes as UINT8
    CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                             : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    // boxes with 128-byte rows use the 128-byte swizzle (every operand and output slab); a box with 64-byte rows (the
    // encoder tail's half slabs) the 64-byte one: 16-byte piece index bits [4:5] ^= address bits [7:8]
    const CUtensorMapSwizzle sw = (uint64_t)box_cols * (uint64_t)elem_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(PB2_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    slot.key = key;
    slot.map = *map;
    slot.valid = true;
    return PB2_OK;
}

}  // namespace pb2

extern "C" const char* pb2_last_error(void) { return pb2::last_error(); }
extern "C" int pb2_version(void) { return PB2_VERSION; }
extern "C" long long pb2_launch_count(void) { return pb2::g_launches.load(std::memory_order_relaxed); }
