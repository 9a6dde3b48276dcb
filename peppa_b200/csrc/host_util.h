// Host-side helpers shared by the C-ABI translation units: error reporting, device queries,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <utility>

namespace pb2 {

int set_error(int code, const char* fmt, ...);      // records a thread-local message, returns code
int check_launch(const char* what);                 // cudaGetLastError -> PB2_OK / PB2_ERR_CUDA
int check_cuda(cudaError_t e, const char* what);
int sm_count();                                     // SMs of the current device (cached per device)

// 2-D row-major tensor map with 128-byte swizzle: inner dimension `cols` (contiguous), outer
// `rows`, row pitch `ld_bytes`; box = box_cols x box_rows elements; out-of-bounds reads give 0.
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_bytes, uint32_t box_rows, uint32_t box_cols);

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
