// Host-side helpers shared by the C-ABI translation units: error reporting, device queries,
// TMA tensor-map encoding through the driver entry point (no link-time libcuda dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>

namespace pb2 {

int set_error(int code, const char* fmt, ...);      // records a thread-local message, returns code
int check_launch(const char* what);                 // cudaGetLastError -> PB2_OK / PB2_ERR_CUDA
int check_cuda(cudaError_t e, const char* what);
int sm_count();                                     // SMs of the current device (cached per device)

// 2-D row-major tensor map with 128-byte swizzle: inner dimension `cols` (contiguous), outer
// `rows`, row pitch `ld_bytes`; box = box_cols x box_rows elements; out-of-bounds reads give 0.
int make_tmap_2d(CUtensorMap* map, const void* base, int elem_bytes, uint64_t rows, uint64_t cols,
                 uint64_t ld_bytes, uint32_t box_rows, uint32_t box_cols);

}  // namespace pb2
