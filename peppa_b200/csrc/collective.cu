// Multi-GPU plumbing of the row-sharded gallery, host-language agnostic:
//
//  * peer memory (NVLink 5 / NVSwitch, one process per GPU on one node): pb2_ipc_export / pb2_ipc_open hand a device
//    buffer to the other ranks as a CUDA IPC handle, and pb2_peer_reduce is the dV reduce-scatter written as our own
//    kernel over peer memory -- every owner rank PULLS the partial gradient rows it owns from all ranks (P2P loads)
//    and sums them in a fixed order.  No NCCL kernel competes with the persistent tensor-core grids for SMs while
//    the step runs, the transfer is one HBM/NVLink-bound launch behind the last gradient GEMM, and the result is
//    deterministic.
//  * NCCL (the embedding all-gather, the merges of column counts / loss / recall hits; optionally the dV
//    reduce-scatter): pb2_nccl_* take the caller's ncclComm_t, so a non-torch host runs the sharded step through
//    this C ABI alone.  NCCL is resolved at run time (dlsym on the process, then dlopen("libnccl.so.2")): the
//    library has no link-time NCCL dependency and uses whichever NCCL the host already loaded.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "host_util.h"
#include "peppa_b200.h"

namespace pb2 {

// out[i] = sum_q src[q][i] (fixed order q = 0 .. n - 1), 16-byte vectors; src[q] may be peer memory
struct PeerPtrs {
    const float4* p[PB2_MAX_PEERS];
};
template <int kUnroll>
__global__ void __launch_bounds__(256)
    peer_reduce_kernel(const PeerPtrs src, int n_src, int64_t n_vec, float4* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n_vec; i0 += stride * kUnroll) {
        float4 acc[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        // all loads of one source in flight before the adds: NVLink latency is hidden by memory-level parallelism
        for (int q = 0; q < n_src; ++q) {
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t i = i0 + u * stride;
                v[u] = i < n_vec ? src.p[q][i] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                acc[u].x += v[u].x;
                acc[u].y += v[u].y;
                acc[u].z += v[u].z;
                acc[u].w += v[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n_vec) out[i] = acc[u];
        }
    }
}

// ------------------------------------------------------------------------------------------------- NCCL, late bound
typedef void* nccl_comm_t;
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_reducescatter_fn)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*nccl_group_fn)(void);
typedef const char* (*nccl_errstr_fn)(int);
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclFloat32 = 7, kNcclSum = 0 };  // nccl.h ncclDataType_t / ncclRedOp_t

struct Nccl {
    nccl_allgather_fn all_gather = nullptr;
    nccl_allreduce_fn all_reduce = nullptr;
    nccl_reducescatter_fn reduce_scatter = nullptr;
    nccl_group_fn group_start = nullptr, group_end = nullptr;
    nccl_errstr_fn err = nullptr;
    bool ok = false;
};
static const Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = RTLD_DEFAULT;
        if (!dlsym(h, "ncclAllGather")) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        n.all_gather = (nccl_allgather_fn)dlsym(h, "ncclAllGather");
        n.all_reduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
        n.reduce_scatter = (nccl_reducescatter_fn)dlsym(h, "ncclReduceScatter");
        n.group_start = (nccl_group_fn)dlsym(h, "ncclGroupStart");
        n.group_end = (nccl_group_fn)dlsym(h, "ncclGroupEnd");
        n.err = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
        n.ok = n.all_gather && n.all_reduce && n.reduce_scatter && n.group_start && n.group_end;
    });
    return n;
}
static int check_nccl(int rc, const char* what) {
    if (rc == 0) return PB2_OK;
    return set_error(PB2_ERR_CUDA, "%s: NCCL error %d (%s)", what, rc, nccl().err ? nccl().err(rc) : "?");
}

}  // namespace pb2

using namespace pb2;

// ------------------------------------------------------------------------------------------------------ peer memory
extern "C" int pb2_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out) {
    if (!ptr || !handle_out || !offset_out) return set_error(PB2_ERR_ARG, "ipc_export: null");
    static_assert(sizeof(cudaIpcMemHandle_t) == PB2_IPC_HANDLE_BYTES, "IPC handle size");
    // the handle names the ALLOCATION the pointer lies in (a caching allocator hands out pieces of larger segments)
    CUdeviceptr base = 0;
    size_t size = 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return set_error(PB2_ERR_CUDA, "ipc_export: cuMemGetAddressRange entry point unavailable");
    typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
    const CUresult r = ((range_fn)fn)(&base, &size, (CUdeviceptr)ptr);
    if (r != CUDA_SUCCESS) return set_error(PB2_ERR_CUDA, "ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
    cudaIpcMemHandle_t h;
    int rc = check_cuda(cudaIpcGetMemHandle(&h, (void*)base), "ipc_export (a buffer from cudaMalloc, not a virtual-memory mapping)");
    if (rc) return rc;
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((CUdeviceptr)ptr - base);
    return PB2_OK;
}

extern "C" int pb2_ipc_open(const void* handle, int64_t offset, void** base_out, void** ptr_out) {
    if (!handle || !base_out || !ptr_out || offset < 0) return set_error(PB2_ERR_ARG, "ipc_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* base = nullptr;
    // opened on the CURRENT device with lazy peer access: the pointer is usable by this device's kernels over NVLink
    int rc = check_cuda(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "ipc_open");
    if (rc) return rc;
    *base_out = base;
    *ptr_out = static_cast<char*>(base) + offset;
    return PB2_OK;
}

extern "C" int pb2_ipc_close(void* base) {
    if (!base) return PB2_OK;
    return check_cuda(cudaIpcCloseMemHandle(base), "ipc_close");
}

extern "C" int pb2_peer_reduce(const void* const* src, int n_src, int64_t n_elems, float* out, void* stream) {
    if (n_elems <= 0) return PB2_OK;
    if (!src || !out || n_src < 1 || n_src > PB2_MAX_PEERS) return set_error(PB2_ERR_ARG, "peer_reduce: 1 .. %d sources", PB2_MAX_PEERS);
    if (n_elems % 4 != 0 || (reinterpret_cast<uintptr_t>(out) & 15)) return set_error(PB2_ERR_ARG, "peer_reduce: whole 16-byte vectors");
    PeerPtrs p;
    for (int q = 0; q < PB2_MAX_PEERS; ++q) {
        p.p[q] = q < n_src ? static_cast<const float4*>(src[q]) : nullptr;
        if (q < n_src && (!src[q] || (reinterpret_cast<uintptr_t>(src[q]) & 15)))
            return set_error(PB2_ERR_ARG, "peer_reduce: source %d is null or not 16-byte aligned", q);
    }
    const int64_t n_vec = n_elems / 4;
    constexpr int kUnroll = 4;
    // NVLink-bound: enough CTAs to keep ~10^5 16-byte loads in flight, but never the whole machine
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_vec + 256 * kUnroll - 1) / (256 * kUnroll), (int64_t)sm_count() * 4));
    peer_reduce_kernel<kUnroll><<<grid, 256, 0, (cudaStream_t)stream>>>(p, n_src, n_vec, reinterpret_cast<float4*>(out));
    return check_launch("peer_reduce");
}

// ------------------------------------------------------------------------------------------------------------- NCCL
extern "C" int pb2_nccl_available(void) { return nccl().ok ? 1 : 0; }

// rank r's rows [n_local, ld] (any 1 / 2 / 4-byte element type: raw bytes) -> full [world * n_local, ld]
extern "C" int pb2_nccl_gallery_allgather(void* comm, const void* local_rows, int64_t n_local, int64_t row_bytes, void* full_out,
                                          void* stream) {
    if (n_local <= 0 || row_bytes <= 0) return PB2_OK;
    if (!comm || !local_rows || !full_out) return set_error(PB2_ERR_ARG, "nccl_gallery_allgather: null");
    if (!nccl().ok) return set_error(PB2_ERR_UNSUPPORTED, "NCCL not found in this process (libnccl.so.2)");
    return check_nccl(nccl().all_gather(local_rows, full_out, (size_t)(n_local * row_bytes), kNcclUint8, comm, (cudaStream_t)stream),
                      "nccl_gallery_allgather");
}

// merge of the per-rank column statistics of the hinge loss and of the scalars, in ONE NCCL group:
// col_cnt int32 [n_total] (sum), loss fp32 [1] (sum), hits fp32 [n_hits] (sum; recall@n numerators); any may be NULL
extern "C" int pb2_nccl_colstat_merge(void* comm, int32_t* col_cnt, int64_t n_total, float* loss, float* hits, int n_hits, void* stream) {
    if (!comm) return set_error(PB2_ERR_ARG, "nccl_colstat_merge: null communicator");
    if (!nccl().ok) return set_error(PB2_ERR_UNSUPPORTED, "NCCL not found in this process (libnccl.so.2)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = check_nccl(nccl().group_start(), "nccl_colstat_merge");
    if (rc) return rc;
    int e = 0;
    if (col_cnt && n_total > 0) e |= nccl().all_reduce(col_cnt, col_cnt, (size_t)n_total, kNcclInt32, kNcclSum, comm, st);
    if (loss) e |= nccl().all_reduce(loss, loss, 1, kNcclFloat32, kNcclSum, comm, st);
    if (hits && n_hits > 0) e |= nccl().all_reduce(hits, hits, (size_t)n_hits, kNcclFloat32, kNcclSum, comm, st);
    const int e2 = nccl().group_end();
    return check_nccl(e ? e : e2, "nccl_colstat_merge");
}

// dV partials [world * n_local, dim] fp32 on every rank -> this rank's rows [n_local, dim], summed
extern "C" int pb2_nccl_dv_reduce_scatter(void* comm, const float* partial_full, int64_t n_local, int dim, float* out_local,
                                          void* stream) {
    if (n_local <= 0 || dim <= 0) return PB2_OK;
    if (!comm || !partial_full || !out_local) return set_error(PB2_ERR_ARG, "nccl_dv_reduce_scatter: null");
    if (!nccl().ok) return set_error(PB2_ERR_UNSUPPORTED, "NCCL not found in this process (libnccl.so.2)");
    return check_nccl(nccl().reduce_scatter(partial_full, out_local, (size_t)(n_local * dim), kNcclFloat32, kNcclSum, comm,
                                            (cudaStream_t)stream),
                      "nccl_dv_reduce_scatter");
}
