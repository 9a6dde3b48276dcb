// Probe for the round-2 lever of DESIGN section 8 item 2 (one-byte gradient matrix): does tcgen05.mma.kind::i8 run on
// this B200 (sm_100a), is u8 x s8 -> s32 exact with the K-major 128B-swizzled operand layout the kernels already use,
// and what is its issue rate next to kind::f16 on the same bytes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I peppa_b200/csrc -o tools/ubench/i8mma tools/ubench/i8mma.cu
//   tools/ubench/i8mma
// A [128 x 128] u8 in {0,1,2} (the hinge gradient matrix' values), B [128 x 128] s8, D = A B^T [128 x 128] s32:
// four MMAs of K = 32 (32 bytes of a 128-byte swizzled row per instruction, like K = 16 for 16-bit types).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace pb2;

// instruction descriptor: [4,6) D format (1 = f32, 2 = s32), [7,10) A format, [10,13) B format (kind::i8: 0 = u8,
// 1 = s8; kind::f16: 0 = f16, 1 = bf16), bit 15 / 16 A / B major (0 = K), [17,23) N >> 3, [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_of(uint32_t dfmt, uint32_t afmt, uint32_t bfmt, uint32_t m, uint32_t n) {
    return (dfmt << 4) | (afmt << 7) | (bfmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <bool kI8>
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (kI8)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
                     "l"(da), "l"(db), "r"(idesc), "r"(acc)
                     : "memory");
}

// reps == 0: one product, D written to out (correctness).  reps > 0: reps x 4 MMAs back to back (issue rate).
// reps < 0: correctness with B given TRANSPOSED, bt [K = 128][N = 128] row-major, and read MN-major (the layout of the
// embedding operand of the gradient GEMMs): 128 k-rows of 128 bytes, 128B swizzle on (k-row % 8), descriptor SBO =
// 1024 (8 k-rows), 32 k-rows = 4096 bytes per instruction, B-major bit set; -reps - 1 = LBO >> 4 to try.
template <bool kI8>
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* a, const uint8_t* b, int32_t* out, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sa = smem;
    uint8_t* sb = smem + 128 * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 128 * 128);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // row r of the [128 x 128-byte] tile: 16-byte chunk c lands at chunk (c ^ (r & 7)) of the row (128B swizzle)
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a + r * 128 + c * 16);
        *reinterpret_cast<uint4*>(sb + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * 128 + c * 16);
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(slot, 128);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (tid == 0) {
        uint32_t idesc = kI8 ? idesc_of(2, 0, 1, 128, 128) : idesc_of(1, 1, 1, 128, 128);
        const uint64_t da = make_smem_desc(smem_u32(sa), 16, 1024);
        if (reps < 0) {
            idesc |= 1u << 16;  // B is MN-major
            const uint32_t code = (uint32_t)(-reps - 1);  // (LBO >> 4) * 4096 + (SBO >> 4)
            const uint64_t db = make_smem_desc(smem_u32(sb), (code >> 12) << 4, (code & 4095u) << 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) mma<kI8>(tmem, da + 2 * k, db + 256 * k, idesc, k != 0 ? 1u : 0u);  // +32 k-rows per step
        } else {
            const uint64_t db = make_smem_desc(smem_u32(sb), 16, 1024);
            const int n = reps > 0 ? reps : 1;
            for (int it = 0; it < n; ++it)
#pragma unroll
                for (int k = 0; k < 4; ++k) mma<kI8>(tmem, da + 2 * k, db + 2 * k, idesc, (it | k) != 0 ? 1u : 0u);  // +32 bytes per step
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    if (reps <= 0) {
        for (int ch = 0; ch < 4; ++ch) {
            uint32_t v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + ch * 32, v);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 128 + ch * 32 + j] = (int32_t)v[j];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

int main() {
    std::vector<uint8_t> ha(128 * 128), hb(128 * 128);
    srand(7);
    for (auto& x : ha) x = (uint8_t)(rand() % 3);
    for (auto& x : hb) x = (uint8_t)(int8_t)(rand() % 256 - 128);
    uint8_t *da, *db;
    int32_t* dout;
    cudaMalloc(&da, ha.size());
    cudaMalloc(&db, hb.size());
    cudaMalloc(&dout, 128 * 128 * 4);
    cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice);
    const int smem = 2 * 128 * 128 + 64;
    cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<true><<<1, 128, smem>>>(da, db, dout, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("kind::i8 probe failed: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<int32_t> ho(128 * 128);
    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int i = 0; i < 128; ++i)
        for (int j = 0; j < 128; ++j) {
            int32_t ref = 0;
            for (int k = 0; k < 128; ++k) ref += (int32_t)ha[i * 128 + k] * (int32_t)(int8_t)hb[j * 128 + k];
            bad += ref != ho[i * 128 + j];
        }
    printf("kind::i8 u8 x s8 -> s32, 128 x 128 x 128, K-major 128B swizzle: %ld mismatches (D[0][0..3] = %d %d %d %d)\n", bad, ho[0], ho[1],
           ho[2], ho[3]);
    // B transposed in memory, read MN-major
    std::vector<uint8_t> hbt(128 * 128);
    for (int k = 0; k < 128; ++k)
        for (int j = 0; j < 128; ++j) hbt[k * 128 + j] = hb[j * 128 + k];
    uint8_t* dbt;
    cudaMalloc(&dbt, hbt.size());
    cudaMemcpy(dbt, hbt.data(), hbt.size(), cudaMemcpyHostToDevice);
    // (LBO, SBO) candidates in bytes: with one 128-element MN chunk the stride between 8-k-row groups (1024 B) is
    // the one that matters; which field carries it is what the sweep answers
    // (measured: SBO = 1024 is exact for every LBO; SBO = 16 gives 16378 mismatches; SBO = 4096 walks out of the tile)
    const int cand[][2] = {{16, 1024}, {1024, 1024}, {4096, 1024}, {16384, 1024}};
    for (auto& c2 : cand) {
        const int lbo = c2[0] >> 4, sbo = c2[1] >> 4;
        cudaMemset(dout, 0xff, 128 * 128 * 4);
        probe<true><<<1, 128, smem>>>(da, dbt, dout, -(lbo * 4096 + sbo) - 1);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("MN-major B probe (LBO %d, SBO %d) failed: %s\n", lbo << 4, sbo << 4, cudaGetErrorString(e));
            return 1;
        }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        long badt = 0;
        for (int i = 0; i < 128; ++i)
            for (int j = 0; j < 128; ++j) {
                int32_t ref = 0;
                for (int k = 0; k < 128; ++k) ref += (int32_t)ha[i * 128 + k] * (int32_t)(int8_t)hb[j * 128 + k];
                badt += ref != ho[i * 128 + j];
            }
        printf("kind::i8, B [K][N] read MN-major (+4096 B per K = 32; LBO %5d, SBO %5d): %ld mismatches\n", lbo << 4, sbo << 4, badt);
    }
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0);
    cudaEventCreate(&t1);
    const int reps = 20000;
    for (int kind = 0; kind < 2; ++kind) {
        for (int w = 0; w < 2; ++w) {  // warm-up, then timed
            cudaEventRecord(t0);
            if (kind) probe<true><<<sms, 128, smem>>>(da, db, dout, reps);
            else probe<false><<<sms, 128, smem>>>(da, db, dout, reps);
            cudaEventRecord(t1);
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("timing run failed: %s\n", cudaGetErrorString(e));
                return 1;
            }
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t0, t1);
        const double macs = (double)sms * reps * 4 * 128.0 * 128.0 * (kind ? 32.0 : 16.0);
        printf("%s: %d SMs x %d x 4 MMAs (128 x 128 x %d): %.3f ms  %.1f T%s/s\n", kind ? "kind::i8 " : "kind::f16", sms, reps, kind ? 32 : 16,
               ms, 2.0 * macs / ms / 1e9, kind ? "OP" : "FLOP");
    }
    return bad != 0;
}
