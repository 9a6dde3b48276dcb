// Issue-rate microbenchmark for the instructions of the hinge epilogue (sm_100a): cycles per warp
// instruction per SM sub-partition with 1, 2, 3, 4 warps per sub-partition, 8 independent chains per warp.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 4096
enum Op { FFMA, FFMA_SAT, FFMA2, FADD2, FMUL2, F2FP, REDUX, LOP3, FSET, MIX_FMA_ALU, FMNMX, IADD3, F2I, MIX_SAT_FADD2 };

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
    float a[8], b = seed, c = seed * 0.5f;
    float2 p[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i; p[i] = make_float2(seed + i, seed - i); u[i] = threadIdx.x * 7 + i; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (OP == FFMA_SAT) asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (OP == FFMA2) p[i] = __ffma2_rn(p[i], make_float2(b, b), make_float2(c, c));
            if (OP == FADD2) p[i] = __fadd2_rn(p[i], make_float2(b, c));
            if (OP == FMUL2) p[i] = __fmul2_rn(p[i], make_float2(b, c));
            if (OP == F2FP) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(__uint_as_float(u[i]))); }
            if (OP == REDUX) u[i] = __reduce_add_sync(0xffffffffu, u[i]);
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)));
            if (OP == FSET) asm volatile("{.reg .pred q; setp.ge.f32 q, %0, %1; selp.f32 %0, %2, %0, q;}" : "+f"(a[i]) : "f"(b), "f"(c));
            if (OP == FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(__float_as_uint(b)));
            if (OP == F2I) asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(u[i]) : "f"(a[i] + __uint_as_float(u[i])));
            if (OP == MIX_FMA_ALU) {
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(__float_as_uint(b)), "r"(__float_as_uint(c)));
            }
            if (OP == MIX_SAT_FADD2) {
                asm volatile("fma.rn.sat.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                p[i] = __fadd2_rn(p[i], make_float2(b, c));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + p[i].x + p[i].y + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_iter) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    printf("%-14s", name);
    for (int wps = 1; wps <= 4; ++wps) {  // warps per sub-partition
        k<OP><<<148, wps * 128>>>(out, cyc, 1.0001f);
        cudaDeviceSynchronize();
        k<OP><<<148, wps * 128>>>(out, cyc, 1.0001f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double per = (double)c / ((double)ITERS * 8 * instr_per_iter * wps);  // cycles per warp-instr per SMSP
        printf("  %dw/smsp: %.2f cyc/instr", wps, per);
    }
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<FFMA>("FFMA", 1); run<FFMA_SAT>("FFMA.SAT", 1); run<FFMA2>("FFMA2", 1); run<FADD2>("FADD2", 1); run<FMUL2>("FMUL2", 1);
    run<F2FP>("F2FP", 1); run<REDUX>("REDUX", 1); run<LOP3>("LOP3", 1); run<FSET>("FSETP+SEL", 2); run<FMNMX>("FMNMX", 1);
    run<IADD3>("IADD", 1); run<F2I>("F2I(+FADD)", 2); run<MIX_FMA_ALU>("FFMA+LOP3", 2); run<MIX_SAT_FADD2>("SAT+FADD2", 2);
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
