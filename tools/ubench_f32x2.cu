// Microbenchmark: issue cost of packed fp32 (FFMA2/FADD2, sm_100a) against scalar FFMA/FADD, alone and mixed with
// integer work.  Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/ub tools/ubench_f32x2.cu && /tmp/ub
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float add1(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) { unsigned d; asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

// MODE 0: 16 scalar FFMA / iter; 1: 8 FFMA2 / iter (same flops); 2: 16 FADD; 3: 8 FADD2;
// 4: 16 FFMA + 16 IADD; 5: 8 FFMA2 + 16 IADD; 6: 16 IADD only
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float x[16]; u64 y[8]; unsigned z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x * 0.001f + i; z[i] = threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(x[2 * i], x[2 * i + 1]); y[i] = *reinterpret_cast<u64*>(&t); }
    float2 s2 = make_float2(s, s); const u64 sp = *reinterpret_cast<u64*>(&s2);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma1(x[i], s, s);
        }
        if (MODE == 1 || MODE == 5) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = fma2(y[i], sp, sp);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = add1(x[i], s);
        }
        if (MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = add2(y[i], sp);
        }
        if (MODE >= 4) {
#pragma unroll
            for (int i = 0; i < 16; ++i) z[i] = iadd(z[i], (unsigned)it);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += x[i] + (float)z[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&y[i]); acc += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* out, double flops_per_iter_thread) {
    const int iters = 20000, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, 100, 0.999f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<blocks, 256>>>(out, iters, 0.999f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double thr = (double)blocks * 256;
    printf("%-28s %8.3f ms  %7.2f Gelem-op/s per SM-clk-ish: %6.1f elem-ops/clk/SM @1.9GHz\n", name, ms,
           thr * iters * flops_per_iter_thread / ms * 1e-6, thr * iters * flops_per_iter_thread / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("16 FFMA", out, 16);
    run<1>("8 FFMA2", out, 16);
    run<2>("16 FADD", out, 16);
    run<3>("8 FADD2", out, 16);
    run<4>("16 FFMA + 16 IADD", out, 16);
    run<5>("8 FFMA2 + 16 IADD", out, 16);
    run<6>("16 IADD", out, 16);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
