// Micro-benchmark: issue rate of scalar FFMA / FADD against the packed fma.rn.f32x2 / add.rn.f32x2 of sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(u64 v) { float2 r; asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

constexpr int ITERS = 4096, ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float s) {
    const float t = threadIdx.x * 1e-3f;
    if (MODE == 0) {            // scalar FFMA: 2*ILP independent chains = the same FLOPs as MODE 1
        float a[2 * ILP];
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) a[i] = t + i;
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 2 * ILP; ++i) a[i] = fmaf(a[i], s, t);
        float r = 0;
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 1) {     // packed FFMA2
        u64 a[ILP];
        const u64 s2 = pk(s, s), t2 = pk(t, t);
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = pk(t + i, t - i);
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma2(a[i], s2, t2);
        float r = 0;
#pragma unroll
        for (int i = 0; i < ILP; ++i) { float2 v = up(a[i]); r += v.x + v.y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else if (MODE == 2) {     // scalar FADD
        float a[2 * ILP];
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) a[i] = t + i;
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < 2 * ILP; ++i) a[i] = a[i] + s;
        float r = 0;
#pragma unroll
        for (int i = 0; i < 2 * ILP; ++i) r += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    } else {                    // packed FADD2
        u64 a[ILP];
        const u64 s2 = pk(s, s);
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = pk(t + i, t - i);
        for (int it = 0; it < ITERS; ++it)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = add2(a[i], s2);
        float r = 0;
#pragma unroll
        for (int i = 0; i < ILP; ++i) { float2 v = up(a[i]); r += v.x + v.y; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    }
}

template <int MODE>
void run(const char* name, float* out) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8;
    k<MODE><<<grid, 256>>>(out, 1.0001f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(out, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double lane_ops = 5.0 * grid * 256 * (double)ITERS * 2 * ILP;
    printf("%-12s %8.3f ms  %7.2f T lane-ops/s  (%s)\n", name, ms / 5, lane_ops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA", out); run<1>("FFMA2", out); run<2>("FADD", out); run<3>("FADD2", out);
    return 0;
}
