// How fast can ONE warp (and 2, 4 warps on one scheduler) issue FP32 work on sm_100a?  Independent FFMA / FFMA2 chains
// (ILP 1..8) in a long unrolled loop; cycles per instruction from clock64 on one SM.  A CTA of 32 * W * 4 threads puts W
// warps on each of the four schedulers; only warp 0 reports.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_issue tools/ubench_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int ILP, bool PACKED>
__global__ void k(float *out, long long *cyc, int n)
{
    float x[ILP]; u64 y[ILP];
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 0.001f + i; y[i] = (u64)__float_as_uint(x[i]) | ((u64)__float_as_uint(x[i] + 1.f) << 32); }
    const float c1 = 0.999f, c2 = 0.001f; const u64 p1 = 0x3f7fbe773f7fbe77ull, p2 = 0x3a83126f3a83126full;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) { if (PACKED) y[i] = fma2(y[i], p1, p2); else x[i] = fma1(x[i], c1, c2); }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < ILP; i++) acc += x[i] + __uint_as_float((unsigned)y[i]) + __uint_as_float((unsigned)(y[i] >> 32));
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int ILP, bool PACKED> void run(const char *name, float *o, long long *c)
{
    const int n = 4000;
    for (int w = 1; w <= 4; w *= 2) {
        k<ILP, PACKED><<<1, 128 * w>>>(o, c, n); k<ILP, PACKED><<<1, 128 * w>>>(o, c, n);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / ((double)n * 16 * ILP);
        printf("%-6s ILP %d, %d warp(s) per scheduler: %.2f cycles per instruction per warp, %.2f instructions per cycle per scheduler\n", name, ILP, w, per, w / per);
    }
}

int main()
{
    float *o; long long *c; cudaMalloc(&o, 4096 * 4); cudaMalloc(&c, 8);
    run<1, false>("FFMA", o, c); run<2, false>("FFMA", o, c); run<4, false>("FFMA", o, c); run<8, false>("FFMA", o, c);
    run<1, true>("FFMA2", o, c); run<2, true>("FFMA2", o, c); run<4, true>("FFMA2", o, c); run<8, true>("FFMA2", o, c);
    return 0;
}
