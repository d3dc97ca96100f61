// Microbenchmark: issue rate of packed fma.rn.f32x2 on sm_100a, as a function of independent chains per warp,
// warps per scheduler, and whether the multiplier is warp-uniform (kernel parameter) or a per-thread register.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_ffma2 ubench_ffma2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int CH, bool UNI>
__global__ void k(float *out, float cu, int iters)
{
    u64 acc[CH], m[CH];
    const float t = threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = pk(t + i, t - i); m[i] = UNI ? pk(cu, cu) : pk(cu + t, cu - t); }
    const u64 add = pk(0.5f, 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < CH; i++) acc[i] = fma2(m[i], acc[i], add);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(acc[i])); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA2 mixed with integer (alu pipe) instructions: does the second cycle of an FFMA2 hide another issue?
template <int CH, int NI>
__global__ void kmix(float *out, float cu, int iters)
{
    u64 acc[CH]; unsigned z[4] = { threadIdx.x, 3u, 5u, 7u };
    const float t = threadIdx.x * 1e-3f;
    const u64 m = pk(cu, cu), add = pk(0.5f, 0.25f);
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = pk(t + i, t - i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < CH; i++) {
                acc[i] = fma2(m, acc[i], add);
#pragma unroll
                for (int q = 0; q < NI; q++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(z[(i + q) & 3]) : "r"(z[(i + q + 1) & 3]), "r"(0x9e3779b9u));
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { float2 v; asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(acc[i])); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(z[0] ^ z[1] ^ z[2] ^ z[3]);
}
template <int CH, int NI> void runmix(int warps_per_sm, float *out)
{
    const int iters = 20000, threads = 32 * warps_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kmix<CH, NI><<<148, threads>>>(out, 0.999f, 100);
    cudaEventRecord(e0);
    kmix<CH, NI><<<148, threads>>>(out, 0.999f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double ffma2_per_sched = (double)iters * 8 * CH * warps_per_sm / 4.0;
    printf("chains %2d  + %d LOP3 per FFMA2  warps/SM %2d : %.2f cycles per FFMA2 per scheduler\n", CH, NI, warps_per_sm, cycles / ffma2_per_sched);
}

template <int CH, bool UNI> void run(int warps_per_sm, float *out)
{
    const int iters = 20000, threads = 32 * warps_per_sm;     // one CTA per SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<CH, UNI><<<148, threads>>>(out, 0.999f, 100);
    cudaEventRecord(e0);
    k<CH, UNI><<<148, threads>>>(out, 0.999f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double ffma2_per_sched = (double)iters * 8 * CH * warps_per_sm / 4.0;
    printf("chains %2d  %s  warps/SM %2d : %.2f cycles per FFMA2 per scheduler (%.1f %% of 1 per 2 cycles)\n", CH, UNI ? "uniform mult " : "register mult",
           warps_per_sm, cycles / ffma2_per_sched, 200.0 * ffma2_per_sched / cycles);
}

int main()
{
    float *out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    run<1, true>(4, out); run<2, true>(4, out); run<4, true>(4, out); run<8, true>(4, out); run<16, true>(4, out);
    run<4, true>(8, out); run<8, true>(8, out); run<4, true>(16, out); run<8, true>(16, out);
    run<4, false>(4, out); run<8, false>(4, out); run<8, false>(8, out); run<8, false>(16, out);
    runmix<8, 0>(8, out); runmix<8, 1>(8, out); runmix<8, 2>(8, out); runmix<8, 1>(4, out); runmix<8, 1>(16, out);
    return 0;
}
