// Latency of the single-precision gain loop's dependent chain (devmath.cuh agc_step_fast), one warp, constant input:
// cycles per step for the whole step and for pieces of it.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdio>
#include <cuda_runtime.h>
#include "../python-liquiddsp_b200/csrc/devmath.cuh"
using namespace lqb;

template <int MODE>
__global__ void k(float *out, long long *cyc, int n, AgcFast kk, float xr, float xi)
{
    float g = 1.f, y2p = 1.f, acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (MODE == 0) { float2 y = agc_step_fast<false>(make_float2(xr, xi), g, y2p, kk); acc += y.x; }
        if (MODE == 1) { float2 y = agc_step_fast<true>(make_float2(xr, xi), g, y2p, kk); acc += y.x; }
        if (MODE == 2) { float l2; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(g)); g = l2 + 2.0f; }        // MUFU + FADD
        if (MODE == 3) { g = __fmaf_rn(g, 0.999f, 0.001f); }                                                           // one FFMA
        if (MODE == 4) { g = __fmaf_rn(g, 0.999f, 0.001f); g = fminf(g, 1e6f); }                                       // FFMA + FMNMX
        if (MODE == 5) { g = __fmaf_rn(g, 0.999f, 0.001f); g = g > 0.5f ? g : 0.25f; }                                  // FFMA + FSETP/FSEL
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = acc + g + y2p; }
}

int main()
{
    float *o; long long *c; cudaMalloc(&o, 4); cudaMalloc(&c, 8);
    const float alpha = 0.01f; const double c1 = 1.0 - (double)alpha;
    AgcFast kk{alpha, (float)c1, (float)(c1 - (double)(float)c1), (float)(-0.5 * alpha * 0.6931471805599453), -0.5f * alpha, 1.0f};
    const int n = 100000; long long h;
    const char *names[] = {"agc_step_fast<false>", "agc_step_fast<true>", "MUFU.LG2 + FADD", "FFMA", "FFMA + FMNMX", "FFMA + FSETP + FSEL"};
    for (int m = 0; m < 6; m++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
            case 0: k<0><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break; case 1: k<1><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break;
            case 2: k<2><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break; case 3: k<3><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break;
            case 4: k<4><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break; default: k<5><<<1, 32>>>(o, c, n, kk, 0.3f, 0.2f); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-24s %.1f cycles per step\n", names[m], (double)h / n);
    }
    return 0;
}
