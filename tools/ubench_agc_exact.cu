// Latency of the bit-exact gain loop's dependent chain (am.cu agc_tmajor_kernel<false>) and of its pieces on sm_100a:
// one warp, cycles per step.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_agc_exact ubench_agc_exact.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../python-liquiddsp_b200/csrc/devmath.cuh"
using namespace lqb;

template <int MODE>
__global__ void k(float *out, long long *cyc, int n, const double2 *logtab, float alpha, double oma, float xr, float xi)
{
    __shared__ double2 s_log[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) s_log[i] = logtab[i];
    __syncthreads();
    float g = 1.f, y2p = 1.f, acc = 0.f; double d = 1.0 + 1e-9 * threadIdx.x; int ii = threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < n; i++) {
        if (MODE == 0) {              // the whole exact step
            const float yr = __fmul_rn(xr, g), yi = __fmul_rn(xi, g);
            const float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
            y2p = (float)fma(oma, (double)y2p, (double)__fmul_rn(alpha, y2));
            const float ge = __fmul_rn(g, exp_rn_warp(__fmul_rn(__fmul_rn(-0.5f, alpha), log_rn(fmaxf(y2p, 1e-30f), s_log))));
            g = y2p > 1e-6f ? ge : g; g = g > 1e6f ? 1e6f : g; acc += yr;
        }
        if (MODE == 1) { d = fma(d, 0.999999, 1e-7); }                                         // DFMA
        if (MODE == 2) { g = (float)((double)g * 1.0000001); }                                  // F2F.F64.F32 + DMUL + F2F.F32.F64
        if (MODE == 3) { g = log_rn(fmaxf(g, 1e-30f), s_log) + 2.0f; }                         // log_rn + FADD
        if (MODE == 4) { g = exp_rn_warp(__fmul_rn(g, 1e-3f)); }                                // FMUL + exp_rn_warp
        if (MODE == 5) { y2p = (float)fma(oma, (double)y2p, (double)__fmul_rn(alpha, g)); g = y2p; }   // the smoothing update
        if (MODE == 6) { ii = (int)s_log[ii & 127].x + ii; }                                    // LDS.64 + F2I + IADD
        if (MODE == 7) { d = d * d; d = d + 0.5; }                                              // DMUL + DADD
        if (MODE == 8) { g = exp_rn_small(__fmul_rn(g, 1e-3f)); }                               // FMUL + exp_rn_small (no vote)
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = acc + g + y2p + (float)d + ii; }
}

int main()
{
    float *o; long long *c; cudaMalloc(&o, 4); cudaMalloc(&c, 8);
    std::vector<double> t(256);
    for (int i = 0; i < 128; i++) { const double ci = 1.0 + (i + 0.5) / 128.0; const double inv = (double)(float)(1.0 / ci); t[2 * i] = inv; t[2 * i + 1] = -std::log(inv); }
    double2 *dt; cudaMalloc(&dt, 128 * sizeof(double2)); cudaMemcpy(dt, t.data(), 128 * sizeof(double2), cudaMemcpyHostToDevice);
    const float alpha = 0.01f; const double oma = 1.0 - (double)alpha;
    const int n = 50000; long long h;
    const char *names[] = {"exact gain step", "DFMA", "F2F.F64 + DMUL + F2F.F32", "log_rn + FADD", "FMUL + exp_rn_warp", "smoothing (cvt, cvt, DFMA, cvt)", "LDS.64 + F2I + IADD", "DMUL + DADD", "FMUL + exp_rn_small"};
    for (int m = 0; m < 9; m++) {
        for (int rep = 0; rep < 2; rep++) {
            switch (m) {
            case 0: k<0><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break; case 1: k<1><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break;
            case 2: k<2><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break; case 3: k<3><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break;
            case 4: k<4><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break; case 5: k<5><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break;
            case 6: k<6><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break; case 7: k<7><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break;
            default: k<8><<<1, 32>>>(o, c, n, dt, alpha, oma, 0.7f, 0.6f); break;
            }
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-36s %.1f cycles per step (%s)\n", names[m], (double)h / n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
