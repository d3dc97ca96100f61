// bam.cu -- BroadcastAM: the reference author's AM broadcast demodulator (demod.hpp:94-153).
//
// Per sample (demod_one, demod.hpp:133-152): x0 = lowpass(x) (Kaiser, 2m+1 taps), x1 = x delayed by m;
// both are mixed down by the carrier oscillator; the PLL is stepped with arg(v0); the output is real(v1)
// through a third-order Chebyshev-II high-pass (two iirfilt_rrrf sections).  The PLL is a nonlinear feedback
// loop, so one thread owns one channel and walks it in time.  The lowpass is feed-forward: samples are taken
// G = 8 at a time and the 8 dot products share one register window (8 shared-memory loads and 8 tap loads per
// 64 FFMA2), each still summing oldest sample first into its own accumulator -- the order of liquid's dotprod.
// The window is a linear per-thread column in shared memory, slid every NG groups; between calls the newest
// ntaps-1 inputs are parked in HBM.  Input may arrive time-major from the decimating kernel (coalesced).
#include <cuda_runtime.h>
#include <math.h>
#include "params.h"
#include "devmath.cuh"
#include "bam.h"

namespace lqb {
namespace {

constexpr int BT = 64;
constexpr int G  = 8;                // samples per register-blocked group
constexpr int NG = 4;                // groups between window slides

__global__ void __launch_bounds__(BT) bam_kernel(const __grid_constant__ BamArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int NTP = a.p.ntaps_pad, H = NTP - 1, W = H + NG * G;
    float2 *s_w   = (float2 *)smem;                       // [W][BT]
    float2 *s_h   = s_w + (size_t)W * BT;                 // [NTP] taps, duplicated for both lanes
    float  *s_sin = (float *)(s_h + NTP);                 // [1024]
    double *s_at  = (double *)(s_sin + 1024);             // [65][kAtanPitch] atan2_rn table

    const int tid = threadIdx.x;
    const long long chl = (long long)blockIdx.x * BT + tid;
    const bool active = chl < a.C;
    const long long cl = active ? chl : 0;                // inactive lanes shadow channel 0 and store nothing
    const long long gch = a.ch0 + cl, CT = a.Ctot, N = a.n;

    for (int i = tid; i < 1024; i += BT) s_sin[i] = a.p.sincos[i].x;
    for (int i = tid; i < 65 * 8; i += BT) s_at[(i >> 3) * kAtanPitch + (i & 7)] = a.p.atantab[i];
    for (int i = tid; i < NTP; i += BT) { const float h = a.p.hrev[i]; s_h[i] = make_float2(h, h); }
    float2 *lp = s_w + tid;
    for (int i = 0; i < H; i++) lp[i * BT] = a.p.hist[i * CT + gch];
    uint32_t theta = a.p.theta[gch], dtheta = a.p.dtheta[gch];
    float v1[2], v2[2];
#pragma unroll
    for (int s = 0; s < 2; s++) { v1[s] = a.p.dcv[(2 * s) * CT + gch]; v2[s] = a.p.dcv[(2 * s + 1) * CT + gch]; }
    float de_v1 = a.has_de ? a.de.v1[gch] : 0.f;
    __syncthreads();                                      // tables; the only barrier

    auto load_x = [&](long long k) -> float2 { return a.in_tmajor ? a.x[k * a.in_pitch + cl] : a.x[cl * a.in_pitch + k]; };
    float *yrow = a.y + cl * a.out_pitch;
    const int dly = H - a.p.m;                            // window slot of x[k - m] relative to the group's output index

    float2 zn[G];
#pragma unroll
    for (int g = 0; g < G; g++) zn[g] = g < N ? load_x(g) : make_float2(0.f, 0.f);
    for (long long kk = 0; kk < N; kk += NG * G) {
        const int consumed = (int)((N - kk) < NG * G ? (N - kk) : NG * G);
#pragma unroll 1
        for (int gi = 0; gi < NG; gi++) {
            const int ng = consumed - gi * G < G ? consumed - gi * G : G;
            if (ng <= 0) break;
            const long long k0 = kk + gi * G;
            float2 *lw = lp + gi * G * BT;                // output g of this group sees lw[g .. g + H]
#pragma unroll
            for (int g = 0; g < G; g++) {
                lw[(H + g) * BT] = zn[g];
                zn[g] = k0 + G + g < N ? load_x(k0 + G + g) : make_float2(0.f, 0.f);
            }
            // lowpass: acc[g] = sum_i hrev[i] * lw[g + i], i ascending
            u64 acc[G], R[2 * G - 1];
#pragma unroll
            for (int g = 0; g < G; g++) acc[g] = 0ull;
#pragma unroll
            for (int j = 0; j < G - 1; j++) R[j] = pk(lw[j * BT]);
#pragma unroll 1
            for (int c = 0; c < NTP; c += G) {
#pragma unroll
                for (int j = 0; j < G; j++) R[G - 1 + j] = pk(lw[(c + G - 1 + j) * BT]);
#pragma unroll
                for (int t = 0; t < G; t++) {
                    const u64 tap = pk(s_h[c + t]);
#pragma unroll
                    for (int g = 0; g < G; g++) acc[g] = fma2(tap, R[t + g], acc[g]);
                }
#pragma unroll
                for (int j = 0; j < G - 1; j++) R[j] = R[j + G];
            }
            float out[G];
#pragma unroll
            for (int g = 0; g < G; g++) {
                out[g] = 0.f;
                if (g < ng) {
                    const unsigned idx = nco_index(theta);
                    const float2 sc = make_float2(s_sin[idx], s_sin[(idx + 256) & 0x3ffu]);
                    const float2 v0 = mix_down(upk(acc[g]), sc), w1 = mix_down(lw[(dly + g) * BT], sc);
                    // arg(v0), taken correctly rounded (the last bit of atan2f differs between math libraries)
                    const float pe = atan2_rn(v0.y, v0.x, s_at);
                    dtheta += nco_constrain_dev(__fmul_rn(pe, a.p.pll_alpha));
                    theta  += nco_constrain_dev(__fmul_rn(pe, a.p.pll_beta));
                    theta  += dtheta;
                    float r = w1.x;
#pragma unroll
                    for (int s = 0; s < 2; s++) {         // iirfiltsos_execute_df2 on real samples
                        const float t  = __fmaf_rn(-a.p.a[s][1], v1[s], r);
                        const float u0 = __fmaf_rn(-a.p.a[s][2], v2[s], t);
                        float y = __fmul_rn(a.p.b[s][1], v1[s]);
                        y = __fmaf_rn(a.p.b[s][0], u0, y);
                        y = __fmaf_rn(a.p.b[s][2], v2[s], y);
                        v2[s] = v1[s]; v1[s] = u0; r = y;
                    }
                    if (a.has_de) { de_v1 = __fmaf_rn(-a.de.a1, de_v1, r); r = __fmul_rn(a.de.b0, de_v1); }
                    out[g] = r;
                }
            }
            if (active) {
                float *yo = yrow + k0;
                if (ng == G && ((((size_t)yo) & 15) == 0)) {
                    *(float4 *)yo = make_float4(out[0], out[1], out[2], out[3]);
                    *(float4 *)(yo + 4) = make_float4(out[4], out[5], out[6], out[7]);
                } else {
#pragma unroll
                    for (int g = 0; g < G; g++) if (g < ng) yo[g] = out[g];
                }
            }
        }
        // slide the window by the samples consumed (ascending: every source lies above its destination)
        const float2 *ls = lp + consumed * BT;
#pragma unroll 4
        for (int i = 0; i < H; i++) lp[i * BT] = ls[i * BT];
    }

    if (active) {
        for (int i = 0; i < H; i++) a.p.hist[i * CT + gch] = lp[i * BT];
        a.p.theta[gch] = theta; a.p.dtheta[gch] = dtheta;
#pragma unroll
        for (int s = 0; s < 2; s++) { a.p.dcv[(2 * s) * CT + gch] = v1[s]; a.p.dcv[(2 * s + 1) * CT + gch] = v2[s]; }
        if (a.has_de) a.de.v1[gch] = de_v1;
    }
}

}  // namespace

cudaError_t bam_launch(const BamArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const int NTP = a.p.ntaps_pad, W = NTP - 1 + NG * G;
    if (NTP % G || a.p.m < 1 || a.p.m > kBamMaxM || a.p.m > NTP - 1) return cudaErrorInvalidValue;
    const size_t smem = (size_t)W * BT * sizeof(float2) + (size_t)NTP * sizeof(float2) + 1024 * sizeof(float) + 65 * kAtanPitch * sizeof(double);
    cudaError_t rc = cudaFuncSetAttribute((const void *)bam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    bam_kernel<<<(unsigned)((a.C + BT - 1) / BT), BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
