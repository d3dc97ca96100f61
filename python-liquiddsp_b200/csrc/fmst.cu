// fmst.cu -- FMStereo (reference demod.hpp:4-85): per sample, frequency discriminator -> mix down by the pilot
// oscillator -> one-pole phase-error filter -> mix down again -> PLL step -> de-emphasis of s + re(sc) and
// s - re(sc) -> two resamp_rrrf with a common phase; one (left, right) pair leaves per resampler output.
//
// The PLL closes a nonlinear loop around every sample, so one thread owns one channel and walks it in time.
// Input tiles [64 channels x 16 samples] are staged through shared memory with coalesced 16-byte loads (each warp
// stages its own 32 rows, double-buffered with cp.async); the two resampler windows are per-thread columns in shared
// memory indexed by the absolute sample count, and the polyphase bank sits in shared memory.  Which samples carry
// an output, and which sub-filter it uses, is the same for every channel (liquid's 8.24 fixed-point phase).
#include <cuda_runtime.h>
#include <math.h>
#include "params.h"
#include "devmath.cuh"
#include "fmst.h"

namespace lqb {
namespace {

constexpr int BT = 64, TS = 16, PITCH = TS * 8 + 16;     // 144-byte rows: conflict-free per-lane LDS.128

template <bool UP>
__global__ void __launch_bounds__(BT) fmstereo_kernel(const __grid_constant__ FmstArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = a.p.rs.sublen;
    unsigned char *s_in = smem;                                   // 2 stages of [BT][PITCH]
    float *s_L   = (float *)(s_in + 2 * BT * PITCH);              // [L][BT]
    float *s_R   = s_L + L * BT;
    float *s_b   = s_R + L * BT;                                  // bank [npfb][L]
    float *s_sin = s_b + ((a.p.rs.npfb * L + 3) & ~3);            // [1024]
    double *s_at = (double *)(s_sin + 1024);                      // [65][kAtanPitch] atan2_rn table

    // cpw channels per warp: the loop is one long FP64 dependency chain per sample, so with few channels they are
    // spread over more warps (8 or 16 working lanes each) and the schedulers get independent chains to interleave
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31, cpw = a.cpw, RC = (BT / 32) * cpw;
    const bool worker = lane < cpw;
    const int myrow = wid * cpw + (worker ? lane : 0);
    const long long chl = (long long)blockIdx.x * RC + myrow;
    const bool active = worker && chl < a.C;
    const long long gch = a.ch0 + (active ? chl : 0), CT = a.Ctot, N = a.n;

    for (int i = tid; i < 1024; i += BT) s_sin[i] = a.p.sincos[i].x;
    for (int i = tid; i < 65 * 8; i += BT) s_at[(i >> 3) * kAtanPitch + (i & 7)] = a.p.atantab[i];
    for (int i = tid; i < a.p.rs.npfb * L; i += BT) s_b[i] = a.p.rs.bank[i];
    if (worker) for (int i = 0; i < L; i++) { s_L[i * BT + myrow] = a.p.ringL[i * CT + gch]; s_R[i * BT + myrow] = a.p.ringR[i * CT + gch]; }
    float2 prev = a.p.rprime[gch];
    uint32_t theta = a.p.theta[gch], dtheta = a.p.dtheta[gch];
    float pe = a.p.pe[gch], vL = a.p.vL[gch], vR = a.p.vR[gch];
    uint32_t phase = a.p.rs.phase;
    int slot = (int)(a.p.rs.count % (unsigned)L);                 // where the next sample goes (= oldest sample)
    long long kout = 0;
    float *yrow = a.y + (active ? chl : 0) * 2 * a.n_out;

    // a warp stages its own cpw rows: 8 lanes cover one 128-byte row segment, 4 rows per pass
    const long long row0 = (long long)blockIdx.x * RC + wid * cpw;
    const bool vec = ((N & 1) == 0) && ((((size_t)a.x) & 15) == 0);
    auto load_tile = [&](long long t, int stage) {
        unsigned char *dst = s_in + stage * (BT * PITCH) + (wid * cpw) * PITCH;
        const long long n0 = t * TS;
#pragma unroll
        for (int ps = 0; ps < 8; ps++) {
            const int r = ps * 4 + (lane >> 3), c = lane & 7;     // row within the warp, 16-byte chunk within the row
            if (r >= cpw) break;
            const long long ch = row0 + r, s0 = n0 + 2 * c;
            unsigned char *d = dst + r * PITCH + c * 16;
            if (ch < a.C && vec && s0 + 1 < N) cp_async16(d, a.x + ch * N + s0, 16);
            else {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ch < a.C) {
                    if (s0 < N)     { const float2 u = a.x[ch * N + s0];     v.x = u.x; v.y = u.y; }
                    if (s0 + 1 < N) { const float2 u = a.x[ch * N + s0 + 1]; v.z = u.x; v.w = u.y; }
                }
                *(float4 *)d = v;
            }
        }
    };
    __syncthreads();

    const long long ntiles = (N + TS - 1) / TS;
    load_tile(0, 0); cp_async_commit();
    for (long long t = 0; t < ntiles; t++) {
        const int stage = (int)(t & 1);
        if (t + 1 < ntiles) load_tile(t + 1, stage ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        unsigned char *row = s_in + stage * (BT * PITCH) + myrow * PITCH;
        const int nv = worker ? (int)((N - t * TS) < TS ? (N - t * TS) : TS) : 0;
        if (worker) {
        // the discriminator has no feedback: its 16 arguments are independent work, done ahead of the PLL loop (the
        // result replaces the sample in the thread's own staged row)
#pragma unroll 4
        for (int j = 0; j < TS; j++) {
            const float2 z = *(const float2 *)(row + j * 8);
            // freqdem_demodulate: arg(conj(r') r) / (2 pi kf), the argument taken correctly rounded
            const float re = __fmaf_rn(prev.x, z.x, __fmul_rn(prev.y, z.y));
            const float im = __fmaf_rn(prev.x, z.y, -__fmul_rn(prev.y, z.x));
            *(float *)(row + j * 8) = __fmul_rn(atan2_rn(im, re, s_at), a.p.ref);
            if (j < nv) prev = z;
        }
        }
#pragma unroll 1
        for (int j = 0; j < nv; j++) {
            const float s = *(const float *)(row + j * 8);
            const unsigned idx = nco_index(theta);
            const float2 osc = make_float2(s_sin[idx], s_sin[(idx + 256) & 0x3ffu]);
            float2 sc = mix_down(make_float2(s, 0.f), osc);
            const float arg = atan2_rn(sc.y, sc.x, s_at);
            pe = (float)fma(0.999, (double)pe, 0.001 * (double)arg);
            sc = mix_down(sc, osc);
            dtheta += nco_constrain_dev(__fmul_rn(pe, a.p.pll_alpha));
            theta  += nco_constrain_dev(__fmul_rn(pe, a.p.pll_beta));
            theta  += dtheta;
            // de-emphasis (iirfilt_rrrf, b = [b0], a = [1, a1]) of the sum and the difference
            vL = __fmaf_rn(-a.p.a1, vL, __fadd_rn(s, sc.x)); const float left  = __fmaf_rn(a.p.b0, vL, 0.f);
            vR = __fmaf_rn(-a.p.a1, vR, __fsub_rn(s, sc.x)); const float right = __fmaf_rn(a.p.b0, vR, 0.f);
            // resamp_rrrf_execute on both: push, then one output for every phase position inside this sample
            auto dot = [&](const float *win, uint32_t ph) -> float {
                const float *h = s_b + (ph >> (24 - a.p.rs.bits)) * L;
                float acc = 0.f;
                int q = slot;                                      // oldest sample first (slot already points past the newest)
                for (int i = 0; i < L; i++) { acc = __fmaf_rn(h[i], win[q * BT + myrow], acc); q = q + 1 == L ? 0 : q + 1; }
                return acc;
            };
            const int here = slot;
            slot = slot + 1 == L ? 0 : slot + 1;
            s_L[here * BT + myrow] = left;
            if constexpr (!UP) {
                s_R[here * BT + myrow] = right;
                if (phase <= 0x00ffffffu) {                        // rate <= 1: at most one output per input
                    const float aL = dot(s_L, phase), aR = dot(s_R, phase);
                    if (active) *(float2 *)(yrow + 2 * kout) = make_float2(aL, aR);
                    kout++;
                    phase += a.p.rs.step;
                }
            } else {
                // pcm_rate > iq_rate: a sample can yield several outputs.  demod_one (demod.hpp:79-83) passes &y[nw] and
                // &y[nw + 1] to the resamplers as input AND output, so the left resampler's second output overwrites the right
                // resampler's input before it is read, and execute() (:45-48) keeps a pair only when each yields exactly one.
                int nl = 0; uint32_t ph = phase;
                while (ph <= 0x00ffffffu) { nl++; ph += a.p.rs.step; }             // (the same for every channel)
                s_R[here * BT + myrow] = nl >= 2 ? dot(s_L, phase + a.p.rs.step) : right;
                if (nl == 1) {
                    const float aL = dot(s_L, phase), aR = dot(s_R, phase);
                    if (active) *(float2 *)(yrow + 2 * kout) = make_float2(aL, aR);
                    kout++;
                }
                phase = ph;
            }
            phase -= (1u << 24);
        }
        __syncwarp();
    }

    if (active) {
        for (int i = 0; i < L; i++) { a.p.ringL[i * CT + gch] = s_L[i * BT + myrow]; a.p.ringR[i * CT + gch] = s_R[i * BT + myrow]; }
        a.p.rprime[gch] = prev; a.p.theta[gch] = theta; a.p.dtheta[gch] = dtheta;
        a.p.pe[gch] = pe; a.p.vL[gch] = vL; a.p.vR[gch] = vR;
    }
}

}  // namespace

cudaError_t fmstereo_launch(const FmstArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const int L = a.p.rs.sublen;
    if (L < 1 || L > kFmstMaxSub || a.p.rs.step == 0) return cudaErrorInvalidValue;
    const size_t smem = (size_t)2 * BT * PITCH + (size_t)2 * L * BT * sizeof(float) + (size_t)((a.p.rs.npfb * L + 3) & ~3) * sizeof(float)
                      + 1024 * sizeof(float) + 65 * kAtanPitch * sizeof(double);
    void (*fn)(const FmstArgs) = a.p.rs.step < (1u << 24) ? fmstereo_kernel<true> : fmstereo_kernel<false>;
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    if (a.cpw != 8 && a.cpw != 16 && a.cpw != 32) return cudaErrorInvalidValue;
    const int rc_cta = (BT / 32) * a.cpw;
    fn<<<(unsigned)((a.C + rc_cta - 1) / rc_cta), BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
