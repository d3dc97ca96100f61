// am.cu -- the decimated-rate tail of the AM receiver: [AGC ->] ampmodem (DSB) [-> de-emphasis].
//
// Stands in for the reference's per-sample loops AGC::execute (agc.hpp:109-128),
// ampmodem_demodulate_block (demod.hpp:294) and DeemphasisFilter::execute (iirfilter.hpp:384-391).
// The AGC and the carrier PLL are nonlinear feedback loops, so one thread owns one channel and walks
// it in time; what can be restructured without touching the arithmetic is everything around them:
//   * samples are taken G = 8 at a time: AGC x8, lowpass FIR x8, PLL x8, DC-block FIR x8, de-emphasis x8.
//     The two 51-tap FIRs are feed-forward, so the 8 outputs of a group share every window load
//     (58 loads feed 8 x 51 multiply-adds) and give the scheduler 8 independent accumulation chains.
//     Each output still sums its taps oldest-first into one accumulator: bit-identical to the
//     sample-by-sample order;
//   * the windows are linear per-thread columns in shared memory (no ring index arithmetic); they slide
//     after every two groups.  A thread touches only its own column, so the sample loop has no barrier;
//   * input arrives time-major [sample][channel] from the full-rate kernel (coalesced 8-byte loads, no
//     staging) or row-major from a caller; output goes out 8 floats (one 32-byte sector) per thread.
// Between calls the windows live in the 64-slot rings of params.h, indexed by absolute sample count.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>
#include "params.h"
#include "devmath.cuh"
#include "am.h"
#include "bam.h"

namespace lqb {
namespace {

constexpr int BT = kAmBT;
constexpr int G  = 8;                // samples per register-blocked group
constexpr int NG = 2;                // groups between window slides
constexpr int H  = kAmTaps - 1;      // history samples a window needs: 50
constexpr int W  = H + NG * G;       // window length: 66
constexpr int DOFF = 2;              // amtail_kernel: the DC window starts two slots into its row, so that the slot a group's results go to
                                     // (H + DOFF = 52) is 16-byte aligned
constexpr int WDP = W + DOFF;        // DC row length: 68 floats (68 mod 32 = 4: a quarter warp's 16-byte accesses cover all 32 banks, as
                                     // do the lowpass rows of 66 float2 = 132 words)

// OUT_V1 (ampmodem USB / LSB with carrier): the kernel stops after the carrier loop and writes the mixed-down delayed
// branch v1 as complex samples; the Hilbert pair and the DC blocker are feed-forward and follow as FIR launches
template <bool HAS_AGC, bool HAS_DE, bool OUT_V1>
__global__ void __launch_bounds__(BT, 4) amtail_kernel(const __grid_constant__ AmTailArgs a)
{
    extern __shared__ __align__(16) unsigned char smem[];
    float2 *s_lp  = (float2 *)smem;                       // [BT][W]   one row per thread: 16-byte loads, stores and slides
    float  *s_dc  = (float *)(s_lp + W * BT);             // [BT][WDP]
    float  *s_sin = s_dc + WDP * BT;                      // [1024]
    double2 *s_log = (double2 *)(s_sin + 1024);           // [128] (only with the AGC)

    const int tid = threadIdx.x;
    const long long chl = (long long)blockIdx.x * BT + tid;
    const bool active = chl < a.C;
    const long long cl = active ? chl : 0;                // inactive lanes shadow channel 0 and store nothing
    const long long gch = a.ch0 + cl;
    const long long CT = a.Ctot, N = a.n;

    for (int i = tid; i < 1024; i += BT) s_sin[i] = a.am.sincos[i].x;
    if (HAS_AGC) for (int i = tid; i < 128; i += BT) s_log[i] = a.agc.logtab[i];
    float agc_g = 1.f, agc_y2p = 1.f; int agc_mode = 7; unsigned agc_timer = 0, agc_rises = 0;
    if (HAS_AGC) { agc_g = a.agc.g[gch]; agc_y2p = a.agc.y2p[gch]; agc_mode = a.agc.mode[gch]; agc_timer = a.agc.timer[gch]; }
    uint32_t theta = a.am.theta[gch], dtheta = a.am.dtheta[gch];
    float de_v1 = HAS_DE ? a.de.v1[gch] : 0.f;
    // history: the H samples before this call, oldest first, from the rings
    float2 *lp = s_lp + tid * W; float *dc = s_dc + tid * WDP;
    if (!a.am.suppressed) {
        for (int i = 0; i < H; i++) {
            const unsigned slot = (a.am.count + (unsigned)(kAmRing - H) + i) & (kAmRing - 1);
            lp[i] = a.am.lp_ring[slot * CT + gch];
            dc[DOFF + i] = a.am.dc_ring[slot * CT + gch];
        }
    }
    __syncthreads();                                      // the sine table; the only barrier

    auto load_x = [&](long long k) -> float2 {
        return a.in_tmajor ? a.x[k * a.in_pitch + cl] : a.x[cl * a.in_pitch + k];
    };
    const AgcFast agck{a.agc.alpha, a.agc.chi, a.agc.clo, a.agc.cl2, a.agc.chalf, a.agc.scale};
    auto agc_step = [&](float2 z) -> float2 {
        if (HAS_AGC && a.agc.fast) return a.agc.big ? agc_step_fast<true>(z, agc_g, agc_y2p, agck) : agc_step_fast<false>(z, agc_g, agc_y2p, agck);
        // agc_crcf_execute (liquid agc.proto.c) then the wrapper's status poll, agc.hpp:115-125
        float yr = __fmul_rn(z.x, agc_g), yi = __fmul_rn(z.y, agc_g);
        const float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
        agc_y2p = (float)fma(a.agc.one_minus_alpha, (double)agc_y2p, (double)__fmul_rn(a.agc.alpha, y2));
        if (!a.agc.locked) {
            if (agc_y2p > 1e-6f) agc_g = __fmul_rn(agc_g, exp_rn_small(__fmul_rn(__fmul_rn(-0.5f, a.agc.alpha), log_rn(agc_y2p, s_log))));
            if (agc_g > 1e6f) agc_g = 1e6f;
            if (agc_mode != 7) {
                const bool ex = (float)(-20.0 * log10((double)agc_g)) > a.agc.threshold;
                const int before = agc_mode;
                switch (agc_mode) {
                case 1: agc_mode = ex ? 2 : 1; break;
                case 2: agc_mode = ex ? 3 : 4; break;
                case 3: agc_mode = ex ? 3 : 4; break;
                case 4: agc_timer = a.agc.timeout; agc_mode = ex ? 3 : 5; break;
                case 5: agc_timer--; if (agc_timer == 0) agc_mode = 6; else if (ex) agc_mode = 3; break;
                case 6: agc_mode = 1; break;
                default: break;
                }
                if (agc_mode == 2 && before != 2) agc_rises++;
            }
            yr = __fmul_rn(yr, a.agc.scale); yi = __fmul_rn(yi, a.agc.scale);
        }
        if (agc_mode == 5 || agc_mode == 1) { yr = __fmul_rn(yr, 0.0f); yi = __fmul_rn(yi, 0.0f); }
        return make_float2(yr, yi);
    };
    auto nco_sc = [&]() -> float2 {
        const unsigned idx = nco_index(theta);
        return make_float2(s_sin[idx], s_sin[(idx + 256) & 0x3ffu]);
    };
    auto pll = [&](float pe) {
        dtheta += nco_constrain_dev(__fmul_rn(pe, a.am.pll_alpha));
        theta  += nco_constrain_dev(__fmul_rn(pe, a.am.pll_beta));
        theta  += dtheta;
    };
    auto deemph = [&](float r) -> float {
        if (!HAS_DE) return r;
        de_v1 = __fmaf_rn(-a.de.a1, de_v1, r);
        return __fmul_rn(a.de.b0, de_v1);
    };
    float *yrow = a.y + cl * a.out_pitch;

    if (a.am.suppressed) {
        // ampmodem_demod_dsb_pll_costas: no filters, one sample at a time
        for (long long k = 0; k < N; k++) {
            float2 z = load_x(k);
            if (HAS_AGC) z = agc_step(z);
            const float2 v = mix_down(z, nco_sc());
            pll(__fmul_rn(v.y, v.x > 0.f ? 1.f : -1.f));
            const float r = deemph(__fdiv_rn(v.x, a.am.mod_index));
            if (active) yrow[k] = r;
        }
    } else {
        // ampmodem_demod_dsb_pll_carrier, G samples per pass, NG passes per window slide
        float2 zn[G];                           // next group's input, fetched one group ahead
#pragma unroll
        for (int g = 0; g < G; g++) zn[g] = g < N ? load_x(g) : make_float2(0.f, 0.f);
        for (long long kk = 0; kk < N; kk += NG * G) {
            const int consumed = (int)((N - kk) < NG * G ? (N - kk) : NG * G);
#pragma unroll 1
            for (int gi = 0; gi < NG; gi++) {
                const int ng = consumed - gi * G < G ? consumed - gi * G : G;
                if (ng <= 0) break;
                const long long k0 = kk + gi * G;
                float2 *lw = lp + gi * G; float *dw = dc + gi * G;               // this group's window base (16-byte aligned)
                float2 z[G];
#pragma unroll
                for (int g = 0; g < G; g++) { z[g] = zn[g]; zn[g] = k0 + G + g < N ? load_x(k0 + G + g) : make_float2(0.f, 0.f); }
#pragma unroll
                for (int g = 0; g < G; g++) if (HAS_AGC && g < ng) z[g] = agc_step(z[g]);
#pragma unroll
                for (int g = 0; g < G; g += 2) *(float4 *)&lw[H + g] = make_float4(z[g].x, z[g].y, z[g + 1].x, z[g + 1].y);
                // lowpass: x0[g] = sum_i lp[i] * window[g + i], i ascending (oldest sample first)
                u64 s2[G];                      // (re, im) accumulators: one FFMA2 per tap and output
#pragma unroll
                for (int g = 0; g < G; g++) s2[g] = 0ull;
#pragma unroll
                for (int i0 = 0; i0 < H + G; i0 += 2) {
                    const float4 q = *(const float4 *)&lw[i0];                 // window samples i0 and i0 + 1
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int i = i0 + h;
                        const u64 w = h ? pk(q.z, q.w) : pk(q.x, q.y);
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            const int t = i - g;
                            if (t >= 0 && t < kAmTaps) s2[g] = fma2(pk(a.am.lp[t], a.am.lp[t]), w, s2[g]);
                        }
                    }
                }
                // carrier PLL on the filtered branch, mix the delayed branch with the same phase
                float m[G];
#pragma unroll
                for (int g = 0; g < G; g++) {
                    m[g] = 0.f;
                    if (g < ng) {
                        const float2 sc = nco_sc();
                        const float2 x1 = lw[H + g - kAmDelay];
                        const float2 v0 = mix_down(upk(s2[g]), sc), v1 = mix_down(x1, sc);
                        pll(v0.y);
                        if (OUT_V1) { if (active) ((float2 *)a.y)[cl * a.out_pitch + k0 + g] = v1; }
                        else m[g] = __fdiv_rn(v1.x, a.am.mod_index);
                    }
                }
                *(float4 *)&dw[H + 2] = make_float4(m[0], m[1], m[2], m[3]);   // (H + 2 = 52: 16-byte aligned; the window starts two slots in)
                *(float4 *)&dw[H + 6] = make_float4(m[4], m[5], m[6], m[7]);
                if constexpr (!OUT_V1) {
                // dc blocker, same blocking
                float acc[G];
#pragma unroll
                for (int g = 0; g < G; g++) acc[g] = 0.f;
#pragma unroll
                for (int i0 = 0; i0 < H + G + 2; i0 += 4) {                       // window slots 2 .. H + G + 1 of the row (DOFF = 2)
                    const float4 q = *(const float4 *)&dw[i0];
                    const float wv[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
                    for (int h = 0; h < 4; h++) {
                        const int i = i0 + h - DOFF;
                        if (i < 0 || i >= H + G) continue;
#pragma unroll
                        for (int g = 0; g < G; g++) {
                            const int t = i - g;
                            if (t >= 0 && t < kAmTaps) acc[g] = __fmaf_rn(a.am.dc[t], wv[h], acc[g]);
                        }
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++) if (g < ng) acc[g] = deemph(acc[g]);
                if (active) {
                    float *yo = yrow + k0;
                    if (ng == G && ((((size_t)yo) & 15) == 0)) {
                        *(float4 *)yo = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        *(float4 *)(yo + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    } else {
#pragma unroll
                        for (int g = 0; g < G; g++) if (g < ng) yo[g] = acc[g];
                    }
                }
                }   // !OUT_V1
            }
            // slide both windows by the samples consumed
            // (loads in batches ahead of the stores: the source lies above everything a batch writes)
            if (consumed == NG * G) {
                // (rows are 16-byte aligned and a whole frame is a multiple of 16 bytes: the slide moves four floats at a time)
                const float4 *ls = (const float4 *)(lp + NG * G); float4 *ld = (float4 *)lp;
                const float4 *ds = (const float4 *)(dc + NG * G); float4 *dd = (float4 *)dc;
                constexpr int NL = H / 2, ND = (DOFF + H + 3) / 4;                   // 25 and 13 vectors
                float4 tl[NL], td[ND];
#pragma unroll
                for (int i = 0; i < NL; i++) tl[i] = ls[i];
#pragma unroll
                for (int i = 0; i < ND; i++) td[i] = ds[i];
#pragma unroll
                for (int i = 0; i < NL; i++) ld[i] = tl[i];
#pragma unroll
                for (int i = 0; i < ND; i++) dd[i] = td[i];
            } else {
                // the call's last, shorter frame: one entry at a time (ascending: the source lies above the destination)
                for (int i = 0; i < H; i++) { lp[i] = lp[consumed + i]; dc[DOFF + i] = dc[DOFF + consumed + i]; }
            }
        }
    }

    if (active) {
        if (HAS_AGC) {
            a.agc.g[gch] = agc_g; a.agc.y2p[gch] = agc_y2p; a.agc.mode[gch] = agc_mode; a.agc.timer[gch] = agc_timer;
            if (agc_rises) atomicAdd(a.agc.rise_count, agc_rises);
        }
        a.am.theta[gch] = theta; a.am.dtheta[gch] = dtheta;
        if (HAS_DE) a.de.v1[gch] = de_v1;
        if (!a.am.suppressed) {
            // the windows now hold the last H samples; park them in the rings by absolute index
            const unsigned base = a.am.count + (unsigned)(N % kAmRing) + (unsigned)(kAmRing - H);
            for (int i = 0; i < H; i++) {
                const unsigned slot = (base + i) & (kAmRing - 1);
                a.am.lp_ring[slot * CT + gch] = lp[i];
                a.am.dc_ring[slot * CT + gch] = dc[DOFF + i];
            }
        }
    }
}

// AGC alone over a time-major [sample][channel] block, in place.  The gain loop is one long dependent chain per
// sample (double-precision smoothing, log, exp); on its own it needs no shared-memory windows, so all 2048
// threads of an SM can be resident and the chains of ~14 warps hide each other.
// FAST: unlocked, squelch disabled -- the single-precision gain loop (devmath.cuh agc_step_fast).  Its chain is short
// enough that the loads become the limit, so 16 samples per thread are in flight instead of 4.
template <bool FAST, int U, bool BIG>
__global__ void __launch_bounds__(128) agc_tmajor_kernel(const __grid_constant__ AmTailArgs a)
{
    __shared__ double2 s_log[FAST ? 1 : 128];
    const int tid = threadIdx.x;
    const long long chl = (long long)blockIdx.x * blockDim.x + tid;
    if constexpr (!FAST) {
        for (int i = tid; i < 128; i += blockDim.x) s_log[i] = a.agc.logtab[i];
        __syncthreads();
    }
    if (chl >= a.C) return;
    const long long gch = a.ch0 + chl, N = a.n, P = a.in_pitch;
    float agc_g = a.agc.g[gch], agc_y2p = a.agc.y2p[gch]; int agc_mode = a.agc.mode[gch]; unsigned agc_timer = a.agc.timer[gch], agc_rises = 0;
    float2 *x = const_cast<float2 *>(a.x) + chl;
    if constexpr (FAST) {
        const AgcFast k{a.agc.alpha, a.agc.chi, a.agc.clo, a.agc.cl2, a.agc.chalf, a.agc.scale};
        float2 zn[U];
#pragma unroll
        for (int u = 0; u < U; u++) zn[u] = u < N ? x[u * P] : make_float2(0.f, 0.f);
        long long k0 = 0;
        for (; k0 + U <= N; k0 += U) {
            float2 z[U];
#pragma unroll
            for (int u = 0; u < U; u++) { z[u] = zn[u]; zn[u] = k0 + U + u < N ? x[(k0 + U + u) * P] : make_float2(0.f, 0.f); }
#pragma unroll
            for (int u = 0; u < U; u++) x[(k0 + u) * P] = agc_step_fast<BIG>(z[u], agc_g, agc_y2p, k);
        }
        for (int u = 0; k0 + u < N; u++) x[(k0 + u) * P] = agc_step_fast<BIG>(x[(k0 + u) * P], agc_g, agc_y2p, k);
        a.agc.g[gch] = agc_g; a.agc.y2p[gch] = agc_y2p;
    } else {
    float2 zn[U];
#pragma unroll
    for (int u = 0; u < U; u++) zn[u] = u < N ? x[u * P] : make_float2(0.f, 0.f);
    for (long long k0 = 0; k0 < N; k0 += U) {
        float2 z[U];
#pragma unroll
        for (int u = 0; u < U; u++) { z[u] = zn[u]; zn[u] = k0 + U + u < N ? x[(k0 + U + u) * P] : make_float2(0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (k0 + u < N) {
                float yr = __fmul_rn(z[u].x, agc_g), yi = __fmul_rn(z[u].y, agc_g);
                const float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
                agc_y2p = (float)fma(a.agc.one_minus_alpha, (double)agc_y2p, (double)__fmul_rn(a.agc.alpha, y2));
                if (!a.agc.locked) {
                    // the gain update without data-dependent branches: the logarithm takes a harmless argument when the
                    // level is below 1e-6 and its result is then not used
                    const float ge = __fmul_rn(agc_g, exp_rn_warp(__fmul_rn(__fmul_rn(-0.5f, a.agc.alpha), log_rn(fmaxf(agc_y2p, 1e-30f), s_log))));
                    agc_g = agc_y2p > 1e-6f ? ge : agc_g;
                    agc_g = agc_g > 1e6f ? 1e6f : agc_g;
                    if (agc_mode != 7) {
                        const bool ex = (float)(-20.0 * log10((double)agc_g)) > a.agc.threshold;
                        const int before = agc_mode;
                        switch (agc_mode) {
                        case 1: agc_mode = ex ? 2 : 1; break;
                        case 2: agc_mode = ex ? 3 : 4; break;
                        case 3: agc_mode = ex ? 3 : 4; break;
                        case 4: agc_timer = a.agc.timeout; agc_mode = ex ? 3 : 5; break;
                        case 5: agc_timer--; if (agc_timer == 0) agc_mode = 6; else if (ex) agc_mode = 3; break;
                        case 6: agc_mode = 1; break;
                        default: break;
                        }
                        if (agc_mode == 2 && before != 2) agc_rises++;
                    }
                    yr = __fmul_rn(yr, a.agc.scale); yi = __fmul_rn(yi, a.agc.scale);
                }
                if (agc_mode == 5 || agc_mode == 1) { yr = __fmul_rn(yr, 0.0f); yi = __fmul_rn(yi, 0.0f); }
                x[(k0 + u) * P] = make_float2(yr, yi);
            }
        }
    }
    a.agc.g[gch] = agc_g; a.agc.y2p[gch] = agc_y2p; a.agc.mode[gch] = agc_mode; a.agc.timer[gch] = agc_timer;
    if (agc_rises) atomicAdd(a.agc.rise_count, agc_rises);
    }
}

// ---- few channels: eight lanes per channel ----------------------------------------------------------------------
// With one thread per channel a lone warp walks 220 instructions per sample in order, and the two 51-tap filters --
// 80 % of them -- wait behind the carrier loop's dependent chain although they are feed-forward.  Here a channel is
// eight lanes: lane g of the eight computes output g of each group of eight samples -- both filters, each output
// still summing its taps oldest-first into one accumulator (bit-identical) -- while the carrier loop, the only part
// that is sequential in time, runs (redundantly, on identical values) in all eight.  The windows are per-channel rings
// in shared memory, stored twice 64 slots apart so that a 51-sample window is a linear read at any position.
// Input: time-major [n][C] or row-major; AGC is not part of this kernel (it runs in place on the hand-off before it).
constexpr int kA8Lanes = 8, kA8Cpw = 32 / kA8Lanes, kA8Warps = 4, kA8Cpc = kA8Cpw * kA8Warps;   // 4 channels per warp, 16 per CTA
template <bool HAS_DE>
__global__ void __launch_bounds__(32 * kA8Warps) amtail8_kernel(const __grid_constant__ AmTailArgs a)
{
    __shared__ float2 s_lp[kA8Cpc][2 * kAmRing];
    __shared__ float  s_dc[kA8Cpc][2 * kAmRing];
    __shared__ float  s_sin[1024];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int g = lane & (kA8Lanes - 1), cw = wid * kA8Cpw + (lane >> 3);       // lane of its channel, channel within the CTA
    const unsigned gmask = 0xffu << (lane & 24);                                // this channel's eight lanes
    const int glead = lane & 24;
    const long long chl = (long long)blockIdx.x * kA8Cpc + cw;
    const bool active = chl < a.C;
    const long long cl = active ? chl : 0, gch = a.ch0 + cl, CT = a.Ctot, N = a.n;
    for (int i = tid; i < 1024; i += blockDim.x) s_sin[i] = a.am.sincos[i].x;
    float2 *lp = s_lp[cw]; float *dc = s_dc[cw];
    // the rings by absolute sample index: slot (count + k) & 63 holds sample k of this call (k < 0: earlier calls)
    for (int i = g; i < kAmRing; i += kA8Lanes) {
        const float2 v = a.am.lp_ring[i * CT + gch]; const float d = a.am.dc_ring[i * CT + gch];
        lp[i] = v; lp[i + kAmRing] = v; dc[i] = d; dc[i + kAmRing] = d;
    }
    uint32_t theta = a.am.theta[gch], dtheta = a.am.dtheta[gch];
    float de_v1 = HAS_DE ? a.de.v1[gch] : 0.f;
    __syncthreads();                                      // the sine table; windows are per warp from here on

    auto load_x = [&](long long k) -> float2 {
        return k < N ? (a.in_tmajor ? a.x[k * a.in_pitch + cl] : a.x[cl * a.in_pitch + k]) : make_float2(0.f, 0.f);
    };
    float *yrow = a.y + cl * a.out_pitch;
    const unsigned cnt = a.am.count;
    float2 zn = load_x(g);
    for (long long k0 = 0; k0 < N; k0 += kA8Lanes) {
        const int ng = (int)((N - k0) < kA8Lanes ? (N - k0) : kA8Lanes);
        // this group's samples into the ring (both copies), the next group's on their way
        const float2 z = zn;
        zn = load_x(k0 + kA8Lanes + g);
        const unsigned pz = (cnt + (unsigned)k0 + (unsigned)g) & (kAmRing - 1);
        if (g < ng) { lp[pz] = z; lp[pz + kAmRing] = z; }
        __syncwarp();
        // lowpass: output k0 + g = sum_t lp[t] * x[k0 + g - 50 + t], t ascending (oldest sample first)
        const float2 *w = lp + ((cnt + (unsigned)k0 + (unsigned)g + (unsigned)(kAmRing - H)) & (kAmRing - 1));
        u64 s2 = 0ull;
#pragma unroll
        for (int t = 0; t < kAmTaps; t++) s2 = fma2(pk(a.am.lp[t], a.am.lp[t]), pk(w[t]), s2);
        const float2 lpo = upk(s2);
        const float2 x1 = w[H - kAmDelay];                // the delayed branch: sample k0 + g - 25
        // carrier PLL, sample by sample, on every lane of the channel: lane j's filter output and delayed sample by shuffle
        float mine = 0.f;
#pragma unroll
        for (int j = 0; j < kA8Lanes; j++) {
            const float sr = __shfl_sync(0xffffffffu, lpo.x, glead + j), si = __shfl_sync(0xffffffffu, lpo.y, glead + j);
            const float xr = __shfl_sync(0xffffffffu, x1.x, glead + j), xi = __shfl_sync(0xffffffffu, x1.y, glead + j);
            if (j < ng) {
                const unsigned idx = nco_index(theta);
                const float2 sc = make_float2(s_sin[idx], s_sin[(idx + 256) & 0x3ffu]);
                const float2 v0 = mix_down(make_float2(sr, si), sc), v1 = mix_down(make_float2(xr, xi), sc);
                dtheta += nco_constrain_dev(__fmul_rn(v0.y, a.am.pll_alpha));
                theta  += nco_constrain_dev(__fmul_rn(v0.y, a.am.pll_beta));
                theta  += dtheta;
                const float m = __fdiv_rn(v1.x, a.am.mod_index);
                if (j == g) mine = m;
            }
        }
        if (g < ng) { dc[pz] = mine; dc[pz + kAmRing] = mine; }
        __syncwarp();
        // dc blocker, same shape
        const float *wd = dc + ((cnt + (unsigned)k0 + (unsigned)g + (unsigned)(kAmRing - H)) & (kAmRing - 1));
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < kAmTaps; t++) acc = __fmaf_rn(a.am.dc[t], wd[t], acc);
        // de-emphasis is a recurrence over the outputs in order: every lane runs it, lane g keeps output g
        float outv = acc;
        if constexpr (HAS_DE) {
#pragma unroll
            for (int j = 0; j < kA8Lanes; j++) {
                const float r = __shfl_sync(0xffffffffu, acc, glead + j);
                if (j < ng) {
                    de_v1 = __fmaf_rn(-a.de.a1, de_v1, r);
                    if (j == g) outv = __fmul_rn(a.de.b0, de_v1);
                }
            }
        }
        if (active && g < ng) yrow[k0 + g] = outv;
        __syncwarp();                                     // the rings are rewritten by the next group
    }
    (void)gmask;
    if (active) {
        if (g == 0) { a.am.theta[gch] = theta; a.am.dtheta[gch] = dtheta; if (HAS_DE) a.de.v1[gch] = de_v1; }
        for (int i = g; i < kAmRing; i += kA8Lanes) { a.am.lp_ring[i * CT + gch] = lp[i]; a.am.dc_ring[i * CT + gch] = dc[i]; }
    }
}

typedef void (*AmFn)(const AmTailArgs);

}  // namespace

bool amtail_few(bool has_agc, const AmTailArgs &a)
{
    // the eight-lane kernel: DSB with carrier, gain control (if any) already applied in place on the time-major hand-off
    if (getenv("LQB_NO_AMTAIL8")) return false;
    // (up to 2048 channels: beyond that the carrier loop run on all eight lanes costs more issue slots than the filters'
    // parallelism saves -- measured 0.74 vs 0.79 ms at 1024 channels, 1.24 vs 0.81 ms at 8192)
    return a.C <= 2048 && !a.am.suppressed && !a.am.out_v1 && (!has_agc || a.in_tmajor);
}

const char *amtail_kernel_name(bool has_agc, const AmTailArgs &a)
{
    return amtail_few(has_agc, a) ? "amtail8_kernel" : "amtail_kernel";
}

cudaError_t agc_tmajor_launch(const AmTailArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((a.C + 127) / 128);
    if (a.agc.fast && a.agc.big) agc_tmajor_kernel<true, 16, true><<<grid, 128, 0, stream>>>(a);
    else if (a.agc.fast)         agc_tmajor_kernel<true, 16, false><<<grid, 128, 0, stream>>>(a);
    else                         agc_tmajor_kernel<false, 4, false><<<grid, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t amtail_launch(bool has_agc, bool has_de, const AmTailArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    if (has_agc && a.in_tmajor) {
        // time-major hand-off buffer (chain-internal scratch): gain control in place at full occupancy, then the
        // window-bound demodulator without it
        cudaError_t rc0 = agc_tmajor_launch(a, stream);
        if (rc0 != cudaSuccess) return rc0;
        has_agc = false;
    }
    if (a.am.out_v1 && (has_de || a.am.suppressed)) return cudaErrorInvalidValue;
    if (amtail_few(has_agc, a)) {
        const unsigned grid = (unsigned)((a.C + kA8Cpc - 1) / kA8Cpc);
        if (has_de) amtail8_kernel<true><<<grid, 32 * kA8Warps, 0, stream>>>(a); else amtail8_kernel<false><<<grid, 32 * kA8Warps, 0, stream>>>(a);
        return cudaGetLastError();
    }
    AmFn fn = a.am.out_v1 ? (has_agc ? amtail_kernel<true, false, true> : amtail_kernel<false, false, true>)
            : has_agc ? (has_de ? amtail_kernel<true, true, false> : amtail_kernel<true, false, false>)
                      : (has_de ? amtail_kernel<false, true, false> : amtail_kernel<false, false, false>);
    const size_t smem = (size_t)BT * (W * sizeof(float2) + WDP * sizeof(float)) + 1024 * sizeof(float) + (has_agc ? 128 * sizeof(double2) : 0);
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    fn<<<(unsigned)((a.C + BT - 1) / BT), BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
