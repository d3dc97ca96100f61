// seq.cu -- the channel-parallel, time-sequential chain kernel.
//
// One thread owns one channel and walks its samples in order, exactly as the reference's
// per-object loops do (iirfilt_crcf_execute_block iirfilter.hpp:296, the per-sample resampler loop
// resampler.hpp:164-166, the AGC loop agc.hpp:113-127, ampmodem_demodulate_block demod.hpp:294,
// the de-emphasis loop iirfilter.hpp:388-389) -- but every stage of the chain runs back to back in
// registers, so the full-rate intermediate between them never reaches HBM.  The stage set is a
// compile-time mask; channels are the parallel axis (65536 channels = 443 threads per SM).
//
// Data movement: rows are [channel][time] in HBM, so a thread-per-channel access would be strided
// by a whole row.  Each CTA of BT channels instead stages [BT x TS] tiles through shared memory
// with 16-byte cp.async in a NST-deep ring (each warp-wide copy reads whole 128-byte row segments),
// and every thread then reads its own row with conflict-free 16-byte shared loads (row pitch is an
// odd number of 16-byte units).  Full-rate outputs go back the same way; decimated outputs
// (2.4 % of the samples) are stored directly.
#include <cuda_runtime.h>
#include <math.h>
#include "params.h"
#include "devmath.cuh"
#include "seq.h"

namespace lqb {

namespace {

constexpr int BT  = kSeqBT;    // channels (threads) per CTA
constexpr int TS  = kSeqTS;    // samples per tile row
constexpr int NST = kSeqNST;   // cp.async ring depth

template <int ELEM> struct Pitch { static constexpr int v = TS * ELEM + 16; };

// ---- tile movers -------------------------------------------------------------------------------
template <int ELEM>
__device__ __forceinline__ void load_tile(const SeqArgs &a, unsigned char *stage, long long t0, int tid)
{
    constexpr int CH16 = TS * ELEM / 16;      // 16-byte chunks per row
    constexpr int EPC  = 16 / ELEM;           // elements per chunk
    const char *xg = (const char *)a.x;
#pragma unroll
    for (int c = tid; c < BT * CH16; c += BT) {
        const int row = c / CH16, k = c % CH16;
        const long long e0 = t0 + (long long)k * EPC;
        const long long ch = (long long)blockIdx.x * BT + row;
        unsigned char *dst = stage + row * Pitch<ELEM>::v + k * 16;
        if (ch < a.C && e0 < a.n) {
            const char *src = xg + (ch * a.n + e0) * ELEM;
            const long long rem = (a.n - e0) * ELEM;
            if (a.vec_in) {
                cp_async16(dst, src, rem >= 16 ? 16 : (int)rem);
            } else {
#pragma unroll
                for (int e = 0; e < EPC; e++) {
                    if (e0 + e < a.n) {
                        if (ELEM == 8) ((float2 *)dst)[e] = ((const float2 *)src)[e];
                        else           ((float *)dst)[e]  = ((const float *)src)[e];
                    }
                }
            }
        }
    }
}

template <int ELEM>
__device__ __forceinline__ void store_tile(const SeqArgs &a, const unsigned char *tile, long long t0, int tid)
{
    constexpr int CH16 = TS * ELEM / 16;
    constexpr int EPC  = 16 / ELEM;
    char *yg = (char *)a.y;
#pragma unroll
    for (int c = tid; c < BT * CH16; c += BT) {
        const int row = c / CH16, k = c % CH16;
        const long long e0 = t0 + (long long)k * EPC;
        const long long ch = (long long)blockIdx.x * BT + row;
        const unsigned char *src = tile + row * Pitch<ELEM>::v + k * 16;
        if (ch < a.C && e0 < a.n) {
            char *dst = yg + (ch * a.out_pitch + e0) * ELEM;
            if (a.vec_out && e0 + EPC <= a.n) {
                *(float4 *)dst = *(const float4 *)src;
            } else {
#pragma unroll
                for (int e = 0; e < EPC; e++) {
                    if (e0 + e < a.n) {
                        if (ELEM == 8) ((float2 *)dst)[e] = ((const float2 *)src)[e];
                        else           ((float *)dst)[e]  = ((const float *)src)[e];
                    }
                }
            }
        }
    }
}

// ---- the kernel --------------------------------------------------------------------------------
template <unsigned M, int NSOS>
__global__ void __launch_bounds__(BT) seq_kernel(const __grid_constant__ SeqArgs a)
{
    constexpr bool HAS_NCO = (M & F_NCO) != 0, HAS_IIR = (M & F_IIR) != 0, HAS_RS = (M & F_RS) != 0;
    constexpr bool HAS_AGC = (M & F_AGC) != 0, HAS_AM = (M & F_AM) != 0, HAS_FM = (M & F_FM) != 0;
    constexpr bool HAS_DE = (M & F_DE) != 0, IN_REAL = (M & F_INREAL) != 0;
    constexpr bool OUT_REAL = HAS_AM || HAS_FM || IN_REAL;
    constexpr int  IELEM = IN_REAL ? 4 : 8, OELEM = OUT_REAL ? 4 : 8;
    constexpr int  PIN = Pitch<IELEM>::v, POUT = Pitch<OELEM>::v;
    constexpr int  NS = NSOS > 0 ? NSOS : 1;
    constexpr int  UNR = HAS_AM ? 1 : TS / 2;      // the ampmodem body is large: keep one copy of it
    static_assert(HAS_IIR == (NSOS > 0), "section count and mask disagree");

    extern __shared__ __align__(16) unsigned char smem[];
    unsigned char *s_in  = smem;                                   // NST stages of [BT][PIN]
    unsigned char *s_out = s_in + NST * BT * PIN;                  // [BT][POUT] when the output is full rate
    unsigned char *s_nxt = s_out + (HAS_RS ? 0 : BT * POUT);
    float2 *s_bank = (float2 *)s_nxt;                              // duplicated taps (h, h)
    s_nxt += HAS_RS ? (size_t)a.rs.npfb * a.rs.sublen * sizeof(float2) : 0;
    float2 *s_sincos = (float2 *)s_nxt;                            // oscillator table, full-rate mixing only
    s_nxt += HAS_NCO ? 1024 * sizeof(float2) : 0;
    float2 *s_lpr = (float2 *)s_nxt;                               // ampmodem rings [kAmRing][BT]
    s_nxt += HAS_AM ? (size_t)kAmRing * BT * sizeof(float2) : 0;
    float *s_dcr = (float *)s_nxt;

    const int tid = threadIdx.x;
    const long long chl = (long long)blockIdx.x * BT + tid;        // channel within this launch
    const bool active = chl < a.C;
    const long long gch = a.ch0 + (active ? chl : 0);              // index into the state arrays
    const long long CT = a.Ctot;

    // ---- per-channel state into registers ----
    uint32_t nco_theta = 0, nco_dtheta = 0;
    u64 iv1[NS], iv2[NS], ca1[NS], ca2[NS], cb0[NS], cb1[NS], cb2[NS];
    u64 rs_acc = 0; uint32_t rsP = a.rs.phase;
    float agc_g = 1.f, agc_y2p = 1.f; int agc_mode = 7; unsigned agc_timer = 0, agc_rises = 0;
    uint32_t am_theta = 0, am_dtheta = 0, am_cnt = a.am.count;
    float2 fm_prev = make_float2(0.f, 0.f);
    float de_v1 = 0.f;
    long long kout = 0;

    if constexpr (HAS_NCO) {
        for (int i = tid; i < 1024; i += BT) s_sincos[i] = a.nco.sincos[i];
        nco_theta = a.nco.theta[gch]; nco_dtheta = a.nco.dtheta[gch];
    }
    if constexpr (HAS_IIR) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            ca1[s] = pk(-a.iir.a[s][1], -a.iir.a[s][1]); ca2[s] = pk(-a.iir.a[s][2], -a.iir.a[s][2]);
            cb0[s] = pk(a.iir.b[s][0], a.iir.b[s][0]);   cb1[s] = pk(a.iir.b[s][1], a.iir.b[s][1]);
            cb2[s] = pk(a.iir.b[s][2], a.iir.b[s][2]);
            iv1[s] = pk(a.iir.v[(2 * s + 0) * CT + gch]); iv2[s] = pk(a.iir.v[(2 * s + 1) * CT + gch]);
        }
    }
    if constexpr (HAS_RS) {
        const int nb = a.rs.npfb * a.rs.sublen;
        for (int i = tid; i < nb; i += BT) { float h = a.rs.bank[i]; s_bank[i] = make_float2(h, h); }
    }
    if constexpr (HAS_AGC) {
        agc_g = a.agc.g[gch]; agc_y2p = a.agc.y2p[gch]; agc_mode = a.agc.mode[gch]; agc_timer = a.agc.timer[gch];
    }
    if constexpr (HAS_AM) {
        for (int k = 0; k < kAmRing; k++) {
            s_lpr[k * BT + tid] = a.am.lp_ring[k * CT + gch];
            s_dcr[k * BT + tid] = a.am.dc_ring[k * CT + gch];
        }
        am_theta = a.am.theta[gch]; am_dtheta = a.am.dtheta[gch];
    }
    if constexpr (HAS_FM) fm_prev = a.fm.rprime[gch];
    if constexpr (HAS_DE) de_v1 = a.de.v1[gch];
    __syncthreads();

    // resampler: the first output's window may start before this call; that part comes from the ring
    if constexpr (HAS_RS) {
        const int L = a.rs.sublen;
        const long long nnext = rsP >> 24;
        const unsigned f = (rsP & 0xffffffu) >> (24 - a.rs.bits);
        for (long long j = nnext - (L - 1); j < 0; j++) {
            const int slot = (int)(((long long)a.rs.count + j + 4LL * L) % L);
            const int tap = (int)(j - nnext + L - 1);
            rs_acc = add2(rs_acc, mul2(pk(s_bank[f * L + tap]), pk(a.rs.ring[slot * CT + gch])));
        }
    }

    // ---- what happens to one sample of the chain's decimated/demodulated side ----
    auto tail = [&](float2 z, int jtile) {
        float r = 0.f;
        if constexpr (HAS_AGC) {
            // agc_crcf_execute (liquid agc.proto.c) then the wrapper's status poll, agc.hpp:115-125
            float yr = __fmul_rn(z.x, agc_g), yi = __fmul_rn(z.y, agc_g);
            float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
            agc_y2p = (float)fma(a.agc.one_minus_alpha, (double)agc_y2p, (double)__fmul_rn(a.agc.alpha, y2));
            if (!a.agc.locked) {
                if (agc_y2p > 1e-6f) agc_g = __fmul_rn(agc_g, exp_rn_small(__fmul_rn(__fmul_rn(-0.5f, a.agc.alpha), log_rn(agc_y2p))));
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
            z = make_float2(yr, yi);
        }
        if constexpr (HAS_AM) {
            // ampmodem_demod_dsb_pll_carrier / _costas (liquid ampmodem.c)
            const float2 sc = __ldg(&a.am.sincos[nco_index(am_theta)]);
            float pe, mre;
            if (a.am.suppressed) {
                const float2 v = mix_down(z, sc);
                pe = __fmul_rn(v.y, v.x > 0.f ? 1.f : -1.f);
                mre = v.x;
            } else {
                const unsigned slot = am_cnt & (kAmRing - 1);
                s_lpr[slot * BT + tid] = z;
                float sr = 0.f, si = 0.f;
#pragma unroll
                for (int i = 0; i < kAmTaps; i++) {
                    const float2 w = s_lpr[((am_cnt + (kAmRing - kAmTaps + 1) + i) & (kAmRing - 1)) * BT + tid];
                    sr = __fmaf_rn(a.am.lp[i], w.x, sr); si = __fmaf_rn(a.am.lp[i], w.y, si);
                }
                const float2 x1 = s_lpr[((am_cnt + (kAmRing - kAmDelay)) & (kAmRing - 1)) * BT + tid];
                const float2 v0 = mix_down(make_float2(sr, si), sc), v1 = mix_down(x1, sc);
                pe = v0.y; mre = v1.x;
            }
            am_dtheta += nco_constrain_dev(__fmul_rn(pe, a.am.pll_alpha));
            am_theta  += nco_constrain_dev(__fmul_rn(pe, a.am.pll_beta));
            am_theta  += am_dtheta;
            const float m = __fdiv_rn(mre, a.am.mod_index);
            if (a.am.suppressed) {
                r = m;
            } else {
                const unsigned slot = am_cnt & (kAmRing - 1);
                s_dcr[slot * BT + tid] = m;
                float acc = 0.f;
#pragma unroll
                for (int i = 0; i < kAmTaps; i++)
                    acc = __fmaf_rn(a.am.dc[i], s_dcr[((am_cnt + (kAmRing - kAmTaps + 1) + i) & (kAmRing - 1)) * BT + tid], acc);
                r = acc;
            }
            am_cnt++;
        }
        if constexpr (HAS_FM) {
            // freqdem_demodulate: arg(conj(r') * r) / (2 pi kf)
            const float re = __fmaf_rn(fm_prev.x, z.x, __fmul_rn(fm_prev.y, z.y));
            const float im = __fmaf_rn(fm_prev.x, z.y, -__fmul_rn(fm_prev.y, z.x));
            r = __fmul_rn(atan2f(im, re), a.fm.ref);
            fm_prev = z;
        }
        if constexpr (IN_REAL) r = z.x;
        if constexpr (HAS_DE) {
            // one-pole iirfilt_rrrf: v0 = x - a1*v1 ; y = b0*v0
            de_v1 = __fmaf_rn(-a.de.a1, de_v1, r);
            r = __fmul_rn(a.de.b0, de_v1);
        }
        if constexpr (HAS_RS) {
            if (active) {
                if constexpr (OUT_REAL) ((float *)a.y)[chl * a.out_pitch + kout] = r;
                else                    ((float2 *)a.y)[chl * a.out_pitch + kout] = z;
            }
            kout++;
        } else {
            if constexpr (OUT_REAL) *(float *)(s_out + tid * POUT + jtile * 4) = r;
            else                    *(float2 *)(s_out + tid * POUT + jtile * 8) = z;
        }
    };

    // ---- one full-rate input sample ----
    auto front = [&](float2 xin, long long n, int jtile) {
        u64 x = pk(xin);
        if constexpr (HAS_NCO) {
            float2 sc;
            if (a.nco.type == 0) sc = s_sincos[nco_index(nco_theta)];
            else { const float th = (float)(6.283185307179586 * (double)(float)nco_theta / 4294967296.0); sincosf(th, &sc.x, &sc.y); }
            x = pk(a.nco.dir == 2 ? mix_down(xin, sc) : mix_up(xin, sc));
            nco_theta += nco_dtheta;
        }
        if constexpr (HAS_IIR) {
            // iirfiltsos_execute_df2, both lanes at once:
            //   v0 = fma(-a2, v2, fma(-a1, v1, x)) ;  y = fma(b2, v2, fma(b0, v0, b1*v1))
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const u64 t  = fma2(ca1[s], iv1[s], x);
                const u64 v0 = fma2(ca2[s], iv2[s], t);
                u64 y = mul2(cb1[s], iv1[s]);
                y = fma2(cb0[s], v0, y);
                y = fma2(cb2[s], iv2[s], y);
                iv2[s] = iv1[s]; iv1[s] = v0; x = y;
            }
        }
        if constexpr (HAS_RS) {
            // resamp_execute with step >= sublen * 2^24: the windows of consecutive outputs do not
            // overlap, so the dot product of the pending output accumulates as its samples go by.
            const int L = a.rs.sublen;
            const unsigned cnt = rsP >> 24;
            if (cnt < (unsigned)L) {
                const unsigned f = (rsP & 0xffffffu) >> (24 - a.rs.bits);
                rs_acc = add2(rs_acc, mul2(pk(s_bank[f * L + (L - 1 - cnt)]), x));   // cccf: product rounded, then added
            }
            if (n >= a.n - L && active) a.rs.ring[(int)((a.rs.count + n) % L) * CT + gch] = upk(x);
            if (rsP <= 0x00ffffffu) { tail(upk(rs_acc), jtile); rs_acc = 0; rsP += a.rs.step; }
            rsP -= (1u << 24);
        } else {
            tail(upk(x), jtile);
        }
    };

    // ---- stream the tiles ----
    const long long ntiles = (a.n + TS - 1) / TS;
    for (int p = 0; p < NST - 1; p++) {
        if (p < ntiles) load_tile<IELEM>(a, s_in + p * BT * PIN, (long long)p * TS, tid);
        cp_async_commit();
    }
    for (long long t = 0; t < ntiles; t++) {
        const long long tn = t + NST - 1;
        if (tn < ntiles) load_tile<IELEM>(a, s_in + (int)(tn % NST) * BT * PIN, tn * TS, tid);
        cp_async_commit();
        cp_async_wait<NST - 1>();
        __syncthreads();
        const unsigned char *row = s_in + (int)(t % NST) * BT * PIN + tid * PIN;
        const long long n0 = t * TS;
        const int nv = (int)((a.n - n0) < TS ? (a.n - n0) : TS);
        if (nv == TS) {
            if constexpr (IN_REAL) {
#pragma unroll
                for (int j = 0; j < TS; j += 4) {
                    const float4 v = *(const float4 *)(row + j * 4);
                    front(make_float2(v.x, 0.f), n0 + j, j);     front(make_float2(v.y, 0.f), n0 + j + 1, j + 1);
                    front(make_float2(v.z, 0.f), n0 + j + 2, j + 2); front(make_float2(v.w, 0.f), n0 + j + 3, j + 3);
                }
            } else {
#pragma unroll UNR
                for (int j = 0; j < TS; j += 2) {
                    const float4 v = *(const float4 *)(row + j * 8);
                    front(make_float2(v.x, v.y), n0 + j, j);
                    front(make_float2(v.z, v.w), n0 + j + 1, j + 1);
                }
            }
        } else {
            for (int j = 0; j < nv; j++) {
                if constexpr (IN_REAL) front(make_float2(*(const float *)(row + j * 4), 0.f), n0 + j, j);
                else                   front(*(const float2 *)(row + j * 8), n0 + j, j);
            }
        }
        if constexpr (!HAS_RS) {
            __syncthreads();
            store_tile<OELEM>(a, s_out, n0, tid);
        }
        __syncthreads();
    }

    // ---- carried state back to HBM ----
    if (active) {
        if constexpr (HAS_NCO) a.nco.theta[gch] = nco_theta;
        if constexpr (HAS_IIR) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                a.iir.v[(2 * s + 0) * CT + gch] = upk(iv1[s]); a.iir.v[(2 * s + 1) * CT + gch] = upk(iv2[s]);
            }
        }
        if constexpr (HAS_AGC) {
            a.agc.g[gch] = agc_g; a.agc.y2p[gch] = agc_y2p; a.agc.mode[gch] = agc_mode; a.agc.timer[gch] = agc_timer;
            if (agc_rises) atomicAdd(a.agc.rise_count, agc_rises);
        }
        if constexpr (HAS_AM) {
            for (int k = 0; k < kAmRing; k++) {
                a.am.lp_ring[k * CT + gch] = s_lpr[k * BT + tid];
                a.am.dc_ring[k * CT + gch] = s_dcr[k * BT + tid];
            }
            a.am.theta[gch] = am_theta; a.am.dtheta[gch] = am_dtheta;
        }
        if constexpr (HAS_FM) a.fm.rprime[gch] = fm_prev;
        if constexpr (HAS_DE) a.de.v1[gch] = de_v1;
    }
}

// ---- dispatch ----------------------------------------------------------------------------------
typedef void (*SeqFn)(const SeqArgs);
struct Entry { unsigned mask; int nsos; SeqFn fn; };

#define LQB_E(M, S) { (M), (S), seq_kernel<(M), (S)> }
#define LQB_E_IIR(M) LQB_E(M, 1), LQB_E(M, 2), LQB_E(M, 3), LQB_E(M, 4)
const Entry kTable[] = {
    // single stages
    LQB_E(F_NCO, 0), LQB_E(F_RS, 0), LQB_E(F_AGC, 0), LQB_E(F_AM, 0), LQB_E(F_FM, 0), LQB_E(F_DE | F_INREAL, 0),
    LQB_E_IIR(F_IIR), LQB_E(F_IIR, 5), LQB_E(F_IIR, 6), LQB_E(F_IIR, 7), LQB_E(F_IIR, 8),
    // fused runs
    LQB_E(F_NCO | F_RS, 0),
    LQB_E_IIR(F_IIR | F_RS),
    LQB_E_IIR(F_NCO | F_IIR | F_RS),
    LQB_E(F_AGC | F_AM, 0), LQB_E(F_AM | F_DE, 0), LQB_E(F_AGC | F_AM | F_DE, 0),
    LQB_E(F_AGC | F_FM, 0), LQB_E(F_FM | F_DE, 0), LQB_E(F_AGC | F_FM | F_DE, 0),
    LQB_E_IIR(F_IIR | F_AGC),
    LQB_E_IIR(F_IIR | F_AGC | F_FM),
    LQB_E_IIR(F_IIR | F_RS | F_AGC | F_AM | F_DE),
};
#undef LQB_E
#undef LQB_E_IIR

const Entry *find(unsigned mask, int nsos)
{
    for (const Entry &e : kTable) if (e.mask == mask && e.nsos == nsos) return &e;
    return nullptr;
}

size_t smem_bytes(unsigned m, const SeqArgs &a)
{
    const bool in_real = m & F_INREAL, out_real = (m & (F_AM | F_FM | F_INREAL)) != 0;
    const int pin = TS * (in_real ? 4 : 8) + 16, pout = TS * (out_real ? 4 : 8) + 16;
    size_t b = (size_t)NST * BT * pin;
    if (!(m & F_RS)) b += (size_t)BT * pout;
    if (m & F_RS)  b += (size_t)a.rs.npfb * a.rs.sublen * sizeof(float2);
    if (m & F_NCO) b += 1024 * sizeof(float2);
    if (m & F_AM)  b += (size_t)kAmRing * BT * (sizeof(float2) + sizeof(float));
    return b;
}

}  // namespace

bool seq_supported(unsigned mask, int nsos) { return find(mask, nsos) != nullptr; }

cudaError_t seq_launch(unsigned mask, int nsos, const SeqArgs &a, cudaStream_t stream)
{
    const Entry *e = find(mask, nsos);
    if (!e) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const size_t smem = smem_bytes(mask, a);
    cudaError_t rc = cudaFuncSetAttribute((const void *)e->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const unsigned grid = (unsigned)((a.C + BT - 1) / BT);
    e->fn<<<grid, BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
