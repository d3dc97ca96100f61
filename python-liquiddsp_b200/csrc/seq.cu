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
// with 16-byte cp.async in a 3-deep ring (each warp-wide copy reads whole 128-byte row segments),
// and every thread then reads its own row with conflict-free 16-byte shared loads (row pitch is an
// odd number of 16-byte units).  Each warp stages exactly the 32 rows its own lanes consume and keeps
// its own copy of the tap stream, so the sample loop synchronises warps, never the CTA.  Full-rate outputs go back the same way; decimated outputs
// (2.4 % of the samples) are stored directly.
//
// Instruction budget: the per-sample loop carries nothing but the arithmetic.  Everything that is the
// same for all channels -- which polyphase tap multiplies this sample, whether an output falls on it,
// whether the accumulator restarts -- is worked out once per tile by 16 lanes and left in shared
// memory as a "tap stream"; load addresses are per-thread constants plus a running offset; bounds
// checks and the history-ring save exist only in the checked path taken by the last tiles of a call.
#include <cuda_runtime.h>
#include <math.h>
#include <type_traits>
#include "params.h"
#include "devmath.cuh"
#include "seq.h"

namespace lqb {

namespace {

constexpr int BT  = kSeqBT;    // channels (threads) per CTA
constexpr int TS  = kSeqTS;    // samples per tile row
constexpr int NST = 3;         // staging ring: two tiles in flight while one is consumed

template <int ELEM> struct Geo {
    static constexpr int PITCH = TS * ELEM + 16;      // odd multiple of 16 bytes
    static constexpr int CH16  = TS * ELEM / 16;      // 16-byte chunks per row == chunks per thread
    static constexpr int RSTEP = 32 / CH16;           // rows between a lane's consecutive chunks (a warp loads its own 32 rows)
    static constexpr int EPC   = 16 / ELEM;           // elements per chunk
};

// ---- the kernel --------------------------------------------------------------------------------
// MODE 0: 8 / 16 channels per warp (few channels), cp.async staging
// MODE 1: 32 channels per warp, cp.async staging, padded row pitch
// MODE 2: 32 channels per warp, TMA staging: one elected lane issues one cp.async.bulk.tensor.2d per tile for the
//         warp's whole [32 rows x 128 B] box, 128-byte swizzle (row r's 16-byte chunk j sits at chunk j ^ (r & 7), so
//         the 8 lanes of an LDS.128 phase hit 8 distinct bank groups without padding), completion on a per-warp
//         mbarrier; rows past the last channel and samples past the end of the call are zero-filled by the hardware
template <unsigned M, int NSOS, int MODE>
__global__ void __launch_bounds__(BT, (M & F_AM) ? 2 : 8) seq_kernel(const __grid_constant__ SeqArgs a)
{
    constexpr bool FULL = MODE != 0, TMA = MODE == 2;
    constexpr bool HAS_NCO = (M & F_NCO) != 0, HAS_IIR = (M & F_IIR) != 0, HAS_RS = (M & F_RS) != 0;
    constexpr bool HAS_AGC = (M & F_AGC) != 0, HAS_AM = (M & F_AM) != 0, HAS_FM = (M & F_FM) != 0;
    constexpr bool HAS_DE = (M & F_DE) != 0, IN_REAL = (M & F_INREAL) != 0, IN_I16 = (M & F_INI16) != 0;
    constexpr bool HAS_TF = (M & F_TF) != 0;        // transfer-function IIR: NSOS = delay elements kept in registers
    constexpr bool OUT_REAL = HAS_AM || HAS_FM || IN_REAL;
    constexpr int  IELEM = (IN_REAL || IN_I16) ? 4 : 8, OELEM = OUT_REAL ? 4 : 8;
    static_assert(!IN_I16 || HAS_RS, "int16 ingest is compiled for the decimating front kernels");
    static_assert(!TMA || IELEM == 8, "TMA staging is compiled for complex64 input");
    using GI = Geo<IELEM>; using GO = Geo<OELEM>;
    constexpr int  PIN = GI::PITCH, POUT = GO::PITCH;
    constexpr int  NS = NSOS > 0 ? NSOS : 1;
    constexpr bool BIG_TAIL = HAS_AM;               // keep one copy of the ampmodem body
    static_assert((HAS_IIR || HAS_TF) == (NSOS > 0) && !(HAS_IIR && HAS_TF), "section count and mask disagree");
    static_assert(!HAS_TF || NSOS < kMaxTf, "delay line longer than the coefficient arrays");

    // Channels per warp.  With few channels the kernel is bound by the latency of each channel's recurrence, not by
    // throughput, so the same channels are spread over more warps (8 or 16 working lanes each) and the schedulers
    // get 2-4x as many independent chains to interleave.  The idle lanes still help stage the tiles.
    // (FULL: 32 channels per warp known at compile time -- the many-channel instantiation carries no extra arithmetic)
    const int cpw = FULL ? 32 : a.cpw, RCTA = (BT / 32) * cpw;     // rows (channels) per CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // the swizzled TMA boxes need 1024-byte alignment
    // (offset arithmetic on the shared symbol, not an integer round trip: the pointers stay in the shared address space)
    unsigned char *smem = smem_raw;
    if constexpr (TMA) smem += (1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u;
    constexpr int PINS = TMA ? TS * IELEM : PIN;                   // staged row pitch: dense under TMA, padded otherwise
    unsigned char *s_in  = smem;                                   // NST stages of [RCTA][PINS]
    unsigned char *s_out = s_in + NST * RCTA * PINS;               // [RCTA][POUT] when the output is full rate
    // TMA mode with complex output: two dense swizzled [RCTA][128 B] buffers, sent by bulk tensor stores
    constexpr bool TMA_OUT = TMA && !HAS_RS && !OUT_REAL;
    unsigned char *s_nxt = s_out + (HAS_RS ? 0 : (TMA_OUT ? 2 * RCTA * (TS * 8) : RCTA * POUT));
    float2 *s_tap = (float2 *)s_nxt;                               // [warp][NST][TS] (tap, keep)
    s_nxt += HAS_RS ? (BT / 32) * NST * TS * sizeof(float2) : 0;
    int *s_emit = (int *)s_nxt;                                    // [warp][NST] sample of the tile an output falls on, or -1
    s_nxt += HAS_RS ? 32 : 0;
    float *s_bank = (float *)s_nxt;                                // [npfb][sublen]
    s_nxt += HAS_RS ? (((size_t)a.rs.npfb * a.rs.sublen * sizeof(float) + 15) & ~(size_t)15) : 0;
    float2 *s_sincos = (float2 *)s_nxt;                            // oscillator table, full-rate mixing only
    s_nxt += HAS_NCO ? 1024 * sizeof(float2) : 0;
    double2 *s_log = (double2 *)s_nxt;                             // AGC logarithm table
    s_nxt += HAS_AGC ? 128 * sizeof(double2) : 0;
    float2 *s_lpr = (float2 *)s_nxt;                               // ampmodem rings [kAmRing][BT]
    s_nxt += HAS_AM ? (size_t)kAmRing * BT * sizeof(float2) : 0;
    float *s_dcr = (float *)s_nxt;

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int myrow = wid * cpw + lane;                            // this lane's row of the CTA's tile
    const bool worker = lane < cpw;
    const long long chl = (long long)blockIdx.x * RCTA + myrow;    // channel within this launch
    const bool active = worker && chl < a.C;
    const long long gch = a.ch0 + (active ? chl : 0);              // index into the state arrays
    const long long CT = a.Ctot;
    const long long N = a.n;

    // ---- per-channel state into registers ----
    uint32_t nco_theta = 0, nco_dtheta = 0;
    u64 iv1[NS], iv2[NS], ca1[NS], ca2[NS], cb0[NS], cb1[NS], cb2[NS];
    u64 rs_acc = 0;
    u64 tv[NS];                                                    // transfer-function form: v[1..NS] of the delay line
    float agc_g = 1.f, agc_y2p = 1.f; int agc_mode = 7; unsigned agc_timer = 0, agc_rises = 0;
    uint32_t am_theta = 0, am_dtheta = 0, am_cnt = a.am.count;
    float2 fm_prev = make_float2(0.f, 0.f);
    float de_v1 = 0.f;
    long long kout = 0;
    int obuf = 0;                                                  // TMA_OUT: result buffer of the tile in flight
    const unsigned swz = TMA ? (unsigned)(lane & 7) : 0u;          // this lane's chunk permutation under the 128-byte swizzle
    // byte offset of sample j within a staged row (chunks are permuted under the 128-byte swizzle)
    auto soff = [&](int j) -> unsigned { return TMA ? (((((unsigned)j >> 1) ^ swz) << 4) + (unsigned)(j & 1) * 8u) : (unsigned)j * 8u; };

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
    if constexpr (HAS_TF) {
#pragma unroll
        for (int i = 0; i < NS; i++) tv[i] = pk(a.tf.v[i * CT + gch]);
    }
    if constexpr (HAS_RS) {
        const int nb = a.rs.npfb * a.rs.sublen;
        for (int i = tid; i < nb; i += BT) s_bank[i] = a.rs.bank[i];
    }
    if constexpr (HAS_AGC) {
        for (int i = tid; i < 128; i += BT) s_log[i] = a.agc.logtab[i];
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
    __shared__ unsigned long long s_bar[(BT / 32) * NST];          // TMA mode: one mbarrier per warp and stage
    if constexpr (TMA) {
        if (tid == 0) { for (int i = 0; i < (BT / 32) * NST; i++) mbar_init(&s_bar[i], 1); mbar_init_fence(); }
    }
    __syncthreads();

    // resampler: the first output's window may start before this call; that part comes from the ring
    // (dotprod_cccf arithmetic throughout: each product is rounded, then added -- oldest sample first)
    const int L = HAS_RS ? a.rs.sublen : 0;
    if constexpr (HAS_RS) {
        const long long nnext = a.rs.phase >> 24;
        const unsigned f = (a.rs.phase & 0xffffffu) >> (24 - a.rs.bits);
        float ar = 0.f, ai = 0.f;
        for (long long j = nnext - (L - 1); j < 0; j++) {
            const int slot = (int)(((long long)a.rs.count + j + 4LL * L) % L);
            const float h = s_bank[f * L + (int)(j - nnext + L - 1)];
            const float2 w = a.rs.ring[slot * CT + gch];
            ar = __fadd_rn(ar, __fmul_rn(h, w.x)); ai = __fadd_rn(ai, __fmul_rn(h, w.y));
        }
        rs_acc = pk(ar, ai);
    }

    // ---- the tap stream: lanes 0..TS-1 of each warp follow one sample position of every tile ----
    // gP = resampler phase (8.24) at this lane's sample of the tile being generated, liquid's own
    // recurrence: an output falls on a sample iff P <= 0xffffff (then P += step); P -= 2^24 per sample.
    uint32_t gP = a.rs.phase;
    if constexpr (HAS_RS) {
        if (lane < TS) for (int i = 0; i < lane; i++) { if (gP <= 0x00ffffffu) gP += a.rs.step; gP -= (1u << 24); }
    }
    auto gen_taps = [&](int stage, bool first_tile) {
        if constexpr (HAS_RS) {
            {
                const bool lane_on = lane < TS;
                const bool emit = lane_on && gP <= 0x00ffffffu;
                if (lane_on) {
                    const unsigned cnt = gP >> 24;
                    const unsigned f = (gP & 0xffffffu) >> (24 - a.rs.bits);
                    const float h = cnt < (unsigned)L ? s_bank[f * L + (L - 1 - (int)cnt)] : 0.f;
                    // the sample after an output starts a new dot product: its accumulator is multiplied by 0
                    // (never the first sample of a call: there the accumulator holds the ring's contribution)
                    const float keep = (gP < a.rs.step - (1u << 24) || (first_tile && lane == 0)) ? 1.f : 0.f;
                    s_tap[(wid * NST + stage) * TS + lane] = make_float2(h, keep);
                    // advance one tile (step >= TS * 2^24, so at most one output per tile)
                    if (gP < ((unsigned)TS << 24)) gP += a.rs.step;
                    gP -= ((unsigned)TS << 24);
                }
                const unsigned m = __ballot_sync(0xffffffffu, emit);
                if (lane == 0) s_emit[wid * NST + stage] = m ? (__ffs(m) - 1) : -1;
            }
        }
    };

    // ---- tile loads: per-thread constants + a running byte offset ----
    // thread -> chunk column lk of rows lrow0 + i*RSTEP; rows past the last channel are skipped
    const int lk = lane % GI::CH16, lrow0 = wid * cpw + lane / GI::CH16;
    const int npass = cpw / GI::RSTEP;                            // copies per lane per tile (CH16 when cpw = 32)
    unsigned vmask = 0;
#pragma unroll
    for (int i = 0; i < GI::CH16; i++)
        if (i < npass && (long long)blockIdx.x * RCTA + lrow0 + i * GI::RSTEP < a.C) vmask |= 1u << i;
    const char *gsrc = (const char *)a.x + (((long long)blockIdx.x * RCTA + lrow0) * N + (long long)lk * GI::EPC) * IELEM;
    const long long grow = (long long)GI::RSTEP * N * IELEM;      // bytes between this thread's rows
    const unsigned sdst0 = (unsigned)__cvta_generic_to_shared(s_in) + lrow0 * PIN + lk * 16;

    // the common case -- every row of the CTA exists, the tile is complete, rows are 16-byte aligned -- needs no
    // predicate and no size: eight copies at a running row pointer
    const bool fast_cta = a.vec_in && (long long)(blockIdx.x + 1) * RCTA <= (long long)a.C;
    const long long nfull = N / TS;
    auto load_tile_fast = [&](long long t, int stage) {
        const char *src = gsrc + t * (TS * IELEM);
        const unsigned dst = sdst0 + stage * (RCTA * PIN);
#pragma unroll
        for (int i = 0; i < GI::CH16; i++) {
            if (i < npass) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + i * (GI::RSTEP * PIN)), "l"(src) : "memory");
            src += grow;
        }
    };
    const unsigned s_in_sh = (unsigned)__cvta_generic_to_shared(s_in);
    auto load_tile = [&](long long t, int stage) {
        if constexpr (TMA) {
            if (lane == 0) {
                unsigned long long *bar = &s_bar[wid * NST + stage];
                mbar_arrive_expect_tx(bar, 32 * TS * IELEM);
                tma_load_2d(s_in_sh + stage * (RCTA * PINS) + wid * (32 * PINS), &a.tmap, (int)(t * (TS * 2)), (int)(blockIdx.x * RCTA + wid * 32), bar);
            }
            return;
        }
        if (fast_cta && t < nfull) { load_tile_fast(t, stage); return; }
        const long long e0 = t * TS + (long long)lk * GI::EPC;    // first element of this thread's chunks
        const char *src = gsrc + t * (TS * IELEM);
        if (a.vec_in) {
            const unsigned dst = sdst0 + stage * (RCTA * PIN);
            const long long rem = (N - e0) * IELEM;
            const int nb = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
            if (nb > 0) {
#pragma unroll
                for (int i = 0; i < GI::CH16; i++)
                    if ((vmask >> i) & 1u)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::
                                     "r"(dst + i * (GI::RSTEP * PIN)), "l"(src + i * grow), "r"(nb) : "memory");
            }
        } else {
            unsigned char *d = s_in + stage * (RCTA * PIN) + lrow0 * PIN + lk * 16;
#pragma unroll
            for (int i = 0; i < GI::CH16; i++)
                if ((vmask >> i) & 1u)
                    for (int e = 0; e < GI::EPC; e++)
                        if (e0 + e < N) {
                            if (IELEM == 8) ((float2 *)(d + i * (GI::RSTEP * PIN)))[e] = ((const float2 *)(src + i * grow))[e];
                            else            ((float *)(d + i * (GI::RSTEP * PIN)))[e]  = ((const float *)(src + i * grow))[e];
                        }
        }
    };

    // full-rate output tile back to HBM, same chunk geometry
    const int ok = lane % GO::CH16, orow0 = wid * cpw + lane / GO::CH16, opass = cpw / GO::RSTEP;
    auto store_tile = [&](long long t) {
        const long long e0 = t * TS + (long long)ok * GO::EPC;
#pragma unroll
        for (int i = 0; i < GO::CH16; i++) {
            const long long ch = (long long)blockIdx.x * RCTA + orow0 + i * GO::RSTEP;
            if (i < opass && ch < a.C && e0 < N) {
                const unsigned char *src = s_out + (orow0 + i * GO::RSTEP) * POUT + ok * 16;
                char *dst = (char *)a.y + (ch * a.out_pitch + e0) * OELEM;
                if (a.vec_out && e0 + GO::EPC <= N) {
                    *(float4 *)dst = *(const float4 *)src;
                } else {
                    for (int e = 0; e < GO::EPC; e++)
                        if (e0 + e < N) {
                            if (OELEM == 8) ((float2 *)dst)[e] = ((const float2 *)src)[e];
                            else            ((float *)dst)[e]  = ((const float *)src)[e];
                        }
                }
            }
        }
    };

    // ---- what happens to one sample on the decimated / demodulated side of the chain ----
    const AgcFast agck{a.agc.alpha, a.agc.chi, a.agc.clo, a.agc.cl2, a.agc.chalf, a.agc.scale};
    const bool agc_fast = HAS_AGC && a.agc.fast != 0;             // unlocked, no squelch: the single-precision gain loop
    auto agc_apply = [&](float2 z) -> float2 {
        if constexpr (HAS_AGC) {
            if (agc_fast) return a.agc.big ? agc_step_fast<true>(z, agc_g, agc_y2p, agck) : agc_step_fast<false>(z, agc_g, agc_y2p, agck);
            // agc_crcf_execute (liquid agc.proto.c) then the wrapper's status poll, agc.hpp:115-125
            float yr = __fmul_rn(z.x, agc_g), yi = __fmul_rn(z.y, agc_g);
            float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
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
            z = make_float2(yr, yi);
        }
        return z;
    };
    auto post = [&](float2 z, int jtile) {
        float r = 0.f;
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
            r = __fmul_rn(atan2_fast(im, re), a.fm.ref);
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
                const long long o = a.out_tmajor ? kout * a.out_pitch + chl : chl * a.out_pitch + kout;
                if constexpr (OUT_REAL) ((float *)a.y)[o] = r;
                else                    ((float2 *)a.y)[o] = z;
            }
            kout++;
        } else {
            if constexpr (OUT_REAL) *(float *)(s_out + myrow * POUT + jtile * 4) = r;
            else if constexpr (TMA_OUT) *(float2 *)(s_out + obuf * (RCTA * (TS * 8)) + myrow * (TS * 8) + soff(jtile)) = z;
            else                    *(float2 *)(s_out + myrow * POUT + jtile * 8) = z;
        }
    };

    auto tail = [&](float2 z, int jtile) { post(agc_apply(z), jtile); };
    // the same with the gain loop known to be the short one: no branch, so an unrolled tile is one basic block
    auto tail_fast = [&](float2 z, int jtile, auto big) { post(agc_step_fast<decltype(big)::value>(z, agc_g, agc_y2p, agck), jtile); };

    // ---- one full-rate sample: oscillator and IIR; returns the (complex) value handed on ----
    auto head = [&](float2 xin) -> u64 {
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
        if constexpr (HAS_TF) {
            // iirfilt_execute_norm: v0 = x - sum a[i] v[i] (i ascending), y = sum b[i] v[i] (from zero, i ascending),
            // every step one fused multiply-add; tv[i-1] is v[i] after liquid's shift
            u64 v0 = x;
#pragma unroll
            for (int i = 1; i <= NS; i++) if (i < a.tf.nna) v0 = fma2(pk(a.tf.na[i], a.tf.na[i]), tv[i - 1], v0);
            u64 y = fma2(pk(a.tf.b[0], a.tf.b[0]), v0, 0ull);
#pragma unroll
            for (int i = 1; i <= NS; i++) if (i < a.tf.nb) y = fma2(pk(a.tf.b[i], a.tf.b[i]), tv[i - 1], y);
#pragma unroll
            for (int i = NS - 1; i > 0; i--) tv[i] = tv[i - 1];
            tv[0] = v0; x = y;
        }
        return x;
    };
    // one staged sample / a whole staged row as complex floats (int16 I/Q pairs are converted on the way)
    auto ld1 = [&](const unsigned char *rw, int j) -> float2 {
        if constexpr (IN_I16) return i16_to_iq(*(const unsigned *)(rw + j * 4));
        else if constexpr (TMA) return *(const float2 *)(rw + ((((unsigned)j >> 1) ^ swz) << 4) + (j & 1) * 8);
        else return *(const float2 *)(rw + j * 8);
    };
    auto ld_row = [&](const unsigned char *rw, u64 (&xs)[TS]) {
        if constexpr (IN_I16) {
#pragma unroll
            for (int j = 0; j < TS; j += 4) {
                const uint4 v = *(const uint4 *)(rw + j * 4);
                xs[j] = pk(i16_to_iq(v.x)); xs[j + 1] = pk(i16_to_iq(v.y)); xs[j + 2] = pk(i16_to_iq(v.z)); xs[j + 3] = pk(i16_to_iq(v.w));
            }
        } else if constexpr (TMA) {
#pragma unroll
            for (int j = 0; j < TS; j += 2) {
                const float4 v = *(const float4 *)(rw + ((((unsigned)j >> 1) ^ swz) << 4));
                xs[j] = pk(v.x, v.y); xs[j + 1] = pk(v.z, v.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < TS; j += 2) {
                const float4 v = *(const float4 *)(rw + j * 8);
                xs[j] = pk(v.x, v.y); xs[j + 1] = pk(v.z, v.w);
            }
        }
    };
    // resampler step for one sample: acc = acc*keep + round(tap*x).  keep is 1 (0 right after an
    // output) and comes from shared memory, so the two roundings of liquid's complex-tap dot product
    // (product, then sum) survive as FMUL2 + FFMA2; ptxas contracts a plain mul.f32x2 + add.f32x2 pair
    // into one FFMA2 even with .rn and -fmad=false.
    auto rs_step = [&](u64 x, const float2 tk) { rs_acc = fma2(rs_acc, pk(tk.y, tk.y), mul2(pk(tk.x, tk.x), x)); };

    // ---- stream the tiles ----
    const long long ntiles = (N + TS - 1) / TS;
    // tiles [0, nfast) are complete and need no ring save; the rest take the checked path
    const long long nfast = HAS_RS ? ((N - L) > 0 ? (N - L) / TS : 0) : N / TS;
    for (int p = 0; p < NST - 1; p++) {
        if (p < ntiles) { load_tile(p, p); gen_taps(p, p == 0); }
        cp_async_commit();
    }
    int stage = 0;
    unsigned phases = 0;                       // TMA mode: parity of each stage's mbarrier
#pragma unroll 1
    for (long long t = 0; t < ntiles; t++) {
        if constexpr (TMA) {
            mbar_wait(&s_bar[wid * NST + stage], (phases >> stage) & 1u);
            phases ^= 1u << stage;
        } else {
            cp_async_wait<NST - 2>();
        }
        __syncwarp();                          // this warp's rows of tile t are in shared memory; its lanes are done with tile t-1
        {
            const int sn = stage == 0 ? NST - 1 : stage - 1;        // the stage tile t-1 occupied
            if (t + NST - 1 < ntiles) { load_tile(t + NST - 1, sn); gen_taps(sn, false); }
            cp_async_commit();
        }
        const unsigned char *row = s_in + stage * (RCTA * PINS) + myrow * PINS;
        const float2 *tk = s_tap + (wid * NST + stage) * TS;
        int e = -1;
        if constexpr (HAS_RS) e = s_emit[wid * NST + stage];
        if (!worker) {
            // idle lane of a partially used warp: staging only
        } else if (t < nfast) {
            if constexpr (IN_REAL) {
#pragma unroll
                for (int j = 0; j < TS; j += 4) {
                    // real samples ride in the real lane (the filters are real-coefficient, so that lane is exactly
                    // the rrrf arithmetic; the imaginary lane carries zeros)
                    const float4 v = *(const float4 *)(row + j * 4);
                    tail(upk(head(make_float2(v.x, 0.f))), j);     tail(upk(head(make_float2(v.y, 0.f))), j + 1);
                    tail(upk(head(make_float2(v.z, 0.f))), j + 2); tail(upk(head(make_float2(v.w, 0.f))), j + 3);
                }
            } else if constexpr (HAS_RS && !BIG_TAIL && HAS_IIR && !HAS_NCO) {
                // Biquad cascade, skewed: at step k section s works on sample k-s, so the NS section updates of
                // one step are independent of each other (each consumes what the previous section produced one
                // step earlier).  Same operations on the same operands as the sample-by-sample order -- only
                // the issue order changes, which gives the scheduler NS chains to interleave instead of one.
                // 62 % of the tiles carry no output (one every 41.67 samples): they run a copy of the body without
                // the per-sample capture of the finished dot product.
                auto body = [&](auto with_emit) {
                    u64 xs[TS];
                    ld_row(row, xs);
                    u64 yy[NS], outv = 0;
#pragma unroll
                    for (int k = 0; k < TS + NS - 1; k++) {
#pragma unroll
                        for (int sct = NS - 1; sct >= 0; sct--) {
                            const int j = k - sct;
                            if (j >= 0 && j < TS) {
                                const u64 in = sct == 0 ? xs[j] : yy[sct - 1];
                                const u64 t  = fma2(ca1[sct], iv1[sct], in);
                                const u64 v0 = fma2(ca2[sct], iv2[sct], t);
                                u64 y = mul2(cb1[sct], iv1[sct]);
                                y = fma2(cb0[sct], v0, y);
                                y = fma2(cb2[sct], iv2[sct], y);
                                iv2[sct] = iv1[sct]; iv1[sct] = v0; yy[sct] = y;
                                if (sct == NS - 1) {
                                    rs_step(y, tk[j]);
                                    if constexpr (decltype(with_emit)::value) { if (j == e) outv = rs_acc; }
                                }
                            }
                        }
                    }
                    if constexpr (decltype(with_emit)::value) tail(upk(outv), e);     // at most one output per tile
                };
                if (e < 0) body(std::false_type{}); else body(std::true_type{});
            } else if constexpr (HAS_RS && !BIG_TAIL) {
                u64 outv = 0, xs[TS];
                ld_row(row, xs);
#pragma unroll
                for (int j = 0; j < TS; j++) {
                    rs_step(head(upk(xs[j])), tk[j]);
                    if (j == e) outv = rs_acc;
                }
                if (e >= 0) tail(upk(outv), e);
            } else if constexpr (HAS_RS) {
#pragma unroll 1
                for (int j = 0; j < TS; j++) {
                    rs_step(head(ld1(row, j)), tk[j]);
                    if (j == e) tail(upk(rs_acc), j);
                }
            } else if constexpr (BIG_TAIL) {
#pragma unroll 1
                for (int j = 0; j < TS; j++) tail(upk(head(ld1(row, j))), j);
            } else if constexpr (HAS_IIR && !HAS_NCO) {
                // full-rate chains behind the biquad cascade: the cascade runs skewed over the tile (see above) and
                // leaves its outputs in the thread's own staged row; the stages after it then walk the row in a
                // rolled loop -- the AGC / discriminator bodies are long, 16 unrolled copies would not fit the
                // instruction cache
                // fused == true: the single-precision gain loop (and the discriminator) follow each cascade output in
                // the same unrolled tile.  The gain loop is ONE dependent chain per channel (about 70 cycles a sample) and
                // with 16384 channels a scheduler holds a single warp, so nothing but this warp's own independent work
                // can fill the chain's latency: in one basic block the cascade of the samples ahead and the
                // discriminator of the samples behind are exactly that work
                auto cascade = [&](auto fused, auto big) {
                    u64 xs[TS], yy[NS];
                    ld_row(row, xs);
#pragma unroll
                    for (int k = 0; k < TS + NS - 1; k++) {
#pragma unroll
                        for (int sct = NS - 1; sct >= 0; sct--) {
                            const int j = k - sct;
                            if (j >= 0 && j < TS) {
                                const u64 in = sct == 0 ? xs[j] : yy[sct - 1];
                                const u64 t  = fma2(ca1[sct], iv1[sct], in);
                                const u64 v0 = fma2(ca2[sct], iv2[sct], t);
                                u64 y = mul2(cb1[sct], iv1[sct]);
                                y = fma2(cb0[sct], v0, y);
                                y = fma2(cb2[sct], iv2[sct], y);
                                iv2[sct] = iv1[sct]; iv1[sct] = v0; yy[sct] = y;
                                if (sct == NS - 1) {
                                    if constexpr (decltype(fused)::value) tail_fast(upk(y), j, big);
                                    else if constexpr (HAS_AGC || HAS_FM) *(float2 *)(const_cast<unsigned char *>(row) + soff(j)) = upk(y);
                                    else tail(upk(y), j);
                                }
                            }
                        }
                    }
                };
                if (HAS_AGC && agc_fast) {
                    // (blocks of 4 or 8 samples instead of the whole skewed tile measured 4-5 % slower: instruction fetch is not the limit)
                    if constexpr (HAS_AGC) { if (a.agc.big) cascade(std::true_type{}, std::true_type{}); else cascade(std::true_type{}, std::false_type{}); }
                } else {
                    cascade(std::false_type{}, std::false_type{});
                    // then one pass per remaining stage: the general gain loop (locked / squelch: double-precision
                    // functions, state machine) is a long serial chain per sample and runs rolled;
                    // the discriminator has no feedback, so its 16 samples are independent work for the scheduler
                    if constexpr (HAS_AGC) {
                        unsigned char *rw = const_cast<unsigned char *>(row);
#pragma unroll 1
                        for (int j = 0; j < TS; j++) { float2 *pz = (float2 *)(rw + soff(j)); *pz = agc_apply(*pz); }
                    }
                    if constexpr (HAS_AGC || HAS_FM) {
#pragma unroll 4
                        for (int j = 0; j < TS; j++) post(ld1(row, j), j);
                    }
                }
            } else if constexpr (HAS_AGC || HAS_FM) {
                if (HAS_AGC && agc_fast) {
                    if constexpr (HAS_AGC) {
                        u64 xs[TS];
                        ld_row(row, xs);
                        if (a.agc.big) {
#pragma unroll
                            for (int j = 0; j < TS; j++) tail_fast(upk(head(upk(xs[j]))), j, std::true_type{});
                        } else {
#pragma unroll
                            for (int j = 0; j < TS; j++) tail_fast(upk(head(upk(xs[j]))), j, std::false_type{});
                        }
                    }
                } else {
#pragma unroll 2
                    for (int j = 0; j < TS; j++) tail(upk(head(ld1(row, j))), j);
                }
            } else {
#pragma unroll
                for (int j = 0; j < TS; j += 2) {
                    const float4 v = *(const float4 *)(row + soff(j));
                    tail(upk(head(make_float2(v.x, v.y))), j);
                    tail(upk(head(make_float2(v.z, v.w))), j + 1);
                }
            }
        } else {
            const long long n0 = t * TS;
            const int nv = (int)((N - n0) < TS ? (N - n0) : TS);
#pragma unroll 1
            for (int j = 0; j < nv; j++) {
                if constexpr (IN_REAL) {
                    tail(upk(head(make_float2(*(const float *)(row + j * 4), 0.f))), j);
                } else {
                    const u64 x = head(ld1(row, j));
                    if constexpr (HAS_RS) {
                        rs_step(x, tk[j]);
                        if (n0 + j >= N - L && active) a.rs.ring[(int)((a.rs.count + n0 + j) % L) * CT + gch] = upk(x);
                        if (j == e) tail(upk(rs_acc), j);
                    } else {
                        tail(upk(x), j);
                    }
                }
            }
        }
        if constexpr (TMA_OUT) {
            // the warp's [32 rows x 128 B] result box goes out as one bulk tensor store; the other buffer takes the next
            // tile, and a buffer is written again only after the store that last read it has drained
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&a.tmap_out, (int)(t * (TS * 2)), (int)(blockIdx.x * RCTA + wid * 32),
                             (unsigned)__cvta_generic_to_shared(s_out + obuf * (RCTA * (TS * 8)) + wid * (32 * TS * 8)));
                bulk_commit();
                bulk_wait_read<1>();
            }
            obuf ^= 1;
            __syncwarp();
        } else if constexpr (!HAS_RS) {
            __syncwarp();
            store_tile(t);
        }
        stage = stage + 1 == NST ? 0 : stage + 1;
    }

    if constexpr (TMA_OUT) { if (lane == 0) bulk_wait_all<0>(); }
    // ---- carried state back to HBM ----
    if (active) {
        if constexpr (HAS_NCO) a.nco.theta[gch] = nco_theta;
        if constexpr (HAS_IIR) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                a.iir.v[(2 * s + 0) * CT + gch] = upk(iv1[s]); a.iir.v[(2 * s + 1) * CT + gch] = upk(iv2[s]);
            }
        }
        if constexpr (HAS_TF) {
#pragma unroll
            for (int i = 0; i < NS; i++) a.tf.v[i * CT + gch] = upk(tv[i]);
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
struct Entry { unsigned mask; int nsos; SeqFn fn, fn_part, fn_tma; };       // full warps / 8-16 channels per warp / TMA staging

#define LQB_E(M, S) { (M), (S), seq_kernel<(M), (S), 1>, seq_kernel<(M), (S), 0>, nullptr }
#define LQB_T(M, S) { (M), (S), seq_kernel<(M), (S), 1>, seq_kernel<(M), (S), 0>, seq_kernel<(M), (S), 2> }
#define LQB_E_IIR(M) LQB_E(M, 1), LQB_E(M, 2), LQB_E(M, 3), LQB_E(M, 4)
const Entry kTable[] = {
    // single stages
    LQB_T(F_NCO, 0), LQB_T(F_RS, 0), LQB_T(F_AGC, 0), LQB_T(F_FM, 0), LQB_E(F_DE | F_INREAL, 0),
    LQB_T(F_IIR, 1), LQB_T(F_IIR, 2), LQB_T(F_IIR, 3), LQB_T(F_IIR, 4), LQB_E(F_IIR, 5), LQB_E(F_IIR, 6), LQB_E(F_IIR, 7), LQB_E(F_IIR, 8),
    LQB_E_IIR(F_IIR | F_INREAL), LQB_E(F_IIR | F_INREAL, 5), LQB_E(F_IIR | F_INREAL, 6), LQB_E(F_IIR | F_INREAL, 7), LQB_E(F_IIR | F_INREAL, 8),
    // transfer-function IIR (CIIRFilter / RIIRFilter), delay elements padded to 1, 2, 4, 8, 15
    LQB_E(F_TF, 1), LQB_E(F_TF, 2), LQB_E(F_TF, 4), LQB_E(F_TF, 8), LQB_E(F_TF, 15),
    LQB_E(F_TF | F_INREAL, 1), LQB_E(F_TF | F_INREAL, 2), LQB_E(F_TF | F_INREAL, 4), LQB_E(F_TF | F_INREAL, 8), LQB_E(F_TF | F_INREAL, 15),
    // fused runs
    LQB_T(F_NCO | F_RS, 0),
    LQB_T(F_IIR | F_RS, 1), LQB_T(F_IIR | F_RS, 2), LQB_T(F_IIR | F_RS, 3), LQB_T(F_IIR | F_RS, 4),
    LQB_E(F_INI16 | F_NCO | F_RS, 0), LQB_E_IIR(F_INI16 | F_IIR | F_RS),      // int16 I/Q ingest fused into the front kernel
    LQB_E(F_AGC | F_FM, 0), LQB_E(F_FM | F_DE, 0), LQB_E(F_AGC | F_FM | F_DE, 0),
    LQB_E_IIR(F_IIR | F_AGC | F_FM),
    LQB_E(F_IIR | F_RS | F_AGC | F_AM | F_DE, 4),
};
#undef LQB_E
#undef LQB_T
#undef LQB_E_IIR

const Entry *find(unsigned mask, int nsos)
{
    for (const Entry &e : kTable) if (e.mask == mask && e.nsos == nsos) return &e;
    return nullptr;
}

size_t smem_bytes(unsigned m, const SeqArgs &a, bool tma)
{
    const bool in_real = (m & (F_INREAL | F_INI16)) != 0, out_real = (m & (F_AM | F_FM | F_INREAL)) != 0;
    const int pin = TS * (in_real ? 4 : 8) + 16, pout = TS * (out_real ? 4 : 8) + 16;
    const size_t rows = (size_t)(BT / 32) * a.cpw;
    size_t b = tma ? (size_t)NST * rows * (pin - 16) + 1024 : (size_t)NST * rows * pin;
    if (!(m & F_RS)) b += (tma && !out_real) ? 2 * rows * (size_t)(TS * 8) : rows * pout;
    if (m & F_RS)  b += (BT / 32) * NST * TS * sizeof(float2) + 32 + (((size_t)a.rs.npfb * a.rs.sublen * sizeof(float) + 15) & ~(size_t)15);
    if (m & F_NCO) b += 1024 * sizeof(float2);
    if (m & F_AGC) b += 128 * sizeof(double2);
    if (m & F_AM)  b += (size_t)kAmRing * BT * (sizeof(float2) + sizeof(float));
    return b;
}

}  // namespace

bool seq_supported(unsigned mask, int nsos) { return find(mask, nsos) != nullptr; }
bool seq_has_tma(unsigned mask, int nsos) { const Entry *e = find(mask, nsos); return e && e->fn_tma; }

cudaError_t seq_launch(unsigned mask, int nsos, const SeqArgs &a, cudaStream_t stream)
{
    const Entry *e = find(mask, nsos);
    if (!e) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    if (a.cpw != 8 && a.cpw != 16 && a.cpw != 32) return cudaErrorInvalidValue;
    if ((mask & F_AM) && a.cpw != 32) return cudaErrorInvalidValue;       // the in-kernel ampmodem rings are per thread
    const bool tma = a.use_tma && a.cpw == 32 && e->fn_tma;
    const size_t smem = smem_bytes(mask, a, tma);
    SeqFn fn = tma ? e->fn_tma : (a.cpw == 32 ? e->fn : e->fn_part);
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const int rows = (BT / 32) * a.cpw;
    const unsigned grid = (unsigned)((a.C + rows - 1) / rows);
    fn<<<grid, BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
