// pipe.cu -- biquad cascade -> gain loop -> frequency discriminator (BASELINE config 4) as a three-warp pipeline.
//
// Same arithmetic, operand for operand, as seq_kernel<F_IIR | F_AGC | F_FM> with the single-precision gain loop
// (reference loops iirfilter.hpp:296, agc.hpp:113-127, demod.hpp:216).  What changes is who does what, and why:
//   * measured, a lone warp running the whole chain takes 193 cycles per sample however few channels there are -- 93
//     instructions, 20 of them packed, strung on the gain loop's 84-cycle chain; an in-order warp stalls at the first
//     operand that is not ready, and no static interleave of the three stages keeps it issuing (a lone warp CAN issue
//     one FFMA per cycle given four independent chains: tools/ubench_issue.cu).  Config 4 has 16384 channels = 512
//     such warps for 592 schedulers, so the one-warp-per-32-channels kernel runs at that single-warp pace;
//   * with 16384 channels a row must deliver four times the bandwidth it does in the 65536-channel configs and the rows
//     are 512 KB apart: what counts is how many bytes per row are on their way.
// So a 32-channel group is worked by three warps of one CTA, each with about a third of the instructions,
//     warp 0  TMA ring (NSTG tile boxes of the group in flight) -> cascade (skewed, FFMA2)   -> ring 0
//     warp 1  ring 0 -> gain loop (agc_step_fast: the serial chain, 84 cycles a sample)      -> ring 1
//     warp 2  ring 1 -> discriminator (atan2_fast) -> 64-byte row segments of the output
// handing [32 channels x 16 samples] tiles on through shared-memory rings guarded by mbarriers (full / empty, one
// arrival per hand-off by an elected lane behind a __syncwarp); the sample in front of a tile rides in the tile's row.
// Lane l owns channel l in every warp.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdlib>
#include "params.h"
#include "devmath.cuh"
#include "pipe.h"

namespace lqb {
namespace {

constexpr int PT = 16;                  // samples per tile
constexpr int PR = 2;                   // slots of each warp-to-warp ring
constexpr int NSTG = 6;                 // TMA input ring: 5 tiles of the group (640 B per row) on their way at any time
constexpr int PITCH = PT * 8 + 16;      // bytes per channel row of a tile: odd multiple of 16, conflict-free LDS.128 / STS.128
constexpr int SLOT = 32 * PITCH;
constexpr int BOX = 32 * PT * 8;        // one [32 rows x 128 B] TMA box, 128-byte swizzle

__device__ __forceinline__ void mbar_arrive(void *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int NS, bool BIG>
__global__ void __launch_bounds__(96) pipe_kernel(const __grid_constant__ SeqArgs a)
{
    __shared__ __align__(1024) unsigned char s_in[NSTG][BOX];
    __shared__ __align__(16) unsigned char s_ring[2][PR][SLOT];
    __shared__ unsigned long long s_full[2][PR], s_empty[2][PR], s_inbar[NSTG];

    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long chl = (long long)blockIdx.x * 32 + lane;
    const bool act = chl < a.C;
    const long long cl = act ? chl : 0;                 // idle lanes shadow channel 0 and store nothing
    const long long gch = a.ch0 + cl, CT = a.Ctot, N = a.n;
    const long long ntiles = (N + PT - 1) / PT;

    if (threadIdx.x == 0) {
        for (int r = 0; r < 2; r++) for (int s = 0; s < PR; s++) { mbar_init(&s_full[r][s], 1); mbar_init(&s_empty[r][s], 1); }
        for (int s = 0; s < NSTG; s++) mbar_init(&s_inbar[s], 1);
        mbar_init_fence();
    }
    __syncthreads();

    auto put_row = [&](int ring, int slot, const u64 (&v)[PT]) {
        unsigned char *rw = &s_ring[ring][slot][lane * PITCH];
#pragma unroll
        for (int j = 0; j < PT; j += 2) {
            const float2 p = upk(v[j]), q = upk(v[j + 1]);
            *(float4 *)(rw + j * 8) = make_float4(p.x, p.y, q.x, q.y);
        }
    };
    auto get_row = [&](int ring, int slot, u64 (&v)[PT]) {
        const unsigned char *rw = &s_ring[ring][slot][lane * PITCH];
#pragma unroll
        for (int j = 0; j < PT; j += 2) {
            const float4 f = *(const float4 *)(rw + j * 8);
            v[j] = pk(f.x, f.y); v[j + 1] = pk(f.z, f.w);
        }
    };
    // producer side of a ring: wait until the consumer has released the slot's previous use, fill, publish
    auto acquire_empty = [&](int r, long long t) { if (t >= PR) mbar_wait(&s_empty[r][t % PR], (unsigned)((t / PR - 1) & 1)); };
    auto publish_full  = [&](int r, long long t) { __syncwarp(); if (lane == 0) mbar_arrive(&s_full[r][t % PR]); };
    auto acquire_full  = [&](int r, long long t) { mbar_wait(&s_full[r][t % PR], (unsigned)((t / PR) & 1)); };
    auto release_empty = [&](int r, long long t) { __syncwarp(); if (lane == 0) mbar_arrive(&s_empty[r][t % PR]); };

    if (wid == 0) {
        // ---- biquad cascade: iirfilt_crcf_execute_block, DF-II sections, complex state x real coefficients ----
        u64 iv1[NS], iv2[NS], ca1[NS], ca2[NS], cb0[NS], cb1[NS], cb2[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            ca1[s] = pk(-a.iir.a[s][1], -a.iir.a[s][1]); ca2[s] = pk(-a.iir.a[s][2], -a.iir.a[s][2]);
            cb0[s] = pk(a.iir.b[s][0], a.iir.b[s][0]);   cb1[s] = pk(a.iir.b[s][1], a.iir.b[s][1]);
            cb2[s] = pk(a.iir.b[s][2], a.iir.b[s][2]);
            iv1[s] = pk(a.iir.v[(2 * s + 0) * CT + gch]); iv2[s] = pk(a.iir.v[(2 * s + 1) * CT + gch]);
        }
        const unsigned s_in_sh = (unsigned)__cvta_generic_to_shared(&s_in[0][0]);
        const unsigned swz = (unsigned)(lane & 7);
        // rows past the last channel and samples past the end of the call are zero-filled by the TMA engine
        auto issue = [&](long long t) {
            if (lane == 0) {
                unsigned long long *bar = &s_inbar[t % NSTG];
                mbar_arrive_expect_tx(bar, BOX);
                tma_load_2d(s_in_sh + (unsigned)(t % NSTG) * BOX, &a.tmap, (int)(t * (PT * 2)), (int)(blockIdx.x * 32), bar);
            }
        };
        for (int p = 0; p < NSTG - 1; p++) if (p < ntiles) issue(p);
#pragma unroll 1
        for (long long t = 0; t < ntiles; t++) {
            u64 xs[PT], out[PT], yy[NS];
            mbar_wait(&s_inbar[t % NSTG], (unsigned)((t / NSTG) & 1));
            {
                const unsigned char *rw = &s_in[t % NSTG][lane * (PT * 8)];
#pragma unroll
                for (int j = 0; j < PT; j += 2) {
                    const float4 f = *(const float4 *)(rw + ((((unsigned)j >> 1) ^ swz) << 4));
                    xs[j] = pk(f.x, f.y); xs[j + 1] = pk(f.z, f.w);
                }
            }
            __syncwarp();                          // every lane has its row: the stage read one tile ago can be refilled
            if (t + NSTG - 1 < ntiles) issue(t + NSTG - 1);
            const int nv = (int)((N - t * PT) < PT ? (N - t * PT) : PT);
            if (nv == PT) {
                // skewed: at step k section s works on sample k - s, so the NS updates of a step are independent
#pragma unroll
                for (int kk = 0; kk < PT + NS - 1; kk++) {
#pragma unroll
                    for (int sct = NS - 1; sct >= 0; sct--) {
                        const int j = kk - sct;
                        if (j >= 0 && j < PT) {
                            const u64 in = sct == 0 ? xs[j] : yy[sct - 1];
                            const u64 tt = fma2(ca1[sct], iv1[sct], in);
                            const u64 v0 = fma2(ca2[sct], iv2[sct], tt);
                            u64 y = mul2(cb1[sct], iv1[sct]);
                            y = fma2(cb0[sct], v0, y);
                            y = fma2(cb2[sct], iv2[sct], y);
                            iv2[sct] = iv1[sct]; iv1[sct] = v0; yy[sct] = y;
                            if (sct == NS - 1) out[j] = y;
                        }
                    }
                }
            } else {
                // the call's last, partial tile: only its nv samples move the state
#pragma unroll
                for (int j = 0; j < PT; j++) {
                    out[j] = 0ull;
                    if (j < nv) {
                        u64 in = xs[j];
#pragma unroll
                        for (int s = 0; s < NS; s++) {
                            const u64 tt = fma2(ca1[s], iv1[s], in);
                            const u64 v0 = fma2(ca2[s], iv2[s], tt);
                            u64 y = mul2(cb1[s], iv1[s]);
                            y = fma2(cb0[s], v0, y);
                            y = fma2(cb2[s], iv2[s], y);
                            iv2[s] = iv1[s]; iv1[s] = v0; in = y;
                        }
                        out[j] = in;
                    }
                }
            }
            acquire_empty(0, t);
            put_row(0, (int)(t % PR), out);
            publish_full(0, t);
        }
        if (act) {
#pragma unroll
            for (int s = 0; s < NS; s++) { a.iir.v[(2 * s + 0) * CT + gch] = upk(iv1[s]); a.iir.v[(2 * s + 1) * CT + gch] = upk(iv2[s]); }
        }
    } else if (wid == 1) {
        // ---- gain loop: agc_crcf_execute, unlocked, squelch disabled (devmath.cuh agc_step_fast) ----
        const AgcFast k{a.agc.alpha, a.agc.chi, a.agc.clo, a.agc.cl2, a.agc.chalf, a.agc.scale};
        float g = a.agc.g[gch], y2p = a.agc.y2p[gch];
        u64 last = pk(a.fm.rprime[gch]);                  // the discriminator's r': this stage's newest output
#pragma unroll 1
        for (long long t = 0; t < ntiles; t++) {
            u64 z[PT];
            acquire_full(0, t);
            get_row(0, (int)(t % PR), z);
            release_empty(0, t);
            const u64 before = last;
            const int nv = (int)((N - t * PT) < PT ? (N - t * PT) : PT);
            if (nv == PT) {
#pragma unroll
                for (int j = 0; j < PT; j++) z[j] = pk(agc_step_fast<BIG>(upk(z[j]), g, y2p, k));
                last = z[PT - 1];
            } else {
#pragma unroll
                for (int j = 0; j < PT; j++) if (j < nv) { z[j] = pk(agc_step_fast<BIG>(upk(z[j]), g, y2p, k)); last = z[j]; }
            }
            acquire_empty(1, t);
            put_row(1, (int)(t % PR), z);
            *(float2 *)&s_ring[1][t % PR][lane * PITCH + PT * 8] = upk(before);     // the sample in front of the tile
            publish_full(1, t);
        }
        if (act) { a.agc.g[gch] = g; a.agc.y2p[gch] = y2p; a.fm.rprime[gch] = upk(last); }
    } else {
        // ---- discriminator: freqdem_demodulate, arg(conj(r') r) / (2 pi kf) ----
        float *yrow = (float *)a.y + cl * a.out_pitch;
#pragma unroll 1
        for (long long t = 0; t < ntiles; t++) {
            u64 z[PT];
            acquire_full(1, t);
            get_row(1, (int)(t % PR), z);
            float2 prev = *(const float2 *)&s_ring[1][t % PR][lane * PITCH + PT * 8];
            release_empty(1, t);
            const int nv = (int)((N - t * PT) < PT ? (N - t * PT) : PT);
            float r[PT];
#pragma unroll
            for (int j = 0; j < PT; j++) {
                const float2 c = upk(z[j]);
                const float re = __fmaf_rn(prev.x, c.x, __fmul_rn(prev.y, c.y));
                const float im = __fmaf_rn(prev.x, c.y, -__fmul_rn(prev.y, c.x));
                r[j] = __fmul_rn(atan2_fast(im, re), a.fm.ref);
                prev = c;
            }
            if (act) {
                float *yo = yrow + t * PT;
                if (a.vec_out && nv == PT) {
#pragma unroll
                    for (int j = 0; j < PT; j += 4) *(float4 *)(yo + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < PT; j++) if (j < nv) yo[j] = r[j];
                }
            }
        }
    }
}

typedef void (*PipeFn)(const SeqArgs);
template <bool BIG> PipeFn pick2(int nsos)
{
    switch (nsos) {
    case 1: return pipe_kernel<1, BIG>; case 2: return pipe_kernel<2, BIG>; case 3: return pipe_kernel<3, BIG>; case 4: return pipe_kernel<4, BIG>;
    default: return nullptr;
    }
}

}  // namespace

bool pipe_supported(unsigned mask, int nsos) { return mask == (F_IIR | F_AGC | F_FM) && pick2<false>(nsos) != nullptr; }

// needs a.agc.fast (the single-precision gain loop) and a.tmap = the input as [32 rows x 128 B] boxes
cudaError_t pipe_launch(int nsos, const SeqArgs &a, cudaStream_t stream)
{
    PipeFn fn = a.agc.big ? pick2<true>(nsos) : pick2<false>(nsos);
    if (!fn || !a.agc.fast || !a.use_tma) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    fn<<<(unsigned)((a.C + 31) / 32), 96, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
