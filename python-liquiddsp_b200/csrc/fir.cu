// fir.cu -- firfilt_crcf: y[n] = scale * sum_k h[k] x[n-k], real taps, complex samples.
//
// Stands in for the push/execute loop of firfilt_*_execute_block (reference call sites
// firfilter.hpp:33, demod.hpp:135-136).  A FIR has no recurrence, so time is the parallel axis:
// each CTA takes one [1 x TN] stretch of one channel, stages it (plus the ntaps-1 samples of
// history in front of it) in shared memory with cp.async, and every thread produces R = 16
// consecutive outputs from a register-resident sliding window -- per tap one broadcast tap load,
// one new sample load and 16 FFMA2.  A per-channel FIR is a bandwidth-bound matrix-vector product
// and stays on the FP32 pipe (no tensor cores).  History is carried between calls in a ping-pong
// pair of [channel][ntaps-1] arrays.
#include <cuda_runtime.h>
#include "params.h"
#include "devmath.cuh"
#include "fir.h"

namespace lqb {
namespace {

constexpr int R   = 16;            // outputs per thread
constexpr int NT  = kFirThreads;   // threads per CTA
constexpr int TN  = R * NT;        // outputs per CTA

// one pad slot after every 16 samples: a thread's window starts 16 samples after its neighbour's,
// so with the pad the 16 lanes of a half-warp land on 16 distinct bank pairs
__device__ __forceinline__ int phys(int i) { return i + (i >> 4); }

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}

// IN_REAL / OUT_REAL: float rows instead of complex64.  firfilt_rrrf is real in, real out: the samples ride in the
// real lane of the same arithmetic.  The firhilbf users (SSBDemod, HilbertTransform) run the two lanes with different
// taps -- a pure delay in one, the quadrature filter in the other -- and combine them when the tile is written.
// SKIP: 1 / 2 = the taps at even / odd positions are zero in both lanes except (at most) position a.skip_keep -- the
// Hilbert-pair filters (firhilbf: quadrature taps on every other position, the in-phase lane a pure delay); those
// positions' multiply-adds are not issued
template <bool IN_REAL, bool OUT_REAL, int SKIP, bool UTAP>
__global__ void __launch_bounds__(NT, 4) fir_kernel(const __grid_constant__ FirArgs a, const int ntiles, const int ntaps_pad, const int tpc, const int groups)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int halo = ntaps_pad - 1;
    const int bufsz = (phys(TN + halo) + 2) & ~1;                      // float2 slots per staged tile
    float2 *s_buf = (float2 *)smem_raw;                                // two tiles: one computing, one arriving
    float2 *s_h = s_buf + 2 * bufsz;                                   // ntaps_pad taps, (in-phase lane, quadrature lane)

    const int tid = threadIdx.x;
    // a CTA walks tpc consecutive tiles of one channel: the next tile streams in (cp.async) while this one is computed
    const long long ch = blockIdx.x / groups;
    const long long tile_first = (long long)(blockIdx.x % groups) * tpc;
    const int my_tiles = (int)((ntiles - tile_first) < tpc ? (ntiles - tile_first) : tpc);
    // firfilt_rrrf with a.pair: a logical row is a PAIR of real channels (2 ch, 2 ch + 1) riding in the two lanes of the
    // same packed arithmetic -- half the multiply-adds per real sample; histories are kept per pair
    const bool pair = IN_REAL && OUT_REAL && a.pair;
    const long long chA = pair ? 2 * ch : ch, chB = chA + 1;
    const bool hasB = pair && chB < a.C;
    const long long gch = pair ? a.ch0 / 2 + ch : a.ch0 + ch;
    const float2 *xrow = a.x + ch * a.n;
    const float *xrow_r = (const float *)a.x + chA * a.n, *xrow_b = (const float *)a.x + chB * a.n;
    const float2 *hrow = a.hist_in + gch * (long long)(a.ntaps - 1);
    const int nh = a.ntaps - 1;

    const bool dup = a.mode == FIR_R2C;                                // real input feeds both lanes
    auto lift = [&](long long g) { const float v = xrow_r[g]; return make_float2(v, dup ? v : (hasB ? xrow_b[g] : 0.f)); };
    for (int k = tid; k < ntaps_pad; k += NT) {
        const float h = k < a.ntaps ? a.taps[k] : 0.f;
        s_h[k] = make_float2(h, a.taps_q ? (k < a.ntaps ? a.taps_q[k] : 0.f) : h);
    }
    auto stage_tile = [&](long long tile, float2 *s_x) {
        const long long t0 = tile * TN;
        if (!IN_REAL && t0 - halo >= 0 && t0 + TN <= a.n) {
            // interior tile of complex samples: no range checks
            const float2 *src = xrow + (t0 - halo);
            for (int i = tid, pi = phys(tid); i < TN + halo; i += NT, pi += (NT / 16) * 17) cp_async8(&s_x[pi], src + i);
            return;
        }
        // i advances by NT = 8 * 16 per pass, so its padded position advances by a constant 8 * 17
        for (int i = tid, pi = phys(tid); i < TN + halo; i += NT, pi += (NT / 16) * 17) {
            const long long g = t0 - halo + i;                         // global sample index
            float2 *dst = &s_x[pi];
            if (g >= 0) {
                if (g >= a.n) *dst = make_float2(0.f, 0.f);
                else if (IN_REAL) {
                    if (dup) *dst = lift(g);
                    else { cp_async4(&dst->x, xrow_r + g); if (hasB) cp_async4(&dst->y, xrow_b + g); else dst->y = 0.f; }
                }
                else cp_async8(dst, xrow + g);
            }
            else if (g + nh >= 0) cp_async8(dst, hrow + (g + nh));
            else *dst = make_float2(0.f, 0.f);
        }
    };
    stage_tile(tile_first, s_buf);
    cp_async_commit();

    // the CTA holding a channel's last tile hands the newest ntaps-1 inputs to the next call
    if (tile_first + my_tiles == ntiles) {
        float2 *ho = a.hist_out + gch * (long long)nh;
        for (int j = tid; j < nh; j += NT) {
            const long long g = a.n - nh + j;
            ho[j] = g >= 0 ? (IN_REAL ? lift(g) : xrow[g]) : hrow[g + nh];
        }
    }

    for (int tt = 0; tt < my_tiles; tt++) {
        float2 *s_x = s_buf + (tt & 1) * bufsz;
        const long long t0 = (tile_first + tt) * TN;                   // first output of this tile
        if (tt + 1 < my_tiles) stage_tile(tile_first + tt + 1, s_buf + ((tt + 1) & 1) * bufsz);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        // window slot q holds the sample at logical index ibase + q; output r with tap k = kb + kk reads
        // logical index o + r + halo - k = ibase + (r - kk + 15) with ibase = o + halo - kb - 15
        u64 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = 0ull;
        // halo + 1 = ntaps_pad is a multiple of 16 and o = 16 * tid, so ibase = 16 * m with m = tid + ntaps_pad/16 - 1 - kb/16:
        // the padded position of ibase + q is 17 * m + q + (q >> 4) -- a per-chunk base plus compile-time offsets.
        // Two register windows that swap roles every 16 taps: the outputs of a chunk read window slots 0..30, of which
        // 0..15 are the chunk's own (older) samples and 16..30 are slots 0..14 of the chunk before -- so the previous
        // chunk's array simply becomes the upper half and nothing is moved.
        u64 Wa[R], Wb[R];
        const float2 *wb = s_x + 17 * (tid + ntaps_pad / R - 1);
#pragma unroll
        for (int q = 0; q < R; q++) Wa[q] = pk(wb[q]);
#pragma unroll
        for (int q = 0; q < R - 1; q++) Wb[q] = pk(wb[R + 1 + q]);      // slots 16..30 sit one pad further (q + (q >> 4))
        auto chunk = [&](const u64 (&lo)[R], const u64 (&hi)[R], int kb) {
#pragma unroll
            for (int kk = 0; kk < R; kk += 2) {
                u64 tap0, tap1;
                if constexpr (UTAP) {                                   // warp-uniform taps from the constant bank
                    const float h0 = a.taps_c[kb + kk], h1 = a.taps_c[kb + kk + 1];
                    tap0 = pk(h0, h0); tap1 = pk(h1, h1);
                } else {
                    const float4 t2 = *(const float4 *)&s_h[kb + kk];   // two taps per load
                    tap0 = pk(t2.x, t2.y); tap1 = pk(t2.z, t2.w);
                }
                if (SKIP != 1 || kb + kk == a.skip_keep) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int q = r - kk + (R - 1);
                        acc[r] = fma2(tap0, q < R ? lo[q] : hi[q - R], acc[r]);
                    }
                }
                if (SKIP != 2 || kb + kk + 1 == a.skip_keep) {
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const int q = r - kk - 1 + (R - 1);
                        acc[r] = fma2(tap1, q < R ? lo[q] : hi[q - R], acc[r]);
                    }
                }
            }
        };
        for (int kb = 0; kb < ntaps_pad; kb += 2 * R) {
            chunk(Wa, Wb, kb);
            if (kb + R < ntaps_pad) {
                wb -= 17;
#pragma unroll
                for (int q = 0; q < R; q++) Wb[q] = pk(wb[q]);
                chunk(Wb, Wa, kb + R);
                if (kb + 2 * R < ntaps_pad) {
                    wb -= 17;
#pragma unroll
                    for (int q = 0; q < R; q++) Wa[q] = pk(wb[q]);
                }
            }
        }
        __syncthreads();                                               // every thread is done reading this tile's samples
        const u64 sc = pk(a.scale, a.scale);
#pragma unroll
        for (int r = 0; r < R; r++) {
            float2 v = upk(mul2(acc[r], sc));
            if (a.mode != FIR_PLAIN) {
                const unsigned long long kabs = a.count + (unsigned long long)(t0 + R * tid + r);
                const long long kloc = t0 + R * tid + r;
                if (a.mode == FIR_SSB_LSB || a.mode == FIR_SSB_USB) {
                    v.x = a.mode == FIR_SSB_LSB ? __fadd_rn(v.x, v.y) : __fsub_rn(v.x, v.y);
                    if (a.post_div != 0.f) v.x = __fdiv_rn(__fmul_rn(0.5f, v.x), a.post_div);
                }
                else if (a.mode == FIR_C2R) v.x = kabs < (unsigned long long)a.delay ? 0.f : (((kabs - a.delay) & 1ull) ? -v.y : v.y);
                else {                                                  // FIR_R2C
                    if (kloc == a.zero_at[0] || kloc == a.zero_at[1] || kloc == a.zero_at[2] || kloc == a.zero_at[3]) v.x = 0.f;
                    if (kabs & 1ull) { v.x = -v.x; v.y = -v.y; }
                }
            }
            s_x[17 * tid + r] = v;                                      // phys(16 * tid + r): the tile buffer doubles as output staging
        }
        __syncthreads();
        float2 *yrow = a.y + ch * a.n;
        if (OUT_REAL) {
            float *yr = (float *)a.y + chA * a.n, *yb = (float *)a.y + chB * a.n;
            for (int i = tid; i < TN; i += NT) {
                const long long g = t0 + i;
                if (g < a.n) { const float2 v = s_x[phys(i)]; yr[g] = v.x; if (hasB) yb[g] = v.y; }
            }
        } else {
            const bool vec = ((a.n & 1) == 0) && ((((size_t)a.y) & 15) == 0);
            if (vec && t0 + TN <= a.n) {
                // sample pair 2i, 2i+1 sits at padded position 2i + (i >> 3); i advances by NT = 128, the position by 272
                float2 *yo = yrow + t0;
#pragma unroll
                for (int it = 0; it < TN / 2 / NT; it++) {
                    const int i = tid + it * NT, pi = 2 * tid + (tid >> 3) + it * (2 * NT + NT / 8);
                    const float2 u = s_x[pi], v = s_x[pi + 1];
                    *(float4 *)(yo + 2 * i) = make_float4(u.x, u.y, v.x, v.y);
                }
            } else if (vec) {
                for (int i = tid, pi = 2 * tid + (tid >> 3); i < TN / 2; i += NT, pi += 2 * NT + NT / 8) {
                    const long long g = t0 + 2 * i;
                    if (g < a.n) {
                        const float2 u = s_x[pi], v = s_x[pi + 1];
                        *(float4 *)(yrow + g) = make_float4(u.x, u.y, v.x, v.y);
                    }
                }
            } else {
                for (int i = tid; i < TN; i += NT) { const long long g = t0 + i; if (g < a.n) yrow[g] = s_x[phys(i)]; }
            }
        }
        __syncthreads();                                               // staging slots are free before the tile after next lands here
    }
}

}  // namespace

cudaError_t fir_launch(const FirArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const int ntaps_pad = (a.ntaps + R - 1) / R * R;
    const int halo = ntaps_pad - 1;
    const long long ntiles = (a.n + TN - 1) / TN;
    const long long rows = (a.pair && a.real_io) ? ((long long)a.C + 1) / 2 : (long long)a.C;      // logical rows (pairs of real channels)
    // consecutive tiles per CTA: as many (up to 8) as still leave ~8 CTAs per SM
    long long tpc = ntiles * rows / (148 * 8);
    tpc = tpc < 1 ? 1 : (tpc > 8 ? 8 : tpc);
    if (tpc > ntiles) tpc = ntiles;
    const long long groups = (ntiles + tpc - 1) / tpc;
    const size_t bufsz = (size_t)((TN + halo + ((TN + halo) >> 4) + 2) & ~1);
    const size_t smem = (2 * bufsz + (size_t)ntaps_pad) * sizeof(float2);
    if (smem > 200 * 1024 || groups * rows > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (a.pair && (!a.real_io || a.mode != FIR_PLAIN || (a.ch0 & 1))) return cudaErrorInvalidValue;
    const bool in_real = a.real_io || a.in_real, out_real = a.real_io || a.out_real;
    const int skip = in_real ? 0 : a.skip;                             // compiled for complex input (the Hilbert-pair users)
    const bool ut = a.utap && a.taps_q == nullptr && ntaps_pad <= kFirUTaps;
    typedef void (*FirFn)(const FirArgs, const int, const int, const int, const int);
    FirFn fn = in_real ? (out_real ? (ut ? fir_kernel<true, true, 0, true> : fir_kernel<true, true, 0, false>) : fir_kernel<true, false, 0, false>)
             : out_real ? (skip == 1 ? fir_kernel<false, true, 1, false> : skip == 2 ? fir_kernel<false, true, 2, false> : fir_kernel<false, true, 0, false>)
                        : (skip == 1 ? fir_kernel<false, false, 1, false> : skip == 2 ? fir_kernel<false, false, 2, false>
                                     : (ut ? fir_kernel<false, false, 0, true> : fir_kernel<false, false, 0, false>));
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    fn<<<(unsigned)(groups * rows), NT, smem, stream>>>(a, (int)ntiles, ntaps_pad, (int)tpc, (int)groups);
    return cudaGetLastError();
}

}  // namespace lqb
