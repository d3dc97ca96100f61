// par.cu -- time-parallel kernels: stages whose outputs are closed-form functions of the input
// position, so time (not channels) can be the parallel axis.
//
// resamp_par: resamp_cccf_execute (reference loop resampler.hpp:164-166) for ANY rate.  From the
// phase p0 at the first input of the call, output k sits at fixed-point position p0 + k*step:
// its newest input is (p0 + k*step) >> 24 and its sub-filter is the next `bits` bits -- the same
// integers liquid's per-sample loop reaches, with no loop-carried dependence.  One CTA takes KT
// consecutive outputs of one channel, stages the input span they touch in shared memory with
// coalesced 8-byte cp.async (history ring for indices before the call), and each thread forms one
// output as a sublen-tap dot product against the bank held in shared memory.
#include <cuda_runtime.h>
#include "params.h"
#include "devmath.cuh"
#include "par.h"
#include <algorithm>
#include <cstdlib>

namespace lqb {
namespace {

constexpr int NT = 128;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}

// oscillator value for absolute sample g of this call: theta_g = theta_0 + g * d_theta (mod 2^32) -- closed form
__device__ __forceinline__ float2 nco_at(const NcoP &q, const float2 *tab, uint32_t th0, uint32_t dth, long long g)
{
    const uint32_t th = th0 + (uint32_t)g * dth;
    if (q.type == 0) return tab[nco_index(th)];
    float2 sc; const float f = (float)(6.283185307179586 * (double)(float)th / 4294967296.0);
    sincosf(f, &sc.x, &sc.y);
    return sc;
}

__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(d), "l"(gsrc) : "memory");
}

// VAR 0: resamp_cccf (ComplexResampler) -- complex taps: each product rounded, then added
// VAR 1: resamp_crcf (CResampler)       -- real taps on complex samples: fused multiply-add chain per lane
// VAR 2: resamp_rrrf (RResampler, RealResampler) -- the same on real samples (float rows in and out)
template <bool HAS_NCO, int VAR>
__global__ void __launch_bounds__(NT)
resamp_par_kernel(const ResampP p, const NcoP q, const void *__restrict__ xv, void *__restrict__ yv, int ch0, int Ctot,
                  long long n, long long n_out, int KT, int ntiles, int span_max, long long nwork)
{
    constexpr bool REAL = VAR == 2;
    static_assert(!(REAL && HAS_NCO), "the mixer runs on complex samples");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s_x0 = (float2 *)smem_raw;                  // span_max samples (float2 slots also when the samples are real)
    // bank [npfb][LP]: rows padded to an odd pitch -- a warp's threads read tap i of up to 16 different sub-filters at once,
    // and with a pitch of 40 floats those rows start on only four distinct banks
    float  *s_b = (float *)(s_x0 + span_max);
    const int tid = threadIdx.x, L = p.sublen, LP = p.sublen | 1;
    // The oscillator's 1024-entry (sin, cos) table in shared memory; the CTAs are persistent (as many as fit at once, each
    // walking many (channel, tile) work items), so the table and the filter bank are loaded once per CTA
    float2 *s_tab = (float2 *)(s_b + ((p.npfb * LP + 3) & ~3));
    __shared__ unsigned long long s_bar;
    const bool bulk = !REAL && ((((size_t)xv) & 15) == 0) && ((n & 1) == 0);
    if (bulk && tid == 0) { mbar_init(&s_bar, 1); mbar_init_fence(); }
    for (int i = tid; i < p.npfb * L; i += NT) s_b[(i / L) * LP + i % L] = p.bank[i];
    if (HAS_NCO && q.type == 0) for (int i = tid; i < 1024; i += NT) s_tab[i] = q.sincos[i];
    __syncthreads();
    unsigned parity = 0;
    for (long long work = blockIdx.x; work < nwork; work += gridDim.x) {
    float2 *s_x = s_x0; float *s_xr = (float *)smem_raw;
    const long long ch = work / ntiles, tile = work % ntiles;
    const long long k0 = tile * KT;
    const int nk = (int)((n_out - k0) < KT ? (n_out - k0) : KT);
    const unsigned long long P0 = (unsigned long long)p.phase + (unsigned long long)k0 * p.step;
    const unsigned long long P1 = (unsigned long long)p.phase + (unsigned long long)(k0 + nk - 1) * p.step;
    const long long i_hi = (long long)(P1 >> 24);
    long long i_lo = (long long)(P0 >> 24) - (L - 1);
    const float2 *xrow = (const float2 *)xv + ch * n;
    const float *xrow_r = (const float *)xv + ch * n;
    const long long gch = ch0 + ch;
    // Complex rows that are 16-byte aligned arrive as ONE bulk copy (cp.async.bulk, the TMA engine) issued by one thread:
    // the staged origin moves down to an even sample so source and destination are 16-byte aligned; the few samples from
    // before the call (history ring) and an odd last sample are filled in by the threads.
    if (bulk) {
        if (i_lo >= 0) i_lo &= ~1LL;
        else if ((-i_lo) & 1) s_x += 1;                 // sample 0 of the call lands on an even slot
    }
    const int span = (int)(i_hi - i_lo + 1);
    const long long gA = i_lo > 0 ? i_lo : 0;           // first sample of this call in the span (even when bulk)
    const int cnt2 = bulk ? (int)((i_hi + 1 - gA) & ~1LL) : 0;
    if (bulk && tid == 0 && cnt2 > 0) {
        fence_async_smem();                             // (the previous work item's generic-proxy writes to the span come first)
        mbar_arrive_expect_tx(&s_bar, (unsigned)cnt2 * 8u);
        bulk_g2s((unsigned)__cvta_generic_to_shared(&s_x[gA - i_lo]), xrow + gA, (unsigned)cnt2 * 8u, &s_bar);
    }
    for (int i = tid; i < span; i += NT) {
        const long long g = i_lo + i;
        if (REAL) {
            if (g >= 0) cp_async4(&s_xr[i], xrow_r + g);
            else if (g >= -(long long)L) s_xr[i] = p.ring[(long long)(((long long)p.count + g + 4LL * L) % L) * Ctot + gch].x;
            else s_xr[i] = 0.f;
        } else {
            if (g >= 0) { if (!bulk || g >= gA + cnt2) cp_async8(&s_x[i], xrow + g); }
            else if (g >= -(long long)L) s_x[i] = p.ring[(long long)(((long long)p.count + g + 4LL * L) % L) * Ctot + gch];
            else s_x[i] = make_float2(0.f, 0.f);
        }
    }
    cp_async_commit(); cp_async_wait<0>();
    if (bulk && cnt2 > 0) { mbar_wait(&s_bar, parity); parity ^= 1u; }
    __syncthreads();
    if (HAS_NCO) {
        // the mixer runs in front of the filter: rotate the staged samples of this call in place (history samples
        // in the ring were rotated by the call that saw them)
        const uint32_t th0 = q.theta[gch], dth = q.dtheta[gch];
        const bool down = q.dir == 2;

        if (q.type == 0) {
            // table oscillator: four samples per thread and pass, so four table reads and four shared loads are in
            // flight together instead of one dependent pair per loop trip
            const int i0 = i_lo < 0 ? (int)(-i_lo) : 0;                  // first sample of this call in the span
            int i = i0 + tid;
            for (; i + 3 * NT < span; i += 4 * NT) {
                float2 sc[4], v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t th = th0 + (uint32_t)(i_lo + i + u * NT) * dth;
                    sc[u] = s_tab[nco_index(th)]; v[u] = s_x[i + u * NT];
                }
#pragma unroll
                for (int u = 0; u < 4; u++) s_x[i + u * NT] = down ? mix_down(v[u], sc[u]) : mix_up(v[u], sc[u]);
            }
            for (; i < span; i += NT) {
                const uint32_t th = th0 + (uint32_t)(i_lo + i) * dth;
                const float2 sc = s_tab[nco_index(th)];
                s_x[i] = down ? mix_down(s_x[i], sc) : mix_up(s_x[i], sc);
            }
        } else {
            for (int i = tid; i < span; i += NT) {
                const long long g = i_lo + i;
                if (g >= 0) { const float2 sc = nco_at(q, q.sincos, th0, dth, g); s_x[i] = down ? mix_down(s_x[i], sc) : mix_up(s_x[i], sc); }
            }
        }
        __syncthreads();
    }

    if (tid < nk) {
        const unsigned long long P = P0 + (unsigned long long)tid * p.step;
        const int off = (int)((long long)(P >> 24) - (L - 1) - i_lo);
        const unsigned f = (unsigned)((P & 0xffffffull) >> (24 - p.bits));
        const float *h = s_b + f * LP;
        // dotprod_cccf order and rounding: product rounded, then added, oldest sample first (scalar intrinsics:
        // ptxas contracts a packed mul.f32x2 + add.f32x2 pair into FFMA2 even with .rn)
        float ar = 0.f, ai = 0.f;
        if (VAR == 0) {
            for (int i = 0; i < L; i++) {
                const float2 w = s_x[off + i];
                ar = __fadd_rn(ar, __fmul_rn(h[i], w.x)); ai = __fadd_rn(ai, __fmul_rn(h[i], w.y));
            }
        } else if (VAR == 1) {
            for (int i = 0; i < L; i++) { const float2 w = s_x[off + i]; ar = __fmaf_rn(h[i], w.x, ar); ai = __fmaf_rn(h[i], w.y, ai); }
        } else {
            for (int i = 0; i < L; i++) ar = __fmaf_rn(h[i], s_xr[off + i], ar);
        }
        if (REAL) ((float *)yv)[ch * n_out + k0 + tid] = ar;
        else ((float2 *)yv)[ch * n_out + k0 + tid] = make_float2(ar, ai);
    }
    __syncthreads();                                    // the staged span is overwritten by the next work item
    }
}

// the newest min(n, sublen) inputs of the call go into the history ring
template <bool HAS_NCO, bool REAL>
__global__ void ring_update_kernel(const ResampP p, const NcoP q, const void *__restrict__ xv, int nch, int ch0, int Ctot, long long n)
{
    const int L = p.sublen;
    const long long first = n > L ? n - L : 0, cnt = n - first;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt * nch; i += (long long)gridDim.x * blockDim.x) {
        const long long j = first + i / nch, ch = i % nch;
        float2 v = REAL ? make_float2(((const float *)xv)[ch * n + j], 0.f) : ((const float2 *)xv)[ch * n + j];
        if (HAS_NCO) {
            const float2 sc = nco_at(q, q.sincos, q.theta[ch0 + ch], q.dtheta[ch0 + ch], j);
            v = q.dir == 2 ? mix_down(v, sc) : mix_up(v, sc);
        }
        p.ring[(long long)((p.count + j) % L) * Ctot + ch0 + ch] = v;
    }
}

// after a time-parallel pass over n samples every oscillator has advanced n steps
__global__ void nco_advance_kernel(const NcoP q, int nch, int ch0, long long n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nch) q.theta[ch0 + i] += (uint32_t)n * q.dtheta[ch0 + i];
}

// y[ch][k] = x[ch][k] * exp(+-j theta_k): nco_crcf_mix_block_up/down (nco.hpp:70,78) with the phase in closed form
__global__ void nco_par_kernel(const NcoP q, const float2 *__restrict__ x, float2 *__restrict__ y, int nch, int ch0, long long n)
{
    __shared__ float2 s_t[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_t[i] = q.sincos[i];
    __syncthreads();
    const long long total = (long long)nch * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long ch = i / n, k = i % n;
        const float2 sc = nco_at(q, s_t, q.theta[ch0 + ch], q.dtheta[ch0 + ch], k);
        const float2 v = x[i];
        y[i] = q.dir == 2 ? mix_down(v, sc) : mix_up(v, sc);
    }
}

// wdelay read-then-push over a block (Delay, utility.hpp:5-59): y[k] = x[k - D]; the D samples before the call come
// from hist_in (oldest first) and the newest D go to hist_out (ping-pong, so no CTA reads what another writes)
template <typename T>
__global__ void delay_kernel(const T *__restrict__ x, T *__restrict__ y, const T *__restrict__ hist_in, T *__restrict__ hist_out,
                             int nch, int ch0, long long n, long long D)
{
    const long long total = (long long)nch * (n + D);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long ch = i / (n + D), k = i % (n + D);
        const T *h = hist_in + (ch0 + ch) * D;
        if (k < n) {
            const long long s = k - D;
            y[ch * n + k] = s >= 0 ? x[ch * n + s] : h[D + s];
        } else {
            const long long t = n - D + (k - n);                 // time index of history slot k - n after this call
            hist_out[(ch0 + ch) * D + (k - n)] = t >= 0 ? x[ch * n + t] : h[D + t];
        }
    }
}

// bytes_to_iq over a whole block: interleaved int16 I/Q -> complex64 (reference utility.hpp:61-69)
__global__ void i16_to_c64_kernel(const unsigned *__restrict__ x, float2 *__restrict__ y, long long total)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        y[i] = i16_to_iq(x[i]);
}

}  // namespace

cudaError_t i16_to_c64_launch(const void *x, float2 *y, long long total, cudaStream_t stream)
{
    if (total <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 32);
    i16_to_c64_kernel<<<blocks, 256, 0, stream>>>((const unsigned *)x, y, total);
    return cudaGetLastError();
}

cudaError_t delay_launch(bool real, const void *x, void *y, const void *hist_in, void *hist_out, int nch, int ch0, long long n, long long D,
                         cudaStream_t stream)
{
    if (nch <= 0 || n <= 0) return cudaSuccess;
    const long long total = (long long)nch * (n + D);
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    if (real) delay_kernel<float><<<blocks, 256, 0, stream>>>((const float *)x, (float *)y, (const float *)hist_in, (float *)hist_out, nch, ch0, n, D);
    else      delay_kernel<float2><<<blocks, 256, 0, stream>>>((const float2 *)x, (float2 *)y, (const float2 *)hist_in, (float2 *)hist_out, nch, ch0, n, D);
    return cudaGetLastError();
}

cudaError_t nco_par_launch(const NcoP &q, const float2 *x, float2 *y, int nch, int ch0, long long n, cudaStream_t stream)
{
    if (nch <= 0 || n <= 0) return cudaSuccess;
    const long long total = (long long)nch * n;
    const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
    nco_par_kernel<<<blocks, 256, 0, stream>>>(q, x, y, nch, ch0, n);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return rc;
    nco_advance_kernel<<<(nch + 255) / 256, 256, 0, stream>>>(q, nch, ch0, n);
    return cudaGetLastError();
}

cudaError_t resamp_par_launch(const ResampP &p, const NcoP *nco, const void *x, void *y, int nch, int ch0, int Ctot,
                              long long n, long long n_out, cudaStream_t stream)
{
    if (nch <= 0 || n <= 0) return cudaSuccess;
    if (p.variant < 0 || p.variant > 2 || (p.variant == 2 && nco)) return cudaErrorInvalidValue;
    const NcoP q = nco ? *nco : NcoP{};
    if (n_out > 0) {
        // outputs per CTA: as many as keep the staged input span within ~48 KB
        int span_budget = 3000;          // staged samples per CTA (24 KB): measured best on config 3 and on CResampler
        if (const char *e = getenv("LQB_PAR_SPAN")) { const int v = atoi(e); if (v >= 256 && v <= 6000) span_budget = v; }
        long long kt = ((long long)(span_budget - p.sublen - 2) << 24) / p.step;
        const int KT = (int)(kt < 1 ? 1 : (kt > NT ? NT : kt));
        const int span_max = (int)((((unsigned long long)KT * p.step) >> 24) + p.sublen + 3) + 3;
        const long long ntiles = (n_out + KT - 1) / KT;
        if (ntiles * (long long)nch > 0x7fffffffLL) return cudaErrorInvalidValue;
        const size_t smem = (size_t)span_max * sizeof(float2) + (size_t)((p.npfb * (p.sublen | 1) + 3) & ~3) * sizeof(float) + (nco ? 1024 * sizeof(float2) : 0);
        if (smem > 200 * 1024) return cudaErrorInvalidValue;
        auto fn = p.variant == 2 ? resamp_par_kernel<false, 2>
                : p.variant == 1 ? (nco ? resamp_par_kernel<true, 1> : resamp_par_kernel<false, 1>)
                                 : (nco ? resamp_par_kernel<true, 0> : resamp_par_kernel<false, 0>);
        cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        // persistent CTAs: as many as are resident at once (shared memory bounds it), each walking the work items with that stride
        int per_sm = 1;
        rc = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fn, NT, smem);
        if (rc != cudaSuccess) return rc;
        int dev = 0, sms = 148;
        cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long nwork = ntiles * (long long)nch;
        const long long resident = (long long)std::max(1, per_sm) * sms;
        fn<<<(unsigned)std::min(nwork, resident), NT, smem, stream>>>(p, q, x, y, ch0, Ctot, n, n_out, KT, (int)ntiles, span_max, nwork);
        rc = cudaGetLastError();
        if (rc != cudaSuccess) return rc;
    }
    const long long work = (n < p.sublen ? n : p.sublen) * (long long)nch;
    const unsigned blocks = (unsigned)((work + 255) / 256 < 1184 ? (work + 255) / 256 : 1184);
    if (nco)                 ring_update_kernel<true, false><<<blocks, 256, 0, stream>>>(p, q, x, nch, ch0, Ctot, n);
    else if (p.variant == 2) ring_update_kernel<false, true><<<blocks, 256, 0, stream>>>(p, q, x, nch, ch0, Ctot, n);
    else                     ring_update_kernel<false, false><<<blocks, 256, 0, stream>>>(p, q, x, nch, ch0, Ctot, n);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess || !nco) return rc;
    nco_advance_kernel<<<(nch + 255) / 256, 256, 0, stream>>>(q, nch, ch0, n);
    return cudaGetLastError();
}

}  // namespace lqb
