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

namespace lqb {
namespace {

constexpr int NT = 128;

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(NT)
resamp_par_kernel(const ResampP p, const float2 *__restrict__ x, float2 *__restrict__ y, int ch0, int Ctot,
                  long long n, long long n_out, int KT, int ntiles, int span_max)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2 *s_x = (float2 *)smem_raw;                   // span_max samples
    float  *s_b = (float *)(s_x + span_max);            // bank [npfb][sublen]
    const int tid = threadIdx.x, L = p.sublen;
    const long long ch = blockIdx.x / ntiles, tile = blockIdx.x % ntiles;
    const long long k0 = tile * KT;
    const int nk = (int)((n_out - k0) < KT ? (n_out - k0) : KT);
    const unsigned long long P0 = (unsigned long long)p.phase + (unsigned long long)k0 * p.step;
    const unsigned long long P1 = (unsigned long long)p.phase + (unsigned long long)(k0 + nk - 1) * p.step;
    const long long i_lo = (long long)(P0 >> 24) - (L - 1), i_hi = (long long)(P1 >> 24);
    const int span = (int)(i_hi - i_lo + 1);
    const float2 *xrow = x + ch * n;
    const long long gch = ch0 + ch;

    for (int i = tid; i < p.npfb * L; i += NT) s_b[i] = p.bank[i];
    for (int i = tid; i < span; i += NT) {
        const long long g = i_lo + i;
        if (g >= 0) cp_async8(&s_x[i], xrow + g);
        else if (g >= -(long long)L) s_x[i] = p.ring[(long long)(((long long)p.count + g + 4LL * L) % L) * Ctot + gch];
        else s_x[i] = make_float2(0.f, 0.f);
    }
    cp_async_commit(); cp_async_wait<0>();
    __syncthreads();

    if (tid < nk) {
        const unsigned long long P = P0 + (unsigned long long)tid * p.step;
        const int off = (int)((long long)(P >> 24) - (L - 1) - i_lo);
        const unsigned f = (unsigned)((P & 0xffffffull) >> (24 - p.bits));
        const float *h = s_b + f * L;
        // dotprod_cccf order and rounding: product rounded, then added, oldest sample first (scalar intrinsics:
        // ptxas contracts a packed mul.f32x2 + add.f32x2 pair into FFMA2 even with .rn)
        float ar = 0.f, ai = 0.f;
        for (int i = 0; i < L; i++) {
            const float2 w = s_x[off + i];
            ar = __fadd_rn(ar, __fmul_rn(h[i], w.x)); ai = __fadd_rn(ai, __fmul_rn(h[i], w.y));
        }
        y[ch * n_out + k0 + tid] = make_float2(ar, ai);
    }
}

// the newest min(n, sublen) inputs of the call go into the history ring
__global__ void ring_update_kernel(const ResampP p, const float2 *__restrict__ x, int nch, int ch0, int Ctot, long long n)
{
    const int L = p.sublen;
    const long long first = n > L ? n - L : 0, cnt = n - first;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cnt * nch; i += (long long)gridDim.x * blockDim.x) {
        const long long j = first + i / nch, ch = i % nch;
        p.ring[(long long)((p.count + j) % L) * Ctot + ch0 + ch] = x[ch * n + j];
    }
}

}  // namespace

cudaError_t resamp_par_launch(const ResampP &p, const float2 *x, float2 *y, int nch, int ch0, int Ctot,
                              long long n, long long n_out, cudaStream_t stream)
{
    if (nch <= 0 || n <= 0) return cudaSuccess;
    if (n_out > 0) {
        // outputs per CTA: as many as keep the staged input span within ~48 KB
        const int span_budget = 6000;
        long long kt = ((long long)(span_budget - p.sublen - 2) << 24) / p.step;
        const int KT = (int)(kt < 1 ? 1 : (kt > NT ? NT : kt));
        const int span_max = (int)((((unsigned long long)KT * p.step) >> 24) + p.sublen + 3);
        const long long ntiles = (n_out + KT - 1) / KT;
        if (ntiles * (long long)nch > 0x7fffffffLL) return cudaErrorInvalidValue;
        const size_t smem = (size_t)span_max * sizeof(float2) + (size_t)p.npfb * p.sublen * sizeof(float);
        if (smem > 200 * 1024) return cudaErrorInvalidValue;
        cudaError_t rc = cudaFuncSetAttribute((const void *)resamp_par_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (rc != cudaSuccess) return rc;
        resamp_par_kernel<<<(unsigned)(ntiles * nch), NT, smem, stream>>>(p, x, y, ch0, Ctot, n, n_out, KT, (int)ntiles, span_max);
        rc = cudaGetLastError();
        if (rc != cudaSuccess) return rc;
    }
    const long long work = (n < p.sublen ? n : p.sublen) * (long long)nch;
    const unsigned blocks = (unsigned)((work + 255) / 256 < 1184 ? (work + 255) / 256 : 1184);
    ring_update_kernel<<<blocks, 256, 0, stream>>>(p, x, nch, ch0, Ctot, n);
    return cudaGetLastError();
}

}  // namespace lqb
