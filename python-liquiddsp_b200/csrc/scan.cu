// scan.cu -- biquad cascade as a blocked linear-recurrence scan: each channel parallelises along time.
//
// iirfilt_crcf_execute_block (reference call iirfilter.hpp:296) is a linear recurrence of order 2*nsos, so a block
// of B samples can be filtered from a zero state independently of every other block (pass 1: the sequential kernel
// of seq.cu run over the [C*K][B] view of the data, one thread per block), the true state entering each block then
// follows from a short prefix pass over the K blocks of a channel (pass 2, below: s <- M s + e_k), and the missing
// homogeneous response is added back per sample (pass 3: y[k] += sum_j H[k][j] s_j).
//
// Rounding: reordering an IIR changes its rounding noise -- for the README filter (poles at radius 0.995, internal
// state gain ~1e3) a naive fp32 scan is 4.5e-4 from the sequential result.  Passes 2 and 3 therefore run in double
// (M, H and the carried state are double; only the block-local pass is fp32), which brings the scan to ~5e-5 of
// the sequential fp32 result -- the sequential filter's own distance from an fp64 evaluation.  The
// scan is therefore opt-in (lqb_iirfilt_crcf_set_mode(2)): the default path is the sequential kernel, which
// matches the oracle bit for bit.
#include <cuda_runtime.h>
#include "scan.h"

namespace lqb {
namespace {

constexpr int SMAX = 2 * kMaxSos;

// pass 2: one thread per channel, K steps of an S x S real matrix on an S-vector of complex states
__global__ void iir_scan_prefix_kernel(const IirScanArgs a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.C) return;
    const int S = 2 * a.nsos, K = (int)(a.n / a.B);
    const long long CK = (long long)a.C * K, gch = a.ch0 + c;
    double2 cur[SMAX], nxt[SMAX];
    for (int j = 0; j < S; j++) { const float2 v = a.v[(long long)j * a.Ctot + gch]; cur[j] = make_double2(v.x, v.y); }
    for (int k = 0; k < K; k++) {
        const long long r = (long long)c * K + k;
        for (int j = 0; j < S; j++) a.sin[r * S + j] = cur[j];
        for (int i = 0; i < S; i++) {
            const float2 e = a.vblk[(long long)i * CK + r];
            double2 s = make_double2(e.x, e.y);
            for (int j = 0; j < S; j++) { const double m = a.M[i * S + j]; s.x = fma(m, cur[j].x, s.x); s.y = fma(m, cur[j].y, s.y); }
            nxt[i] = s;
        }
        for (int j = 0; j < S; j++) cur[j] = nxt[j];
    }
    for (int j = 0; j < S; j++) a.v[(long long)j * a.Ctot + gch] = make_float2((float)cur[j].x, (float)cur[j].y);
}

// pass 3: one CTA per block, one thread per sample
__global__ void iir_scan_fix_kernel(const IirScanArgs a)
{
    __shared__ double2 s_s[SMAX];
    const int S = 2 * a.nsos;
    const long long r = blockIdx.x;                       // block index c*K + k == row of the [C*K][B] view
    if (threadIdx.x < S) s_s[threadIdx.x] = a.sin[r * S + threadIdx.x];
    __syncthreads();
    for (int k = threadIdx.x; k < a.B; k += blockDim.x) {
        const double *h = a.H + (long long)k * S;
        double cr = 0.0, ci = 0.0;
        for (int j = 0; j < S; j++) { cr = fma(h[j], s_s[j].x, cr); ci = fma(h[j], s_s[j].y, ci); }
        float2 *p = a.y + r * a.B + k;
        const float2 v = *p;
        *p = make_float2((float)((double)v.x + cr), (float)((double)v.y + ci));
    }
}

}  // namespace

cudaError_t iir_scan_launch(const IirScanArgs &a, cudaStream_t stream)
{
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    if (a.nsos > kMaxSos || a.n % a.B) return cudaErrorInvalidValue;
    const long long rows = (long long)a.C * (a.n / a.B);
    if (rows > 0x7fffffffLL) return cudaErrorInvalidValue;
    iir_scan_prefix_kernel<<<(a.C + 127) / 128, 128, 0, stream>>>(a);
    cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) return rc;
    iir_scan_fix_kernel<<<(unsigned)rows, a.B < 256 ? a.B : 256, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
