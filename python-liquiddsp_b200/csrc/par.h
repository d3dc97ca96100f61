// par.h -- launchers of the time-parallel kernels (par.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
// arbitrary-rate polyphase resampler, one thread per output sample; x [nch][n] -> y [nch][n_out]
// nco (optional) is the mixer in front of the resampler, applied while the input is staged
// (p.variant: cccf / crcf on complex64 rows, rrrf on float rows)
cudaError_t resamp_par_launch(const ResampP &p, const NcoP *nco, const void *x, void *y, int nch, int ch0, int Ctot,
                              long long n, long long n_out, cudaStream_t stream);
inline int resamp_par_launch_count(bool has_nco, long long n_out) { return (n_out > 0 ? 1 : 0) + 1 + (has_nco ? 1 : 0); }
// wdelay: y[k] = x[k - D] on float2 (or float) rows; hist_in / hist_out are [Ctot][D]
cudaError_t delay_launch(bool real, const void *x, void *y, const void *hist_in, void *hist_out, int nch, int ch0, long long n, long long D,
                         cudaStream_t stream);
// oscillator mix alone, phase in closed form (theta_0 + k * d_theta mod 2^32)
cudaError_t nco_par_launch(const NcoP &q, const float2 *x, float2 *y, int nch, int ch0, long long n, cudaStream_t stream);
// interleaved int16 I/Q pairs -> complex64, (float)s / 32767.0f exactly
cudaError_t i16_to_c64_launch(const void *x, float2 *y, long long total, cudaStream_t stream);
}  // namespace lqb
