// par.h -- launchers of the time-parallel kernels (par.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
// arbitrary-rate polyphase resampler, one thread per output sample; x [nch][n] -> y [nch][n_out]
cudaError_t resamp_par_launch(const ResampP &p, const float2 *x, float2 *y, int nch, int ch0, int Ctot,
                              long long n, long long n_out, cudaStream_t stream);
}  // namespace lqb
