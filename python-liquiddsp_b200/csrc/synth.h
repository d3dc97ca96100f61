// synth.h -- device-side generators of the benchmark inputs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lqb {
cudaError_t synth_launch(int kind, float2 *x, int n_channels, int channel0, long long n, unsigned long long n0,
                         unsigned long long seed, cudaStream_t stream);
}
