// fir.h -- launcher of the time-parallel firfilt_crcf kernel (fir.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
constexpr int kFirThreads = 128;
constexpr int kFirMaxTaps = 1024;
cudaError_t fir_launch(const FirArgs &a, cudaStream_t stream);
}  // namespace lqb
