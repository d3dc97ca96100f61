// pipe.h -- launcher of the two-warp pipeline kernel for cascade -> gain loop -> discriminator (pipe.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
bool pipe_supported(unsigned mask, int nsos);
cudaError_t pipe_launch(int nsos, const SeqArgs &a, cudaStream_t stream);     // needs a.agc.fast (the single-precision gain loop)
}  // namespace lqb
