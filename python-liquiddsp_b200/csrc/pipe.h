// pipe.h -- launcher of the three-warp pipeline kernel for cascade -> gain loop -> discriminator (pipe.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
bool pipe_supported(unsigned mask, int nsos);
cudaError_t pipe_launch(int nsos, const SeqArgs &a, cudaStream_t stream);     // needs a.agc.fast (single-precision gain loop) and a.use_tma (a.tmap: [32 x 128 B] boxes)
}  // namespace lqb
