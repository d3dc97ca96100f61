// am.h -- launcher of the decimated-rate AM tail kernel (am.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
constexpr int kAmBT = 64;     // channels per CTA
cudaError_t amtail_launch(bool has_agc, bool has_de, const AmTailArgs &a, cudaStream_t stream);
}  // namespace lqb
