// fmst.h -- launcher of the FMStereo kernel (fmst.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
cudaError_t fmstereo_launch(const FmstArgs &a, cudaStream_t stream);
}  // namespace lqb
