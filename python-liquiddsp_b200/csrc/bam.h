// bam.h -- launcher of the BroadcastAM kernel (bam.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
cudaError_t bam_launch(const BamArgs &a, cudaStream_t stream);
// AGC alone, in place on a time-major [sample][channel] block (am.cu)
cudaError_t agc_tmajor_launch(const AmTailArgs &a, cudaStream_t stream);
}  // namespace lqb
