// front.h -- launcher of the two-channels-per-thread front kernel (front.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
constexpr int kFront2BoxRows = 64;     // rows of the TMA box the kernel expects in SeqArgs::tmap
bool front2_supported(unsigned mask, int nsos);
cudaError_t front2_launch(int nsos, const SeqArgs &a, cudaStream_t stream, int ring_depth = 3);   // ring_depth 2: leaves shared memory for a co-resident tail CTA
}  // namespace lqb
