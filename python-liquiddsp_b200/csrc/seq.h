// seq.h -- launcher of the channel-parallel sequential chain kernel (seq.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {

constexpr int kSeqBT  = 64;   // channels per CTA
constexpr int kSeqTS  = 16;   // samples per staged tile row (128 B of complex64 per channel)

// true when a kernel for this stage mask / section count was compiled
bool seq_supported(unsigned mask, int nsos);
// true when the kernel also exists with TMA (cp.async.bulk.tensor) input staging
bool seq_has_tma(unsigned mask, int nsos);
cudaError_t seq_launch(unsigned mask, int nsos, const SeqArgs &a, cudaStream_t stream);

}  // namespace lqb
