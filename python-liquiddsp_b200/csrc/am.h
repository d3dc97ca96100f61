// am.h -- launcher of the decimated-rate AM tail kernel (am.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
constexpr int kAmBT = 64;     // channels per CTA
// returns the launch status; *launches (optional) receives the number of kernels launched (1 or 2)
cudaError_t amtail_launch(bool has_agc, bool has_de, const AmTailArgs &a, cudaStream_t stream);
inline int amtail_launch_count(bool has_agc, const AmTailArgs &a) { return (has_agc && a.in_tmajor) ? 2 : 1; }
// true when the launch takes the eight-lanes-per-channel kernel (few channels, DSB with carrier)
bool amtail_few(bool has_agc, const AmTailArgs &a);
// name of the demodulator kernel the launch takes (amtail_kernel or amtail8_kernel)
const char *amtail_kernel_name(bool has_agc, const AmTailArgs &a);
}  // namespace lqb
