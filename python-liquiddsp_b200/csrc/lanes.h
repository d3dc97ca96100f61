// lanes.h -- launchers of the lane-split front kernels (lanes.cu): biquad cascade -> decimating resampler with the real
// and the imaginary part of a channel on two lanes (and, for few channels, the cascade's sections on further lanes).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {
// lanes per channel the launch would use for `nch` channels: 2 (re | im), or 4 / 8 (the cascade's sections spread over
// 2 / 4 lane pairs as a systolic pipeline) when few channels must fill the machine.  0: no kernel for this chain.
int lanes_per_channel(unsigned mask, int nsos, long long nch);
// rows of the TMA box the kernel expects in SeqArgs::tmap for that lane count (= channels per warp)
inline int lanes_box_rows(int lanes) { return 32 / lanes; }
// the per-call tap stream (one 144-byte record per 16-sample tile): bytes needed, and the kernel that fills it
// (records are in step space: the cascade's last section runs lanes_lag() steps behind the input)
int lanes_lag(int nsos, int lanes);
size_t lanes_tapstream_bytes(long long n);
cudaError_t lanes_tapstream_launch(const ResampP &rs, long long n, int lag, void *buf, cudaStream_t stream);
// a.tmap: [lanes_box_rows x 128 B] boxes, a.tapstream: filled by lanes_tapstream_launch for this call's phase and length
cudaError_t lanes_launch(int nsos, int lanes, const SeqArgs &a, cudaStream_t stream);
const char *lanes_kernel_name(int nsos, int lanes);
}  // namespace lqb
