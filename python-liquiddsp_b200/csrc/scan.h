// scan.h -- time-parallel (blocked linear-recurrence scan) evaluation of the biquad cascade (scan.cu).
#pragma once
#include <cuda_runtime.h>
#include "params.h"

namespace lqb {

struct IirScanArgs {
    float2 *y;                 // [C][n]: holds the zero-state block responses on entry, the filter output on exit
    int C, ch0, Ctot, nsos, B; // B = block length, K = n / B blocks per channel
    long long n;
    const double *H;           // [B][S]   response at step k of the block to unit initial state j   (S = 2 * nsos)
    const double *M;           // [S][S]   state after B steps from unit initial state j (column j)
    const float2 *vblk;        // [nsos][2][C*K] end state of every block's zero-state run (seq kernel state layout)
    float2 *v;                 // [nsos][2][Ctot] the channels' carried state
    double2 *sin;              // [C*K][S] scratch: state entering each block
};

cudaError_t iir_scan_launch(const IirScanArgs &a, cudaStream_t stream);

}  // namespace lqb
