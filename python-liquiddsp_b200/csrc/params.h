// params.h -- plain-data kernel arguments shared by the host shim (capi.cu) and the kernels.
//
// Per-channel carried state lives in HBM as structure-of-arrays [slot][n_channels_total] so that a
// warp of 32 adjacent channels reads and writes 32 adjacent words.  History rings (resampler
// window, ampmodem windows, FIR history) are indexed by absolute sample count modulo the ring
// length; the host tracks the count, the kernels never move ring contents.
#pragma once
#include <stdint.h>
#include <cuda.h>          // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace lqb {

constexpr int kMaxSos   = 8;     // sections per launch (longer cascades run as several launches)
constexpr int kAmTaps   = 51;    // ampmodem: 2*m+1, m = 25
constexpr int kAmDelay  = 25;
constexpr int kAmRing   = 64;    // power-of-two ring >= kAmTaps
constexpr int kMaxResampSub = 64; // resampler sub-filter length limit for the fused path (2*m)

enum : unsigned {
    F_NCO = 1u, F_IIR = 2u, F_RS = 4u, F_AGC = 8u, F_AM = 16u, F_FM = 32u, F_DE = 64u, F_INREAL = 128u,
    F_INI16 = 256u,                // input rows are interleaved int16 I/Q (bytes_to_iq fused into the first stage)
    F_TF = 512u                    // IIR in transfer-function form (iirfilt_*_create(b, nb, a, na)); the kernel's section
                                   // count parameter then carries the (padded) number of delay elements
};

struct NcoP {
    uint32_t *theta, *dtheta;      // [Ctot]
    const float2 *sincos;          // device table [1024] of (sin, cos)
    int type;                      // 0 table NCO, 1 VCO (sinf/cosf)
    int dir;                       // 1 up, 2 down
};

struct IirP {
    int nsos;
    float b[kMaxSos][3];
    float a[kMaxSos][3];           // a[.][0] == 1 after normalisation
    float2 *v;                     // [nsos][2][Ctot] : v1, v2 per section (complex)
};

struct ResampP {
    uint32_t step, phase;          // 8.24 fixed point; phase at the first input of this call
    int bits, sublen, npfb;
    const float *bank;             // device [npfb][sublen], sub-filters reversed
    float2 *ring;                  // [sublen][Ctot] last sublen inputs, slot = count % sublen
    uint32_t count;                // inputs consumed since reset, modulo sublen
    int variant;                   // 0 resamp_cccf, 1 resamp_crcf, 2 resamp_rrrf (time-parallel kernel only)
};

struct AgcP {
    float alpha, scale, threshold;
    double one_minus_alpha;
    float chi, clo, cl2, chalf;    // single-precision gain loop (devmath.cuh agc_step_fast): chi + clo = 1 - alpha, cl2 = -alpha/2 ln 2, chalf = -alpha/2
    int fast;                      // unlocked and squelch disabled on every channel: the single-precision loop applies
    int big;                       // alpha > 0.0112: |alpha/2 ln y2'| can exceed 1/2, the single-precision loop keeps its ex2 branch
    int locked;
    unsigned timeout;
    float *g, *y2p;                // [Ctot]
    int *mode;                     // [Ctot] squelch state 0..7
    unsigned *timer;               // [Ctot]
    unsigned *rise_count;          // single counter (atomicAdd)
    const double2 *logtab;         // device [128] (inv, -ln inv) for the gain update's logarithm
};

struct AmP {
    float mod_index, pll_alpha, pll_beta;
    int suppressed;
    int out_v1;                    // USB / LSB with carrier: write the mixed-down delayed branch (complex) instead of audio
    float lp[kAmTaps];             // lowpass taps reversed: lp[i] multiplies the sample (50 - i) steps old
    float dc[kAmTaps];             // dc-block taps reversed
    const float2 *sincos;          // NCO table
    float2 *lp_ring;               // [kAmRing][Ctot]
    float *dc_ring;                // [kAmRing][Ctot]
    uint32_t *theta, *dtheta;      // [Ctot]
    uint32_t count;                // samples consumed since reset, modulo kAmRing
};

struct FmP { float ref; float2 *rprime; };          // [Ctot]
struct DeP { float b0, a1; float *v1; };            // [Ctot]
constexpr int kMaxTf = 16;         // coefficients per polynomial in transfer-function form
struct TfP {
    float b[kMaxTf], na[kMaxTf];   // b[i] / a0 and -(a[i] / a0)
    int nb, nna;                   // lengths of b and a
    float2 *v;                     // [kMaxTf - 1][Ctot] delay line v[1..]
};

struct alignas(64) SeqArgs {
    CUtensorMap tmap_out;          // TMA mode, full-rate complex output: [C rows][2 out_pitch floats], same box and swizzle
    CUtensorMap tmap;              // TMA mode: the input as a 2-D tensor [C rows][2n floats], box 32 floats x 32 rows, 128B swizzle
    const void *x;                 // [C][n] input rows (complex64, or float32 when F_INREAL)
    void *y;                       // [C][out_pitch] output rows
    int tma_out;                   // the output tile leaves through tmap_out (one bulk tensor store per warp and tile)
    int C;                         // channels in this launch
    int ch0;                       // first channel's index into the [.. ][Ctot] state arrays
    int Ctot;                      // channel stride of the state arrays
    int vec_in, vec_out;           // 1 when rows allow 16-byte vector access
    int cpw;                       // channels per warp: 32, or 16 / 8 to spread few channels over more warps
    int use_tma;                   // tmap is valid: stage the input with TMA (full-warp kernels that have the variant)
    int out_tmajor;                // decimated output stored [sample][channel] (hand-off to the AM tail kernel)
    const void *tapstream;         // lane-split front kernels (lanes.cu): per-tile tap records of this call
    int lanes_flags;               // lanes.cu tuning switches (bit 0: no whole-stage bodies)
    alignas(16) float lc[20];      // lanes.cu, one lane pair per channel: -a1[4], -a2[4], b1[4], b0[4], b2[4] of the cascade, in the
                                   // order a step uses them (five 16-byte uniform loads per tile instead of eleven scattered ones)
    long long n, out_pitch;
    NcoP nco; IirP iir; ResampP rs; AgcP agc; AmP am; FmP fm; DeP de; TfP tf;
};

struct AmTailArgs {
    const float2 *x;               // input, row-major [C][in_pitch] or time-major [n][in_pitch]
    float *y;                      // [C][out_pitch]
    int C, ch0, Ctot, in_tmajor;
    long long n, in_pitch, out_pitch;
    AgcP agc; AmP am; DeP de;
};

// BroadcastAM (demod.hpp:94-153): carrier PLL with an arg() phase detector, Kaiser lowpass of 2m+1 taps on the PLL
// branch, the signal branch delayed by m, real part high-passed by a two-section IIR
constexpr int kBamMaxM = 64;
struct BamP {
    int m, ntaps_pad;              // lowpass taps padded (leading zeros) to a multiple of 8
    const double *atantab;         // [65][8] table of atan2_rn
    const float *hrev;             // [ntaps_pad] taps in window order: hrev[i] multiplies the sample (ntaps_pad-1-i) steps old
    float pll_alpha, pll_beta;
    float b[2][3], a[2][3];        // DC-block sections (iirfilt_rrrf, SOS)
    const float2 *sincos;
    float2 *hist;                  // [ntaps_pad - 1][Ctot] newest inputs of the previous calls, oldest first
    float *dcv;                    // [4][Ctot] v1, v2 of each section
    uint32_t *theta, *dtheta;      // [Ctot]
};
struct BamArgs {
    const float2 *x;               // row-major [C][in_pitch] or time-major [n][in_pitch]
    float *y;                      // [C][out_pitch]
    int C, ch0, Ctot, in_tmajor, has_de;
    long long n, in_pitch, out_pitch;
    BamP p; DeP de;
};

// FMStereo (demod.hpp:4-85): discriminator, pilot mixer with its PLL, two one-pole de-emphasis filters, two real
// resamplers with a common phase
constexpr int kFmstMaxSub = 32;    // taps per sub-filter of the audio resamplers (create_default: 14)
struct FmstP {
    float ref;                     // freqdem: 1 / (2 pi kf)
    float pll_alpha, pll_beta;
    float b0, a1;                  // de-emphasis: v0 = x - a1 v1, y = b0 v0
    const double *atantab;         // [65][8] table of atan2_rn
    const float2 *sincos;
    ResampP rs;                    // step, phase, bits, sublen, npfb, bank, count (ring unused)
    float2 *rprime;                // [Ctot] previous input sample
    uint32_t *theta, *dtheta;      // [Ctot]
    float *pe, *vL, *vR;           // [Ctot] filtered phase error, de-emphasis states
    float *ringL, *ringR;          // [sublen][Ctot] resampler windows, slot = count % sublen
};
struct FmstArgs {
    const float2 *x;               // [C][n]
    float *y;                      // [C][2 * n_out] interleaved left, right
    int C, ch0, Ctot, cpw;         // cpw: channels per warp (8, 16 or 32)
    long long n, n_out;
    FmstP p;
};

enum { FIR_PLAIN = 0, FIR_SSB_LSB, FIR_SSB_USB, FIR_R2C, FIR_C2R };
constexpr int kFirUTaps = 128;
struct FirArgs {
    const float2 *x; float2 *y;    // [C][n]
    int C, ch0, Ctot, ntaps;
    int real_io;                   // firfilt_rrrf: x and y are float rows
    int pair;                      // firfilt_rrrf: two real channels per logical row (needs an even ch0)
    // firhilbf built on the same kernel: separate taps for the two lanes and a combining epilogue
    int in_real, out_real;         // element types when they differ (real_io sets both)
    int mode;                      // FIR_PLAIN, FIR_SSB_LSB / FIR_SSB_USB (re + / - im), FIR_R2C, FIR_C2R
    int delay;                     // firhilbf semi-length m (sign toggles and the r2c end-of-call rule)
    unsigned long long count;      // absolute index of this call's first sample since reset
    long long zero_at[4];          // FIR_R2C: local output indices whose in-phase part is forced to zero (-1 = none)
    const float *taps_q;           // taps of the imaginary lane (nullptr: same as taps)
    float post_div;                // SSB modes inside ampmodem: y = (0.5 * side-band) / post_div; 0 = off
    int skip, skip_keep;           // 1 / 2: taps at even / odd positions are zero in both lanes, except position skip_keep (-1: none)
    long long n;
    float scale;
    const float *taps;             // device [ntaps] in design order h[0..ntaps-1]
    // The same taps by value (zero-padded) when both lanes share them and there are at most kFirUTaps: they then reach the
    // multiply-adds as warp-uniform operands (FFMA2 R, R, UR.F32, R: two register pairs per instruction).  A tap that sits
    // in a per-thread register pair makes three, and the register file delivers those in 2.7 cycles instead of 2
    // (tools/ubench_rf.cu) -- that, not the FP32 pipe, held the 64-tap filter at 63 % of its ceiling.
    int utap;
    float taps_c[128];
    const float2 *hist_in;         // [Ctot][ntaps-1] last inputs before this call, oldest first
    float2 *hist_out;              // written by the call (ping-pong with hist_in)
};

}  // namespace lqb
