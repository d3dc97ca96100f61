/*
 * liquiddsp_b200.h -- C ABI of the B200-native streaming baseband chain.
 *
 * This is the drop-in boundary for the hot path of colbyAtCRI/python-liquiddsp.  The reference
 * has no C ABI of its own: its boundary is the pybind11 class surface of src/wrapper.cpp, each
 * class owning one liquid-dsp handle and calling liquid's create / destroy / reset /
 * execute_block C functions.  Each family below replaces exactly that pairing and cites the
 * reference wrapper (file:line under /root/reference/src) whose liquid calls it stands in for.
 *
 * Conventions
 *   - every function returns an int status: LQB_OK (0) on success, a negative LQB_E* code
 *     otherwise; nothing throws across the ABI.  lqb_last_error() gives the text.
 *   - every object is batched: n_channels independent channels, each with its own carried
 *     state, all sharing the object's parameters.  n_channels = 1 is the reference's object.
 *   - sample buffers are dense row-major [n_channels x n] ; complex samples are interleaved
 *     (re, im) float32 pairs == numpy complex64 == liquid_float_complex.
 *   - *_execute()      : HOST pointers; the call does H2D, kernels, D2H and returns when done.
 *     *_execute_dev()  : DEVICE pointers (16-byte aligned) + a cudaStream_t passed as void*;
 *                        asynchronous, state stays resident in HBM between calls.
 *   - there is no CPU fallback: without a CUDA device every execute returns LQB_ECUDA.
 */
#ifndef LIQUIDDSP_B200_H
#define LIQUIDDSP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQB_OK        0
#define LQB_EINVAL   -1   /* bad argument / unsupported configuration                 */
#define LQB_ECUDA    -2   /* CUDA runtime error (no device, launch failure, ...)       */
#define LQB_ENOMEM   -3
#define LQB_ESIZE    -4   /* output capacity too small                                 */
#define LQB_ENOTIMPL -5   /* reference feature outside the built scope (named in text) */

typedef struct { float re, im; } lqb_cf;          /* complex64 */
typedef struct lqb_stage_s *lqb_stage;            /* any stage object below */
typedef struct lqb_chain_s *lqb_chain;

/* enum values follow liquid.h */
enum { LQB_IIRDES_BUTTER = 0, LQB_IIRDES_CHEBY1, LQB_IIRDES_CHEBY2, LQB_IIRDES_ELLIP, LQB_IIRDES_BESSEL };
enum { LQB_IIRDES_LOWPASS = 0, LQB_IIRDES_HIGHPASS, LQB_IIRDES_BANDPASS, LQB_IIRDES_BANDSTOP };
enum { LQB_NCO = 0, LQB_VCO };
enum { LQB_AMPMODEM_DSB = 0, LQB_AMPMODEM_USB, LQB_AMPMODEM_LSB };
enum { LQB_MIX_UP = 1, LQB_MIX_DOWN = 2 };

/* ---------------- library ---------------- */
int         lqb_version(void);
const char *lqb_last_error(void);                      /* thread-local text of the last failure */
int         lqb_device_count(int *count);
int         lqb_set_device(int device);                /* objects are created on the current device */
int         lqb_device_synchronize(void);
/* pinned host memory and raw device memory, so callers without another CUDA binding can stage data */
int         lqb_host_alloc(void **p, size_t bytes);
int         lqb_host_free(void *p);
int         lqb_dev_alloc(void **p, size_t bytes);
int         lqb_dev_free(void *p);
int         lqb_memcpy_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int         lqb_memcpy_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int         lqb_stream_synchronize(void *stream);

/* ---------------- generic stage operations (valid for every stage handle) ---------------- */
int lqb_stage_destroy(lqb_stage s);
int lqb_stage_reset(lqb_stage s);                       /* liquid *_reset(): clears carried state */
int lqb_stage_channels(lqb_stage s, int *n_channels);
/* output length for n input samples per channel, given the current carried state */
int lqb_stage_out_len(lqb_stage s, size_t n, size_t *n_out);
/* Host-pointer execute. x: [n_channels x n] of the stage's input type; y: [n_channels x *n_out].
 * y_capacity is in samples per channel. */
int lqb_stage_execute(lqb_stage s, const void *x, size_t n, void *y, size_t y_capacity, size_t *n_out);
int lqb_stage_execute_dev(lqb_stage s, const void *x_dev, size_t n, void *y_dev, size_t y_capacity,
                          size_t *n_out, void *stream);

/* ---------------- iirfilt_crcf : ComplexIIRFilter, iirfilter.hpp:244-299 ----------------
 * replaces iirfilt_crcf_create_prototype (:275), iirfilt_crcf_execute_block (:296),
 * iirfilt_crcf_freqresponse (:288), iirfilt_crcf_destroy (:279).  complex in -> complex out. */
int lqb_iirfilt_crcf_create_prototype(int ftype, int btype, int order, float fc, float f0,
                                      float ap, float as, int n_channels, lqb_stage *out);
int lqb_iirfilt_crcf_create_sos(const float *B, const float *A, int nsos, int n_channels, lqb_stage *out);
int lqb_iirfilt_crcf_get_sos(lqb_stage s, float *B, float *A, int *nsos);   /* 3 floats per section */
int lqb_iirfilt_crcf_freqresponse(lqb_stage s, float fc, lqb_cf *H);
/* 0 = auto, 1 = channel-parallel sequential (bit-matches the oracle), 2 = time-parallel blocked scan */
int lqb_iirfilt_crcf_set_mode(lqb_stage s, int mode);

/* ---------------- iirfilt_rrrf (SOS) : RealIIRFilter iirfilter.hpp:301-356, RLowpassIIR.. :171-241 ------
 * replaces iirfilt_rrrf_create_prototype (:180,198,216,234,332) and iirfilt_rrrf_execute_block.  real -> real.
 * The C*IIR classes (:61-131) are lqb_iirfilt_crcf_create_prototype with the band type fixed. */
int lqb_iirfilt_rrrf_create_prototype(int ftype, int btype, int order, float fc, float f0,
                                      float ap, float as, int n_channels, lqb_stage *out);
int lqb_iirfilt_rrrf_create_sos(const float *B, const float *A, int nsos, int n_channels, lqb_stage *out);

/* ---------------- iirfilt, transfer-function form : CIIRFilter iirfilter.hpp:23-58, RIIRFilter :133-168 ------
 * replaces iirfilt_crcf_create(b, nb, a, na) (:33) / iirfilt_rrrf_create (:143), _execute_block (:55,:165),
 * _freqresponse (:48,:158), _reset (:43,:153).  Up to 16 coefficients per polynomial; a[0] normalises. */
int lqb_iirfilt_crcf_create(const float *b, int nb, const float *a, int na, int n_channels, lqb_stage *out);
int lqb_iirfilt_rrrf_create(const float *b, int nb, const float *a, int na, int n_channels, lqb_stage *out);
int lqb_iirfilt_tf_freqresponse(lqb_stage s, float fc, lqb_cf *H);

/* ---------------- iirfilt_rrrf one-pole : DeemphasisFilter, iirfilter.hpp:358-392 ----------
 * replaces iirfilt_rrrf_create(b,1,a,2) (:371) and the per-sample iirfilt_rrrf_execute loop
 * (:388-389); coefficient formula of :366-370.  real in -> real out. */
int lqb_deemph_create(float sample_rate, int n_channels, lqb_stage *out);
int lqb_deemph_get_coeffs(lqb_stage s, float *b0, float *a1);
int lqb_deemph_freqresponse(lqb_stage s, float fc, lqb_cf *H);

/* ---------------- firfilt_crcf : new FIRFilter class (BASELINE config 2) -------------------
 * modelled on RealFIRFilter, firfilter.hpp:5-36 (firfilt_rrrf_create :15, _execute_block :33,
 * _freqresponse :27); the crcf form is what demod.hpp:105,135-136 uses.  complex -> complex. */
int lqb_firfilt_crcf_create(const float *h, int h_len, int n_channels, lqb_stage *out);
int lqb_firfilt_crcf_create_kaiser(int h_len, float fc, float as, float mu, int n_channels, lqb_stage *out);
/* firfilt_rrrf : RealFIRFilter firfilter.hpp:5-36, RealDCBlocker :39-50 (create_dc_blocker), RealKaiserBessel
 * :52-67 (create_kaiser, then set_scale(1/|H(0)|)).  real -> real; set_scale / get_taps / freqresponse below apply. */
int lqb_firfilt_rrrf_create(const float *h, int h_len, int n_channels, lqb_stage *out);
int lqb_firfilt_rrrf_create_kaiser(int h_len, float fc, float as, float mu, int n_channels, lqb_stage *out);
int lqb_firfilt_rrrf_create_dc_blocker(int m, float as, int n_channels, lqb_stage *out);
int lqb_firfilt_crcf_set_scale(lqb_stage s, float scale);
int lqb_firfilt_crcf_get_taps(lqb_stage s, float *h, int *h_len);
int lqb_firfilt_crcf_freqresponse(lqb_stage s, float fc, lqb_cf *H);

/* ---------------- resamp_cccf : ComplexResampler, resampler.hpp:127-173 ---------------------
 * replaces resamp_cccf_create (:136), _reset (:144), _set_rate (:153) and the per-sample
 * resamp_cccf_execute loop (:164-166).  complex in -> complex out, *n_out samples per channel
 * (all channels share rate and phase, so the batched output is rectangular). */
int lqb_resamp_create(float rate, int m, float fc, float as, int npfb, int n_channels, lqb_stage *out);
/* resamp_crcf / resamp_rrrf : CResampler resampler.hpp:38-70, RResampler :4-36 (create_default + execute_block),
 * RealResampler :72-125 (resamp_rrrf_create, per-sample execute loop :104-107).  Real taps: fused multiply-add chain.
 * create_default: m = 7, fc = min(0.49, rate/2), As = 60 dB, npfb = 64.  set_rate / get_state / get_bank apply. */
int lqb_resamp_crcf_create(float rate, int m, float fc, float as, int npfb, int n_channels, lqb_stage *out);
int lqb_resamp_rrrf_create(float rate, int m, float fc, float as, int npfb, int n_channels, lqb_stage *out);
int lqb_resamp_crcf_create_default(float rate, int n_channels, lqb_stage *out);
int lqb_resamp_rrrf_create_default(float rate, int n_channels, lqb_stage *out);
/* wdelayf / wdelaycf : Delay utility.hpp:5-59 (read, then push: y[k] = x[k - (delay + 1)]).  real_samples selects
 * float rows (wdelayf) or complex64 rows (wdelaycf). */
int lqb_wdelay_create(int delay, int real_samples, int n_channels, lqb_stage *out);
int lqb_resamp_set_rate(lqb_stage s, float rate);
int lqb_resamp_get_state(lqb_stage s, uint32_t *step, uint32_t *phase);     /* 8.24 fixed point */
int lqb_resamp_get_bank(lqb_stage s, float *bank, int *npfb, int *sublen);  /* [npfb][sublen] as firpfb stores it */

/* ---------------- nco_crcf : NCO, nco.hpp:10-81 ---------------------------------------------
 * replaces nco_crcf_create (:18,22), set/adjust frequency and phase (:34-52), pll (:54-60) and
 * nco_crcf_mix_block_up/down (:70,78).  Phase/frequency are uint32 and bit-exact.  Per-channel
 * frequency/phase arrays are an extension for the batched form. */
int lqb_nco_create(int type, int n_channels, lqb_stage *out);
int lqb_nco_set_direction(lqb_stage s, int dir);                 /* LQB_MIX_UP (default, __call__) / LQB_MIX_DOWN */
int lqb_nco_set_frequency(lqb_stage s, float f);                 /* radians/sample, all channels */
int lqb_nco_adjust_frequency(lqb_stage s, float df);
int lqb_nco_set_phase(lqb_stage s, float phi);
int lqb_nco_adjust_phase(lqb_stage s, float dphi);
int lqb_nco_get_frequency(lqb_stage s, float *f);                /* channel 0 */
int lqb_nco_get_phase(lqb_stage s, float *phi);
int lqb_nco_pll_set_bandwidth(lqb_stage s, float bw);
int lqb_nco_pll_step(lqb_stage s, float dphi);
int lqb_nco_set_frequency_per_channel(lqb_stage s, const float *f, int n);
int lqb_nco_get_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n);   /* per channel */
int lqb_nco_set_u32(lqb_stage s, const uint32_t *theta, const uint32_t *d_theta, int n);

/* ---------------- agc_crcf : AGC, agc.hpp:4-128 ---------------------------------------------
 * replaces agc_crcf_create (:10) and every property call (:17-107), plus the wrapper's own loop
 * (:109-128): per-sample agc_crcf_execute, squelch status poll, zeroing in ENABLED/SIGNALLO and
 * the count of RISE transitions (the host then calls onRise that many times). */
int lqb_agc_create(int n_channels, lqb_stage *out);
int lqb_agc_set_bandwidth(lqb_stage s, float bw);
int lqb_agc_get_bandwidth(lqb_stage s, float *bw);
int lqb_agc_set_signal_level(lqb_stage s, float level);
int lqb_agc_get_signal_level(lqb_stage s, float *level);         /* channel 0 */
int lqb_agc_set_rssi(lqb_stage s, float rssi_dB);
int lqb_agc_get_rssi(lqb_stage s, float *rssi_dB);               /* channel 0 */
int lqb_agc_set_gain(lqb_stage s, float gain);
int lqb_agc_get_gain(lqb_stage s, float *gain);                  /* channel 0 */
int lqb_agc_get_gain_per_channel(lqb_stage s, float *gain, int n);
int lqb_agc_set_scale(lqb_stage s, float scale);
int lqb_agc_get_scale(lqb_stage s, float *scale);
int lqb_agc_lock(lqb_stage s, int locked);
/* Gain-loop arithmetic.  EXACT: liquid's expressions operation for operation (double-precision smoothing, correctly rounded
 * logf / expf) -- bit-identical to the oracle.  FAST: the same loop in single precision with lg2.approx, relative L2 1e-7 from
 * EXACT, 3-4x the throughput.  AUTO (default): FAST inside a chain where a FreqDem follows the AGC and no carrier-PLL
 * demodulator (AmpModem, BroadcastAM, FMStereo) does -- such a loop amplifies a last-bit difference to 1e-4..1e-3 --
 * EXACT everywhere else, including the stage executed on its own.  Locked or squelch-enabled AGCs always run EXACT. */
enum { LQB_AGC_AUTO = 0, LQB_AGC_EXACT = 1, LQB_AGC_FAST = 2 };
int lqb_agc_set_precision(lqb_stage s, int mode);
int lqb_agc_get_precision(lqb_stage s, int *mode);
int lqb_agc_squelch_enable(lqb_stage s, int enabled);
int lqb_agc_squelch_set_threshold(lqb_stage s, float threshold_dB);
int lqb_agc_squelch_get_threshold(lqb_stage s, float *threshold_dB);
int lqb_agc_squelch_set_timeout(lqb_stage s, unsigned timeout);
int lqb_agc_squelch_get_status(lqb_stage s, int *status);        /* channel 0, enum 0..7 of agc_docs.cpp:57-64 */
int lqb_agc_take_rise_count(lqb_stage s, unsigned *count);       /* RISE transitions since last take, all channels */

/* ---------------- ampmodem : AmpModem, demod.hpp:228-307 ------------------------------------
 * replaces ampmodem_create(mod, type, suppressed) (:305), ampmodem_demodulate_block (:294),
 * ampmodem_reset (:287).  complex in -> real out.  DSB with carrier (PLL) or suppressed (Costas); USB / LSB
 * suppressed (Hilbert pair, 0.5 * side-band / mod_index) or with carrier (carrier loop, Hilbert pair, DC blocker). */
int lqb_ampmodem_create(float mod_index, int type, int suppressed_carrier, int n_channels, lqb_stage *out);
int lqb_ampmodem_get_taps(lqb_stage s, float *lowpass, int *n_lowpass, float *dcblock, int *n_dcblock);
int lqb_ampmodem_get_nco_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n);

/* ---------------- firhilbf : SSBDemod demod.hpp:155-187, HilbertTransform utility.hpp:71-108 -------------
 * replaces firhilbf_create(m, As) (demod.hpp:163, utility.hpp:79-80) and the per-sample loops over
 * firhilbf_c2r_execute (demod.hpp:181,184), firhilbf_interp_execute (utility.hpp:93), firhilbf_decim_execute (:101).
 * SSB_LSB / SSB_USB: complex -> real (SSBDemod, m = 25, As = 60).  C2R: complex -> real, R2C: real -> complex, as
 * HilbertTransform's two branches produce them, including the overlapping-pair indexing of the reference loops;
 * the reference's one-element read past the end of the input (R2C) is taken as 0. */
#define LQB_FIRHILB_SSB_LSB 0
#define LQB_FIRHILB_SSB_USB 1
#define LQB_FIRHILB_C2R     2
#define LQB_FIRHILB_R2C     3
int lqb_firhilbf_create(int kind, int m, float as, int n_channels, lqb_stage *out);
int lqb_firhilbf_get_hq(lqb_stage s, float *hq, int *n);

/* ---------------- FMStereo : demod.hpp:4-85, wrapper.cpp:264-267 ------------------------------------------
 * replaces the per-sample demod_one loop (:43-48): freqdem(4.0), nco_crcf pilot mixer + PLL, two iirfilt_rrrf 75 us
 * de-emphasis filters, two resamp_rrrf_create_default(pcm_rate / iq_rate).  complex -> interleaved (left, right)
 * floats; lqb_stage_out_len counts floats (two per output pair).  lqb_stage_reset resets the resamplers only, as
 * FMStereo::reset does (:35-38).  phase_error (uninitialised in the reference) starts at 0. */
int lqb_fmstereo_create(float iq_rate, float pcm_rate, int n_channels, lqb_stage *out);
int lqb_fmstereo_get_state(lqb_stage s, uint32_t *theta, uint32_t *d_theta, float *phase_error, int n);
int lqb_fmstereo_get_deemph(lqb_stage s, float *b0, float *a1);

/* ---------------- BroadcastAM : demod.hpp:94-153, wrapper.cpp:259-262 -----------------------------------
 * replaces the per-sample demod_one loop (:124-131): nco_crcf PLL (bw 0.001, arg() detector), firfilt_crcf
 * kaiser(2*slen+1, 0.01, 40 dB), wdelaycf(slen), iirfilt_rrrf cheby2 order-3 high-pass.  complex -> real.
 * get_design: lowpass taps (2*slen+1) and the two DC-block sections B[6], A[6]. */
int lqb_broadcast_am_create(int slen, int n_channels, lqb_stage *out);
int lqb_broadcast_am_get_design(lqb_stage s, float *lp, int *n_lp, float *B, float *A);
int lqb_broadcast_am_get_nco_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n);

/* ---------------- freqdem : FreqDem, demod.hpp:189-219 --------------------------------------
 * replaces freqdem_create (:197), freqdem_demodulate_block (:216), freqdem_reset (:205). */
int lqb_freqdem_create(float kf, int n_channels, lqb_stage *out);

/* ---------------- fused chain ---------------------------------------------------------------
 * The reference composes stages in Python, one ndarray round trip per stage (README.md:53-54).
 * A chain borrows stage handles (it does not own them; their carried state is the chain's state)
 * and runs them as fused kernels with no HBM round trip between fused stages. */
int lqb_chain_create(lqb_chain *out);
int lqb_chain_append(lqb_chain c, lqb_stage s);
int lqb_chain_destroy(lqb_chain c);
int lqb_chain_out_len(lqb_chain c, size_t n, size_t *n_out);
int lqb_chain_execute(lqb_chain c, const void *x, size_t n, void *y, size_t y_capacity, size_t *n_out);
int lqb_chain_execute_dev(lqb_chain c, const void *x_dev, size_t n, void *y_dev, size_t y_capacity,
                          size_t *n_out, void *stream);
/* ---------------- int16 I/Q ingest : bytes_to_iq, utility.hpp:61-69 (wrapper.cpp:13) ----------------------
 * SDRs deliver interleaved little-endian int16 I/Q; the reference converts with (float)x / 32767.0f on the host
 * and then filters.  Here the conversion is fused into the first kernel of the chain (4 bytes per sample over
 * PCIe and HBM instead of 8); chains whose first kernel has no int16 variant convert in one extra pass.
 * iq: [n_channels x n] pairs (re, im) of int16.  lqb_bytes_to_iq is the stand-alone conversion. */
int lqb_chain_execute_i16(lqb_chain c, const int16_t *iq, size_t n, void *y, size_t y_capacity, size_t *n_out);
int lqb_chain_execute_i16_dev(lqb_chain c, const int16_t *iq_dev, size_t n, void *y_dev, size_t y_capacity,
                              size_t *n_out, void *stream);
int lqb_bytes_to_iq(const int16_t *iq, size_t n, lqb_cf *out);
/* human-readable launch plan ("seq[iir4+resamp] -> seq[agc+am+deemph]") and the number of kernel
 * launches the last execute issued */
int lqb_chain_plan(lqb_chain c, char *buf, size_t buf_len);
int lqb_chain_last_launches(lqb_chain c, int *n_launches);
/* names of the kernels the last execute launched, in launch order, separated by ';' (what the planner dispatched:
 * bench.py reports its roofline kernel from this) */
int lqb_chain_last_kernels(lqb_chain c, char *buf, size_t buf_len);
/* A call that fails after its first kernel launch leaves device state ahead of the host-side bookkeeping; the chain then
 * refuses further calls (LQB_EINVAL) until the stages have been reset and this has been called. */
int lqb_chain_clear_error(lqb_chain c);
/* Per-segment device timing of lqb_chain_execute_dev: while enabled, every call records one CUDA-event pair per
 * launch-plan segment on the caller's stream (up to 1024 calls).  get_timing waits for the recorded events and
 * returns, per segment, the elapsed milliseconds summed over the recorded calls. */
int lqb_chain_set_timing(lqb_chain c, int enabled);
int lqb_chain_get_timing(lqb_chain c, float *ms_per_segment, int capacity, int *n_segments, int *n_calls);
/* 0 = one kernel per stage, 1 = default (full-rate front and decimated tail as two kernels),
 * 2 = longest possible runs (whole AM receiver in one kernel); results are identical */
int lqb_chain_set_fusion(lqb_chain c, int level);
/* Overlapped device calls (default off).  The reference's objects are called block after block with their state carried
 * (README.md:53-58); block k+1's full-rate front depends on block k's front only.  With overlap enabled,
 * lqb_chain_execute_dev on a chain whose plan is "full-rate decimating kernel -> decimated-rate tail" (the AM receivers)
 * issues the tail on the chain's own low-priority stream and returns the caller's stream free for the next block's front,
 * which then runs concurrently with it (the front stages through a shallower ring so that a tail CTA fits beside it on
 * every SM).  Results are identical.  The output of such a call is complete only after
 * lqb_chain_wait(chain, stream) has made `stream` wait for every tail issued so far; any non-overlapped call on the
 * chain waits by itself.  Create the calling stream with lqb_stream_create(&s, 1) (high priority) so that a front is
 * dispatched ahead of the tail queued before it. */
int lqb_chain_set_overlap(lqb_chain c, int enabled);
int lqb_chain_wait(lqb_chain c, void *stream);
int lqb_stream_create(void **stream, int high_priority);
int lqb_stream_destroy(void *stream);

/* ---------------- synthetic input generators, device side -----------------------
 * Fill x_dev [n_channels x n] with the benchmark signal for blocks starting at absolute sample
 * index n0.  kind: 0 = AM IQ (configs 1, 5), 1 = complex Gaussian (config 2), 2 = tone+noise
 * (config 3), 3 = FM IQ with amplitude ramp (config 4). */
int lqb_synth_fill(int kind, void *x_dev, int n_channels, int channel0, size_t n, uint64_t n0,
                   uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif
