/*
 * liquid_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C99, scalar) of the liquid-dsp arithmetic that
 * colbyAtCRI/python-liquiddsp delegates to on the streaming baseband path,
 * plus the wrapper-level semantics of the reference's own src/ headers.
 *
 * PARITY UNPINNED: liquid-dsp is an un-vendored, un-pinned dependency of the
 * reference (README.md:8 "liquid-dsp (brew install)", CMakeLists.txt:8
 * find_library) and is absent from this environment, and the reference holds no
 * tests, golden vectors or fixtures.  This file restates liquid-dsp's published
 * algorithms (target semantics: liquid-dsp >= 1.4, fixed-point resampler phase,
 * 1024-entry NCO table) and is pinned only by known-answer tests against scipy
 * and closed-form integer identities (tests/test_oracle_kat.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * Rounding convention (see DESIGN.md "Oracle arithmetic"): expressions are
 * evaluated in liquid's source order with the multiply-add contraction GCC
 * applies under its default -ffp-contract=fast on an FMA target; the fused
 * operations are written out as fmaf() and the file is compiled with
 * -ffp-contract=off so the result does not depend on the compiler.
 */
#ifndef LIQUID_ORACLE_H
#define LIQUID_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* interleaved complex float (re, im) -- same memory layout as numpy complex64 */
typedef struct { float re, im; } orc_cf;

/* enums follow liquid.h ordering */
enum { ORC_IIRDES_BUTTER = 0, ORC_IIRDES_CHEBY1, ORC_IIRDES_CHEBY2, ORC_IIRDES_ELLIP, ORC_IIRDES_BESSEL };
enum { ORC_IIRDES_LOWPASS = 0, ORC_IIRDES_HIGHPASS, ORC_IIRDES_BANDPASS, ORC_IIRDES_BANDSTOP };
enum { ORC_NCO = 0, ORC_VCO };
enum { ORC_AMPMODEM_DSB = 0, ORC_AMPMODEM_USB, ORC_AMPMODEM_LSB };
enum {
    ORC_AGC_SQUELCH_UNKNOWN = 0, ORC_AGC_SQUELCH_ENABLED, ORC_AGC_SQUELCH_RISE,
    ORC_AGC_SQUELCH_SIGNALHI, ORC_AGC_SQUELCH_FALL, ORC_AGC_SQUELCH_SIGNALLO,
    ORC_AGC_SQUELCH_TIMEOUT, ORC_AGC_SQUELCH_DISABLED
};

/* ---- design (liquid iirdes.c / firdes.c / windows.c restated) ---- */
/* returns number of second-order sections written (3 floats each in B and A), <0 on error */
int   orc_iirdes_sos(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as,
                     float *B, float *A);
/* digital zeros/poles/gain before SOS pairing (for KATs against scipy zpk) */
int   orc_iirdes_dzpk(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as,
                      orc_cf *zd, orc_cf *pd, orc_cf *kd);
float orc_kaiser_beta_As(float as);
float orc_besseli0f(float z);
float orc_lngammaf(float z);
float orc_kaiser(unsigned i, unsigned wlen, float beta);
float orc_sincf(float x);
int   orc_firdes_kaiser(unsigned n, float fc, float as, float mu, float *h);
int   orc_firdes_notch(unsigned m, float f0, float as, float *h);

/* ---- iirfilt_crcf (SOS form) and iirfilt_rrrf (transfer-function form) ---- */
typedef struct orc_iirfilt_crcf_s *orc_iirfilt_crcf;
orc_iirfilt_crcf orc_iirfilt_crcf_create_sos(const float *B, const float *A, unsigned nsos);
orc_iirfilt_crcf orc_iirfilt_crcf_create_prototype(int ftype, int btype, unsigned order,
                                                   float fc, float f0, float ap, float as);
void  orc_iirfilt_crcf_destroy(orc_iirfilt_crcf q);
void  orc_iirfilt_crcf_reset(orc_iirfilt_crcf q);
unsigned orc_iirfilt_crcf_get_sos(orc_iirfilt_crcf q, float *B, float *A);
void  orc_iirfilt_crcf_execute_block(orc_iirfilt_crcf q, const orc_cf *x, unsigned n, orc_cf *y);
void  orc_iirfilt_crcf_freqresponse(orc_iirfilt_crcf q, float fc, orc_cf *H);
/* same recurrence evaluated in double with the same float coefficients ("fp64 truth") */
void  orc_iirfilt_crcf_execute_block_f64(const float *B, const float *A, unsigned nsos,
                                         double *state /* 4*nsos, re/im of v1,v2 */,
                                         const orc_cf *x, unsigned n, double *y_re_im);

typedef struct orc_iirfilt_rrrf_s *orc_iirfilt_rrrf;
orc_iirfilt_rrrf orc_iirfilt_rrrf_create(const float *b, unsigned nb, const float *a, unsigned na);
void  orc_iirfilt_rrrf_destroy(orc_iirfilt_rrrf q);
void  orc_iirfilt_rrrf_reset(orc_iirfilt_rrrf q);
void  orc_iirfilt_rrrf_execute(orc_iirfilt_rrrf q, float x, float *y);
void  orc_iirfilt_rrrf_freqresponse(orc_iirfilt_rrrf q, float fc, orc_cf *H);

/* ---- firfilt ---- */
typedef struct orc_firfilt_s *orc_firfilt; /* real taps; complex (crcf) or real (rrrf) samples */
orc_firfilt orc_firfilt_create(const float *h, unsigned n);
orc_firfilt orc_firfilt_create_kaiser(unsigned n, float fc, float as, float mu);
orc_firfilt orc_firfilt_create_dc_blocker(unsigned m, float as);
void  orc_firfilt_destroy(orc_firfilt q);
void  orc_firfilt_reset(orc_firfilt q);
void  orc_firfilt_set_scale(orc_firfilt q, float scale);
unsigned orc_firfilt_get_taps(orc_firfilt q, float *h); /* in design order h[0..n-1] */
void  orc_firfilt_crcf_push(orc_firfilt q, orc_cf x);
void  orc_firfilt_crcf_execute(orc_firfilt q, orc_cf *y);
void  orc_firfilt_crcf_execute_block(orc_firfilt q, const orc_cf *x, unsigned n, orc_cf *y);
void  orc_firfilt_rrrf_push(orc_firfilt q, float x);
void  orc_firfilt_rrrf_execute(orc_firfilt q, float *y);
void  orc_firfilt_rrrf_execute_block(orc_firfilt q, const float *x, unsigned n, float *y);
void  orc_firfilt_freqresponse(orc_firfilt q, float fc, orc_cf *H);

/* ---- resamp_cccf (fixed-point phase) ---- */
typedef struct orc_resamp_s *orc_resamp;
orc_resamp orc_resamp_create(float rate, unsigned m, float fc, float as, unsigned npfb);
void  orc_resamp_destroy(orc_resamp q);
void  orc_resamp_reset(orc_resamp q);
void  orc_resamp_set_rate(orc_resamp q, float rate);
uint32_t orc_resamp_get_step(orc_resamp q);
uint32_t orc_resamp_get_phase(orc_resamp q);
unsigned orc_resamp_get_npfb(orc_resamp q);
unsigned orc_resamp_get_sublen(orc_resamp q);
/* bank taps as stored by firpfb (reversed): out[i*sublen + k] */
void  orc_resamp_get_bank(orc_resamp q, float *out);
void  orc_resamp_execute(orc_resamp q, orc_cf x, orc_cf *y, unsigned *nw);
/* reference wrapper loop: ComplexResampler::execute, resampler.hpp:160-172 */
void  orc_resamp_set_real_taps(orc_resamp q, int on);   /* resamp_crcf / resamp_rrrf dot product */
orc_resamp orc_resamp_create_default(float rate);       /* resamp_crcf_create_default / resamp_rrrf_create_default */
unsigned orc_resamp_execute_block(orc_resamp q, const orc_cf *x, unsigned n, orc_cf *y);

/* ---- nco_crcf ---- */
typedef struct orc_nco_s *orc_nco;
orc_nco orc_nco_create(int type);
void  orc_nco_destroy(orc_nco q);
void  orc_nco_reset(orc_nco q);
uint32_t orc_nco_constrain(float theta);
void  orc_nco_set_frequency(orc_nco q, float f);
void  orc_nco_adjust_frequency(orc_nco q, float df);
void  orc_nco_set_phase(orc_nco q, float phi);
void  orc_nco_adjust_phase(orc_nco q, float dphi);
float orc_nco_get_frequency(orc_nco q);
float orc_nco_get_phase(orc_nco q);
uint32_t orc_nco_get_theta_u32(orc_nco q);
uint32_t orc_nco_get_dtheta_u32(orc_nco q);
void  orc_nco_set_u32(orc_nco q, uint32_t theta, uint32_t d_theta);
void  orc_nco_pll_set_bandwidth(orc_nco q, float bw);
void  orc_nco_pll_step(orc_nco q, float dphi);
void  orc_nco_step(orc_nco q);
void  orc_nco_sincos(orc_nco q, float *s, float *c);
void  orc_nco_mix_up(orc_nco q, orc_cf x, orc_cf *y);
void  orc_nco_mix_down(orc_nco q, orc_cf x, orc_cf *y);
void  orc_nco_mix_block_up(orc_nco q, const orc_cf *x, orc_cf *y, unsigned n);
void  orc_nco_mix_block_down(orc_nco q, const orc_cf *x, orc_cf *y, unsigned n);
const float *orc_nco_sintab(void); /* 1024 entries */

/* ---- agc_crcf ---- */
typedef struct orc_agc_s *orc_agc;
orc_agc orc_agc_create(void);
void  orc_agc_destroy(orc_agc q);
void  orc_agc_reset(orc_agc q);
void  orc_agc_execute(orc_agc q, orc_cf x, orc_cf *y);
void  orc_agc_lock(orc_agc q);
void  orc_agc_unlock(orc_agc q);
void  orc_agc_set_bandwidth(orc_agc q, float bw);
float orc_agc_get_bandwidth(orc_agc q);
float orc_agc_get_signal_level(orc_agc q);
void  orc_agc_set_signal_level(orc_agc q, float x2);
float orc_agc_get_rssi(orc_agc q);
void  orc_agc_set_rssi(orc_agc q, float rssi);
float orc_agc_get_gain(orc_agc q);
void  orc_agc_set_gain(orc_agc q, float g);
float orc_agc_get_scale(orc_agc q);
void  orc_agc_set_scale(orc_agc q, float s);
void  orc_agc_squelch_enable(orc_agc q);
void  orc_agc_squelch_disable(orc_agc q);
void  orc_agc_squelch_set_threshold(orc_agc q, float t);
float orc_agc_squelch_get_threshold(orc_agc q);
void  orc_agc_squelch_set_timeout(orc_agc q, unsigned t);
int   orc_agc_squelch_get_status(orc_agc q);
float orc_agc_get_y2_prime(orc_agc q);
/* reference wrapper loop AGC::execute (agc.hpp:109-128): zeroing in ENABLED / SIGNALLO, RISE events.
 * state_last is the caller-held equivalent of the wrapper's static; rise_idx receives up to
 * rise_cap sample indices at which onRise would fire; returns the number of RISE events. */
unsigned orc_wrap_agc_execute(orc_agc q, const orc_cf *x, unsigned n, orc_cf *y,
                              int *state_last, unsigned *rise_idx, unsigned rise_cap);

/* ---- ampmodem ---- */
typedef struct orc_ampmodem_s *orc_ampmodem;
orc_ampmodem orc_ampmodem_create(float mod_index, int type, int suppressed_carrier);
void  orc_ampmodem_destroy(orc_ampmodem q);
void  orc_ampmodem_reset(orc_ampmodem q);
void  orc_ampmodem_demodulate_block(orc_ampmodem q, const orc_cf *x, unsigned n, float *y);
unsigned orc_ampmodem_get_lowpass_taps(orc_ampmodem q, float *h);
unsigned orc_ampmodem_get_dcblock_taps(orc_ampmodem q, float *h);
void  orc_ampmodem_get_nco(orc_ampmodem q, uint32_t *theta, uint32_t *d_theta);

/* ---- freqdem ---- */
typedef struct orc_freqdem_s *orc_freqdem;
orc_freqdem orc_freqdem_create(float kf);
void  orc_freqdem_destroy(orc_freqdem q);
void  orc_freqdem_reset(orc_freqdem q);
void  orc_freqdem_demodulate_block(orc_freqdem q, const orc_cf *x, unsigned n, float *y);

/* ---- reference wrapper objects that carry arithmetic of their own ---- */
/* DeemphasisFilter (iirfilter.hpp:358-392) */
void  orc_wrap_deemph_coeffs(float sample_rate, float *b0, float *a1);
orc_iirfilt_rrrf orc_wrap_deemph_create(float sample_rate);
void  orc_wrap_deemph_execute(orc_iirfilt_rrrf q, const float *x, unsigned n, float *y);
/* bytes_to_iq (utility.hpp:61-69) */
typedef struct orc_firhilbf_s *orc_firhilbf;   /* firhilbf: SSBDemod demod.hpp:155-187, HilbertTransform utility.hpp:71-108 */
orc_firhilbf orc_firhilbf_create(unsigned m, float as);
void  orc_firhilbf_destroy(orc_firhilbf q);
void  orc_firhilbf_reset(orc_firhilbf q);
unsigned orc_firhilbf_get_hq(orc_firhilbf q, float *hq);
void  orc_firhilbf_r2c_execute(orc_firhilbf q, float x, orc_cf *y);
void  orc_firhilbf_c2r_execute(orc_firhilbf q, orc_cf x, float *y_lsb, float *y_usb);
void  orc_firhilbf_decim_execute(orc_firhilbf q, const float *x, orc_cf *y);
void  orc_firhilbf_interp_execute(orc_firhilbf q, orc_cf x, float *y);
void  orc_wrap_ssb_execute(orc_firhilbf q, int usb, const orc_cf *x, unsigned n, float *y);
void  orc_wrap_hilbert_c2r(orc_firhilbf q, const orc_cf *z, unsigned n, float *y);
void  orc_wrap_hilbert_r2c(orc_firhilbf q, const float *z, unsigned n, orc_cf *y);
typedef struct orc_fmstereo_s *orc_fmstereo;   /* FMStereo, demod.hpp:4-85 */
orc_fmstereo orc_wrap_fmstereo_create(float iq_rate, float pcm_rate);
void  orc_wrap_fmstereo_destroy(orc_fmstereo q);
void  orc_wrap_fmstereo_reset(orc_fmstereo q);
void  orc_wrap_fmstereo_get_state(orc_fmstereo q, uint32_t *theta, uint32_t *d_theta, float *pe);
void  orc_wrap_fmstereo_get_deemph(orc_fmstereo q, float *b0, float *a1);
unsigned orc_wrap_fmstereo_execute(orc_fmstereo q, const orc_cf *x, unsigned n, float *y);   /* y: interleaved L, R */
typedef struct orc_bam_s *orc_bam;      /* BroadcastAM, demod.hpp:94-153 */
orc_bam orc_wrap_bam_create(int m);
void  orc_wrap_bam_destroy(orc_bam q);
void  orc_wrap_bam_reset(orc_bam q);
void  orc_wrap_bam_get_nco(orc_bam q, uint32_t *theta, uint32_t *d_theta);
unsigned orc_wrap_bam_get_design(orc_bam q, float *lp, float *B, float *A);
void  orc_wrap_bam_set_dcblock(orc_bam q, const float *B, const float *A, unsigned nsos);
void  orc_wrap_bam_execute(orc_bam q, const orc_cf *x, unsigned n, float *y);
void  orc_wrap_bytes_to_iq(const int16_t *iq, unsigned n, orc_cf *y);

/* ---- README AMRadio chain (README.md:41-58) as one object, for the CPU baseline ---- */
typedef struct orc_amradio_s *orc_amradio;
orc_amradio orc_amradio_create(float bandwidth, float iq_rate, float pcm_rate);
void  orc_amradio_destroy(orc_amradio q);
/* returns number of PCM samples written (<= n); scratch-free, block size <= 1<<20 */
unsigned orc_amradio_execute(orc_amradio q, const orc_cf *iq, unsigned n, float *pcm);

#ifdef __cplusplus
}
#endif
#endif
