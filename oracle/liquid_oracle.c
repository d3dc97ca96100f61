/*
 * liquid_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  PARITY UNPINNED (see header).
 *
 * Scalar C99 restatement of the liquid-dsp functions python-liquiddsp calls on the
 * streaming baseband path.  Every function cites (a) the reference call site in
 * /root/reference/src that reaches it and (b) the liquid-dsp source file whose
 * published algorithm is restated (liquid-dsp is NOT vendored by the reference).
 *
 * Compile: gcc -O2 -std=gnu11 -ffp-contract=off -fno-fast-math -fPIC -shared [-mfma]
 * FMA(a,b,c) marks the places where an FMA-target GCC build (-ffp-contract=fast,
 * GCC's default) fuses liquid's source expression; -DORC_NO_FMA gives the unfused
 * variant (used by tests only to report how far apart the two conventions are).
 */
#include "liquid_oracle.h"
#include <math.h>
#include <complex.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifdef ORC_NO_FMA
static inline float  FMA (float a, float b, float c)    { float  p = a * b; return p + c; }
static inline double FMAD(double a, double b, double c) { double p = a * b; return p + c; }
#else
#define FMA(a, b, c)  fmaf((a), (b), (c))
#define FMAD(a, b, c) fma((a), (b), (c))
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------
 * small complex-float helpers with explicit evaluation order (used by the design code, where
 * liquid uses C99 `float complex`; products/quotients follow the textbook formulas GCC emits
 * under -fcx-limited-range-equivalent fast paths for finite operands)
 * ---------------------------------------------------------------------------------------- */
typedef struct { float r, i; } cfx;
static inline cfx cmk(float r, float i) { cfx z = { r, i }; return z; }
static inline cfx cadd(cfx a, cfx b) { return cmk(a.r + b.r, a.i + b.i); }
static inline cfx csub(cfx a, cfx b) { return cmk(a.r - b.r, a.i - b.i); }
static inline cfx cmul(cfx a, cfx b) { return cmk(a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r); }
static inline cfx cscale(cfx a, float s) { return cmk(a.r * s, a.i * s); }
static inline cfx cneg(cfx a) { return cmk(-a.r, -a.i); }
static inline cfx cconj(cfx a) { return cmk(a.r, -a.i); }
static inline cfx cdivf(cfx a, cfx b)
{
    /* Smith's algorithm, as libgcc __divsc3 does for finite operands */
    float ratio, denom;
    if (fabsf(b.r) < fabsf(b.i)) {
        ratio = b.r / b.i; denom = (b.r * ratio) + b.i;
        return cmk(((a.r * ratio) + a.i) / denom, ((a.i * ratio) - a.r) / denom);
    }
    ratio = b.i / b.r; denom = (b.i * ratio) + b.r;
    return cmk(((a.i * ratio) + a.r) / denom, (a.i - (a.r * ratio)) / denom);
}

/* ==========================================================================================
 *  Window functions / FIR design     liquid: src/math/src/windows.c, math.bessel.c,
 *                                            src/filter/src/firdes.c
 *  reached from: resamp_cccf_create (resampler.hpp:136), ampmodem_create (demod.hpp:305),
 *                firfilt_*_create_kaiser (firfilter.hpp:57, demod.hpp:105)
 * ======================================================================================== */
float orc_lngammaf(float z)
{
    float g;
    if (z < 0) return NAN;
    if (z < 10.0f)          /* recursion ln G(z) = ln G(z+1) - ln z */
        return orc_lngammaf(z + 1.0f) - logf(z);
    g  = 0.5 * (logf(2 * M_PI) - log(z));                              /* double arithmetic, float store */
    g += z * (logf(z + (1 / (12.0f * z - 0.1f / z))) - 1);
    return g;
}

float orc_besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    unsigned k; float t, y = 0.0f;
    for (k = 0; k < 32; k++) {
        t = k * logf(0.5f * z) - orc_lngammaf((float)k + 1.0f);
        y += expf(2 * t);
    }
    return y;
}

float orc_kaiser_beta_As(float as)
{
    as = fabsf(as);
    float beta;
    if (as > 50.0f)      beta = 0.1102f * (as - 8.7f);
    else if (as > 21.0f) beta = 0.5842 * powf(as - 21, 0.4f) + 0.07886f * (as - 21);
    else                 beta = 0.0f;
    return beta;
}

/* liquid_kaiser(i, wlen, beta), liquid-dsp >= 1.3.2 form (r normalised by wlen-1) */
float orc_kaiser(unsigned i, unsigned wlen, float beta)
{
    float t = (float)i - (float)(wlen - 1) / 2;
    float r = 2.0f * t / (float)(wlen - 1);
    float a = orc_besseli0f(beta * sqrtf(1 - r * r));
    float b = orc_besseli0f(beta);
    return a / b;
}

float orc_sincf(float x)
{
    if (fabsf(x) < 0.01f)
        return cosf(M_PI * x / 2.0f) * cosf(M_PI * x / 4.0f) * cosf(M_PI * x / 8.0f);
    return sinf(M_PI * x) / (M_PI * x);
}

int orc_firdes_kaiser(unsigned n, float fc, float as, float mu, float *h)
{
    if (mu < -0.5f || mu > 0.5f || fc <= 0.0f || fc > 0.5f || n == 0) return -1;
    float beta = orc_kaiser_beta_As(as);
    for (unsigned i = 0; i < n; i++) {
        float t  = (float)i - (float)(n - 1) / 2 + mu;
        float h1 = orc_sincf(2.0f * fc * t);
        float h2 = orc_kaiser(i, n, beta);
        h[i] = h1 * h2;
    }
    return 0;
}

/* liquid_firdes_notch(m, f0, as, h): used with f0 = 0 by firfilt_rrrf_create_dc_blocker */
int orc_firdes_notch(unsigned m, float f0, float as, float *h)
{
    if (m < 1 || m > 1000 || f0 < -0.5f || f0 > 0.5f || as <= 0.0f) return -1;
    float beta = orc_kaiser_beta_As(as);
    unsigned h_len = 2 * m + 1, i;
    float scale = 0.0f;
    for (i = 0; i < h_len; i++) {
        float p = -cosf(2.0f * M_PI * f0 * ((float)i - (float)m));
        float w = orc_kaiser(i, h_len, beta);
        h[i] = p * w;
        scale += h[i] * p;
    }
    for (i = 0; i < h_len; i++) h[i] /= scale;
    h[m] += 1.0f;
    return 0;
}

/* ==========================================================================================
 *  IIR design    liquid: src/filter/src/iirdes.c, butter.c, cheby1.c, cheby2.c
 *  reached from: iirfilt_crcf_create_prototype (iirfilter.hpp:275)
 * ======================================================================================== */
static void butter_azpkf(unsigned n, cfx *pa)
{
    unsigned r = n % 2, L = (n - r) / 2, i, k = 0;
    for (i = 0; i < L; i++) {
        float theta = (float)(2 * (i + 1) + n - 1) * M_PI / (float)(2 * n);
        pa[k++] = cmk(cosf(theta),  sinf(theta));     /* cexpf( j theta) */
        pa[k++] = cmk(cosf(theta), -sinf(theta));     /* cexpf(-j theta) */
    }
    if (r) pa[k++] = cmk(-1.0f, 0.0f);
}

static void cheby1_azpkf(unsigned n, float ep, cfx *pa)
{
    float t0 = sqrt(1.0 + 1.0 / (ep * ep));
    float tp = powf(t0 + 1.0 / ep, 1.0 / (float)(n));
    float tm = powf(t0 - 1.0 / ep, 1.0 / (float)(n));
    float b = 0.5 * (tp + tm), a = 0.5 * (tp - tm);
    unsigned r = n % 2, L = (n - r) / 2, i, k = 0;
    for (i = 0; i < L; i++) {
        float theta = (float)(2 * (i + 1) + n - 1) * M_PI / (float)(2 * n);
        pa[k++] = cmk(a * cosf(theta), -(b * sinf(theta)));
        pa[k++] = cmk(a * cosf(theta),   b * sinf(theta));
    }
    if (r) pa[k++] = cmk(-a, 0.0f);
}

static void cheby2_azpkf(unsigned n, float es, cfx *za, cfx *pa)
{
    float t0 = sqrt(1.0 + 1.0 / (es * es));
    float tp = powf(t0 + 1.0 / es, 1.0 / (float)(n));
    float tm = powf(t0 - 1.0 / es, 1.0 / (float)(n));
    float b = 0.5 * (tp + tm), a = 0.5 * (tp - tm);
    unsigned r = n % 2, L = (n - r) / 2, i, k = 0;
    for (i = 0; i < L; i++) {
        float theta = (float)(2 * (i + 1) + n - 1) * M_PI / (float)(2 * n);
        pa[k++] = cdivf(cmk(1.0f, 0.0f), cmk(a * cosf(theta), -(b * sinf(theta))));
        pa[k++] = cdivf(cmk(1.0f, 0.0f), cmk(a * cosf(theta),   b * sinf(theta)));
    }
    if (r) pa[k++] = cmk(-1.0f / a, 0.0f);
    k = 0;
    for (i = 0; i < L; i++) {
        float theta = (float)(0.5f * M_PI * (2 * (i + 1) - 1) / (float)(n));
        za[k++] = cdivf(cmk(-1.0f, 0.0f), cmk(0.0f, cosf(theta)));
        za[k++] = cdivf(cmk( 1.0f, 0.0f), cmk(0.0f, cosf(theta)));
    }
}

/* ---- elliptic prototype   liquid: src/filter/src/ellip.c (after Orfanidis, "Lecture notes on elliptic filter
 * design"): Landen sequences, n = 7 iterations, everything in float ---- */
#define ELLIP_NITER 7
static void landenf(float k, unsigned n, float *v)
{
    for (unsigned i = 0; i < n; i++) { float kp = sqrtf(1 - k * k); k = (1 - kp) / (1 + kp); v[i] = k; }
}
static void ellipkf(float k, unsigned n, float *K, float *Kp)
{
    float kmin = 4e-4f, kmax = sqrtf(1 - kmin * kmin), kp = sqrtf(1 - k * k), v[16];
    unsigned i;
    if (k > kmax) { float L = -logf(0.25f * kp); *K = L + 0.25f * (L - 1) * kp * kp; }
    else { landenf(k, n, v); *K = M_PI * 0.5f; for (i = 0; i < n; i++) *K *= (1 + v[i]); }
    if (k < kmin) { float L = -logf(k * 0.25f); *Kp = L + 0.25f * (L - 1) * k * k; }
    else { landenf(kp, n, v); *Kp = M_PI * 0.5f; for (i = 0; i < n; i++) *Kp *= (1 + v[i]); }
}
static float ellipdegf(float N, float k1, unsigned n)
{
    float K1, K1p;
    ellipkf(k1, n, &K1, &K1p);
    float q1 = expf(-M_PI * K1p / K1), q = powf(q1, 1.0f / N), b = 0.0f, a = 0.0f;
    unsigned m;
    for (m = 0; m < n; m++) b += powf(q, (float)(m * (m + 1)));
    for (m = 1; m < n; m++) a += powf(q, (float)(m * m));
    float g = b / (1.0f + 2.0f * a);
    return 4.0f * sqrtf(q) * g * g;
}
static float complex ellip_up(float complex w, float k, unsigned n)     /* ascending Landen recursion of cd / sn */
{
    float v[16]; landenf(k, n, v);
    for (unsigned i = n; i > 0; i--) w = (1 + v[i - 1]) * w / (1 + v[i - 1] * w * w);
    return w;
}
static float complex ellip_cdf(float complex u, float k, unsigned n) { return ellip_up(ccosf(u * (float)(M_PI * 0.5)), k, n); }
static float complex ellip_snf(float complex u, float k, unsigned n) { return ellip_up(csinf(u * (float)(M_PI * 0.5)), k, n); }
static float complex ellip_acdf(float complex w, float k, unsigned n)
{
    float v[16]; landenf(k, n, v);
    for (unsigned i = 0; i < n; i++) {
        float v1 = (i == 0) ? k : v[i - 1];
        w = w / (1 + csqrtf(1 - w * w * v1 * v1)) * 2.0f / (1 + v[i]);
    }
    return cacosf(w) * 2.0f / (float)M_PI;
}
static float complex ellip_asnf(float complex w, float k, unsigned n) { return 1.0f - ellip_acdf(w, k, n); }
static cfx from_c99(float complex z) { return cmk(crealf(z), cimagf(z)); }

static void ellip_azpkf(unsigned n_, float ep, float es, cfx *za, cfx *pa)
{
    const unsigned n = ELLIP_NITER;
    float k1 = ep / es, N = (float)n_;
    float k = ellipdegf(N, k1, n);
    unsigned r = n_ % 2, L = (n_ - r) / 2, i, t = 0;
    float complex v0 = -I * ellip_asnf(I / ep, k1, n) / N;
    for (i = 0; i < L; i++) {
        float u = (2.0f * (i + 1) - 1.0f) / N;
        float complex zeta = ellip_cdf(u, k, n);
        float complex z = I * 1.0f / (k * zeta);
        float complex p = I * ellip_cdf(u - I * v0, k, n);
        za[t] = from_c99(z); pa[t++] = from_c99(p);
        za[t] = from_c99(conjf(z)); pa[t++] = from_c99(conjf(p));
    }
    if (r) pa[t++] = from_c99(I * ellip_snf(I * v0, k, n));
}

/* ---- Bessel prototype   liquid: src/filter/src/bessel.c.  Poles = roots of the reverse Bessel polynomial
 * theta_n(s) = sum_k (2n-k)! / (2^(n-k) k! (n-k)!) s^k, divided by the approximate 3 dB frequency
 * sqrt((2n-1) ln 2).  liquid finds the roots with Orchard's recursion in float; here they are found in double
 * (Durand-Kerner) and narrowed -- the same numbers to float accuracy. ---- */
static void bessel_azpkf(unsigned n, cfx *pa)
{
    double c[18];
    unsigned k, i, it;
    for (k = 0; k <= n; k++)
        c[k] = exp(lgamma(2.0 * n - k + 1) - lgamma((double)k + 1) - lgamma((double)(n - k) + 1) - (double)(n - k) * log(2.0));
    for (k = 0; k <= n; k++) c[k] /= c[n];                 /* monic (c[n] = 1 already; keeps the intent explicit) */
    double complex z[17];
    for (i = 0; i < n; i++) z[i] = cpow(0.4 + 0.9 * I, (double)i) * (double)n;
    for (it = 0; it < 500; it++) {
        double worst = 0;
        for (i = 0; i < n; i++) {
            double complex num = 1.0, den = 1.0;
            for (k = n; k-- > 0;) num = num * z[i] + c[k];
            for (k = 0; k < n; k++) if (k != i) den *= (z[i] - z[k]);
            double complex d = num / den;
            z[i] -= d;
            if (cabs(d) > worst) worst = cabs(d);
        }
        if (worst < 1e-14 * n) break;
    }
    float w3dB = sqrtf((2 * n - 1) * logf(2.0f));
    for (i = 0; i < n; i++) pa[i] = cmk((float)creal(z[i]) / w3dB, (float)cimag(z[i]) / w3dB);
}

static float iirdes_freqprewarp(int btype, float fc, float f0)
{
    float m = 0.0f;
    switch (btype) {
    case ORC_IIRDES_LOWPASS:  m = tanf(M_PI * fc); break;
    case ORC_IIRDES_HIGHPASS: m = -cosf(M_PI * fc) / sinf(M_PI * fc); break;
    case ORC_IIRDES_BANDPASS: m = (cosf(2 * M_PI * fc) - cosf(2 * M_PI * f0)) / sinf(2 * M_PI * fc); break;
    case ORC_IIRDES_BANDSTOP: m = sinf(2 * M_PI * fc) / (cosf(2 * M_PI * fc) - cosf(2 * M_PI * f0)); break;
    }
    return fabsf(m);
}

/* bilinear_zpkf: gain seeded with the NOMINAL digital gain k0 (the analog gain is ignored) */
static void bilinear_zpkf(const cfx *za, unsigned nza, const cfx *pa, unsigned npa, cfx ka, float m,
                          cfx *zd, cfx *pd, cfx *kd)
{
    unsigned n = nza > npa ? nza : npa, i;
    cfx G = ka, one = cmk(1.0f, 0.0f);
    for (i = 0; i < n; i++) {
        if (i < nza) { cfx zm = cscale(za[i], m); zd[i] = cdivf(cadd(one, zm), csub(one, zm)); }
        else zd[i] = cmk(-1.0f, 0.0f);
        if (i < npa) { cfx pm = cscale(pa[i], m); pd[i] = cdivf(cadd(one, pm), csub(one, pm)); }
        else pd[i] = cmk(-1.0f, 0.0f);
        G = cmul(G, cdivf(csub(one, pd[i]), csub(one, zd[i])));
    }
    *kd = G;
}

/* iirdes_dzpk_lp2bp: each low-pass root maps to two band-pass roots */
static cfx csqrtf_(cfx z)
{
    float m = hypotf(z.r, z.i);
    float sr = sqrtf(0.5f * (m + z.r)), si = sqrtf(0.5f * (m - z.r));
    return cmk(sr, z.i < 0 ? -si : si);
}
static void iirdes_dzpk_lp2bp(const cfx *zd, const cfx *pd, unsigned n, float f0, cfx *zdt, cfx *pdt)
{
    float c0 = cosf(2 * M_PI * f0);
    unsigned i;
    for (i = 0; i < n; i++) {
        cfx t0 = cadd(cmk(1.0f, 0.0f), zd[i]);
        cfx s  = csqrtf_(csub(cscale(cmul(t0, t0), c0 * c0), cscale(zd[i], 4.0f)));
        zdt[2 * i + 0] = cscale(cadd(cscale(t0, c0), s), 0.5f);
        zdt[2 * i + 1] = cscale(csub(cscale(t0, c0), s), 0.5f);
        t0 = cadd(cmk(1.0f, 0.0f), pd[i]);
        s  = csqrtf_(csub(cscale(cmul(t0, t0), c0 * c0), cscale(pd[i], 4.0f)));
        pdt[2 * i + 0] = cscale(cadd(cscale(t0, c0), s), 0.5f);
        pdt[2 * i + 1] = cscale(csub(cscale(t0, c0), s), 0.5f);
    }
}

/* liquid_cplxpair + liquid_cplxpair_cleanup (src/math/src/poly... in liquid: math.complex / iirdes.c) */
static void cplxpair(const cfx *z, unsigned n, float tol, cfx *p)
{
    unsigned char paired[64]; memset(paired, 0, sizeof paired);
    unsigned i, j, k = 0, num_pairs = 0;
    for (i = 0; i < n; i++) {
        if (paired[i] || fabsf(z[i].i) < tol) continue;
        for (j = 0; j < n; j++) {
            if (j == i || paired[j] || fabsf(z[j].i) < tol) continue;
            if (fabsf(z[i].i + z[j].i) < tol && fabsf(z[i].r - z[j].r) < tol) {
                p[k++] = z[i]; p[k++] = z[j]; paired[i] = paired[j] = 1; num_pairs++;
                break;
            }
        }
    }
    for (i = 0; i < n; i++) {
        if (paired[i]) continue;
        if (z[i].i > tol) fprintf(stderr, "oracle: cplxpair: complex numbers cannot be paired\n");
        else { p[k++] = z[i]; paired[i] = 1; }
    }
    /* cleanup: perfect conjugates, negative imaginary first */
    for (i = 0; i < num_pairs; i++) {
        p[2 * i + 0] = p[2 * i].i < 0 ? p[2 * i] : cconj(p[2 * i]);
        p[2 * i + 1] = cconj(p[2 * i + 0]);
    }
    /* pairs by increasing real part (bubble sort, as liquid) */
    for (i = 0; i < num_pairs; i++)
        for (j = num_pairs - 1; j > i; j--)
            if (p[2 * (j - 1)].r > p[2 * j].r) {
                cfx t0 = p[2 * (j - 1)], t1 = p[2 * (j - 1) + 1];
                p[2 * (j - 1)] = p[2 * j]; p[2 * (j - 1) + 1] = p[2 * j + 1];
                p[2 * j] = t0; p[2 * j + 1] = t1;
            }
    /* pure-real values by increasing value */
    for (i = 2 * num_pairs; i < n; i++)
        for (j = n - 1; j > i; j--)
            if (p[j - 1].r > p[j].r) { cfx t = p[j - 1]; p[j - 1] = p[j]; p[j] = t; }
}

static unsigned dzpk2sosf(const cfx *zd, const cfx *pd, unsigned n, cfx kd, float *B, float *A)
{
    float tol = 1e-6f;
    cfx zp[64], pp[64];
    cplxpair(zd, n, tol, zp);
    cplxpair(pd, n, tol, pp);
    unsigned r = n % 2, L = (n - r) / 2, i;
    for (i = 0; i < L; i++) {
        cfx p0 = cneg(pp[2 * i + 0]), p1 = cneg(pp[2 * i + 1]);
        cfx z0 = cneg(zp[2 * i + 0]), z1 = cneg(zp[2 * i + 1]);
        A[3 * i + 0] = 1.0f; A[3 * i + 1] = cadd(p0, p1).r; A[3 * i + 2] = cmul(p0, p1).r;
        B[3 * i + 0] = 1.0f; B[3 * i + 1] = cadd(z0, z1).r; B[3 * i + 2] = cmul(z0, z1).r;
    }
    if (r) {
        cfx p0 = cneg(pp[n - 1]), z0 = cneg(zp[n - 1]);
        A[3 * i + 0] = 1.0f; A[3 * i + 1] = p0.r; A[3 * i + 2] = 0.0f;
        B[3 * i + 0] = 1.0f; B[3 * i + 1] = z0.r; B[3 * i + 2] = 0.0f;
    }
    float k = powf(kd.r, 1.0f / (float)(L + r));
    for (i = 0; i < L + r; i++) { B[3 * i + 0] *= k; B[3 * i + 1] *= k; B[3 * i + 2] *= k; }
    return L + r;
}

/* liquid_iirdes up to the digital zpk; returns digital order (doubled for bandpass/bandstop) */
static int iirdes_dzpk(int ftype, int btype, unsigned n, float fc, float f0, float ap, float as,
                       cfx *zd, cfx *pd, cfx *kd)
{
    if (n == 0 || n > 16) return -1;
    if (fc <= 0 || fc >= 0.5f) return -1;
    if ((btype == ORC_IIRDES_BANDPASS || btype == ORC_IIRDES_BANDSTOP) && (f0 < 0 || f0 > 0.5f)) return -1;
    if (ap <= 0 || as <= 0) return -1;
    cfx pa[16], za[16], k0 = cmk(1.0f, 0.0f);
    unsigned npa = n, nza = 0, r = n % 2, L = (n - r) / 2;
    float epsilon;
    switch (ftype) {
    case ORC_IIRDES_BUTTER:
        nza = 0; butter_azpkf(n, pa); break;
    case ORC_IIRDES_CHEBY1:
        nza = 0;
        epsilon = sqrtf(powf(10.0f, ap / 10.0f) - 1.0f);
        k0 = cmk(r ? 1.0f : 1.0f / sqrtf(1.0f + epsilon * epsilon), 0.0f);
        cheby1_azpkf(n, epsilon, pa); break;
    case ORC_IIRDES_CHEBY2:
        nza = 2 * L;
        epsilon = powf(10.0f, -as / 20.0f);
        cheby2_azpkf(n, epsilon, za, pa); break;
    case ORC_IIRDES_ELLIP: {
        nza = 2 * L;
        float Gp = powf(10.0f, -ap / 20.0f), Gs = powf(10.0f, -as / 20.0f);
        float ep = sqrtf(1.0f / (Gp * Gp) - 1.0f), es = sqrtf(1.0f / (Gs * Gs) - 1.0f);
        k0 = cmk(r ? 1.0f : 1.0f / sqrtf(1.0f + ep * ep), 0.0f);
        ellip_azpkf(n, ep, es, za, pa); break;
    }
    case ORC_IIRDES_BESSEL:
        nza = 0; bessel_azpkf(n, pa); break;
    default:
        return -2;
    }
    float m = iirdes_freqprewarp(btype, fc, f0);
    bilinear_zpkf(za, nza, pa, npa, k0, m, zd, pd, kd);
    if (btype == ORC_IIRDES_HIGHPASS || btype == ORC_IIRDES_BANDSTOP)
        for (unsigned i = 0; i < n; i++) { zd[i] = cneg(zd[i]); pd[i] = cneg(pd[i]); }
    if (btype == ORC_IIRDES_BANDPASS || btype == ORC_IIRDES_BANDSTOP) {
        cfx zd1[32], pd1[32];
        iirdes_dzpk_lp2bp(zd, pd, n, f0, zd1, pd1);
        memcpy(zd, zd1, 2 * n * sizeof(cfx)); memcpy(pd, pd1, 2 * n * sizeof(cfx));
        n = 2 * n;
    }
    return (int)n;
}

int orc_iirdes_dzpk(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as,
                    orc_cf *zd, orc_cf *pd, orc_cf *kd)
{
    cfx z[32], p[32], k;
    int n = iirdes_dzpk(ftype, btype, order, fc, f0, ap, as, z, p, &k);
    if (n < 0) return n;
    for (int i = 0; i < n; i++) { zd[i].re = z[i].r; zd[i].im = z[i].i; pd[i].re = p[i].r; pd[i].im = p[i].i; }
    kd->re = k.r; kd->im = k.i;
    return n;
}

int orc_iirdes_sos(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as,
                   float *B, float *A)
{
    cfx z[32], p[32], k;
    int n = iirdes_dzpk(ftype, btype, order, fc, f0, ap, as, z, p, &k);
    if (n < 0) return n;
    return (int)dzpk2sosf(z, p, (unsigned)n, k, B, A);
}

/* ==========================================================================================
 *  iirfilt_crcf, SOS form     liquid: src/filter/src/iirfilt.proto.c, iirfiltsos.proto.c
 *  reached from: ComplexIIRFilter::execute -> iirfilt_crcf_execute_block (iirfilter.hpp:296)
 * ======================================================================================== */
#define ORC_MAX_SOS 16
struct orc_iirfilt_crcf_s {
    unsigned nsos;
    float b[ORC_MAX_SOS][3], a[ORC_MAX_SOS][3];
    float v[ORC_MAX_SOS][3][2];   /* v[s][k][re/im] */
};

orc_iirfilt_crcf orc_iirfilt_crcf_create_sos(const float *B, const float *A, unsigned nsos)
{
    if (nsos == 0 || nsos > ORC_MAX_SOS) return NULL;
    orc_iirfilt_crcf q = calloc(1, sizeof *q);
    q->nsos = nsos;
    for (unsigned s = 0; s < nsos; s++) {          /* iirfiltsos_create: normalise by a0 */
        float a0 = A[3 * s];
        for (int k = 0; k < 3; k++) { q->b[s][k] = B[3 * s + k] / a0; q->a[s][k] = A[3 * s + k] / a0; }
    }
    return q;
}

orc_iirfilt_crcf orc_iirfilt_crcf_create_prototype(int ftype, int btype, unsigned order,
                                                   float fc, float f0, float ap, float as)
{
    float B[3 * ORC_MAX_SOS], A[3 * ORC_MAX_SOS];
    int nsos = orc_iirdes_sos(ftype, btype, order, fc, f0, ap, as, B, A);
    if (nsos <= 0) return NULL;
    return orc_iirfilt_crcf_create_sos(B, A, (unsigned)nsos);
}

void orc_iirfilt_crcf_destroy(orc_iirfilt_crcf q) { free(q); }
void orc_iirfilt_crcf_reset(orc_iirfilt_crcf q) { memset(q->v, 0, sizeof q->v); }
unsigned orc_iirfilt_crcf_get_sos(orc_iirfilt_crcf q, float *B, float *A)
{
    for (unsigned s = 0; s < q->nsos; s++)
        for (int k = 0; k < 3; k++) { B[3 * s + k] = q->b[s][k]; A[3 * s + k] = q->a[s][k]; }
    return q->nsos;
}

/* iirfiltsos_execute_df2, one real lane (the filter is real-coefficient, so the real and imaginary
 * lanes of the complex sample run the identical recurrence):
 *     v0 = x - a1*v1 - a2*v2        -> fnma(a2, v2, fnma(a1, v1, x))
 *     y  = b0*v0 + b1*v1 + b2*v2    -> fma(b2, v2, fma(b0, v0, b1*v1))                          */
static inline float sos_df2_lane(const float *b, const float *a, float *v0, float *v1, float *v2, float x)
{
    *v2 = *v1; *v1 = *v0;
    float t = FMA(-a[1], *v1, x);
    *v0 = FMA(-a[2], *v2, t);
    float y = b[1] * *v1;
    y = FMA(b[0], *v0, y);
    y = FMA(b[2], *v2, y);
    return y;
}

void orc_iirfilt_crcf_execute_block(orc_iirfilt_crcf q, const orc_cf *x, unsigned n, orc_cf *y)
{
    for (unsigned i = 0; i < n; i++) {
        float tr = x[i].re, ti = x[i].im;
        for (unsigned s = 0; s < q->nsos; s++) {
            tr = sos_df2_lane(q->b[s], q->a[s], &q->v[s][0][0], &q->v[s][1][0], &q->v[s][2][0], tr);
            ti = sos_df2_lane(q->b[s], q->a[s], &q->v[s][0][1], &q->v[s][1][1], &q->v[s][2][1], ti);
        }
        y[i].re = tr; y[i].im = ti;
    }
}

void orc_iirfilt_crcf_execute_block_f64(const float *B, const float *A, unsigned nsos, double *st,
                                        const orc_cf *x, unsigned n, double *y)
{
    for (unsigned i = 0; i < n; i++) {
        double t[2] = { x[i].re, x[i].im };
        for (unsigned s = 0; s < nsos; s++)
            for (int c = 0; c < 2; c++) {
                double *v1 = &st[4 * s + c], *v2 = &st[4 * s + 2 + c];
                double v0 = t[c] - (double)A[3 * s + 1] * *v1 - (double)A[3 * s + 2] * *v2;
                t[c] = (double)B[3 * s] * v0 + (double)B[3 * s + 1] * *v1 + (double)B[3 * s + 2] * *v2;
                *v2 = *v1; *v1 = v0;
            }
        y[2 * i] = t[0]; y[2 * i + 1] = t[1];
    }
}

/* iirfilt_freqresponse, SOS branch: product of section responses evaluated in float complex */
void orc_iirfilt_crcf_freqresponse(orc_iirfilt_crcf q, float fc, orc_cf *H)
{
    cfx h = cmk(1.0f, 0.0f);
    for (unsigned s = 0; s < q->nsos; s++) {
        cfx e1 = cmk(cosf(2 * M_PI * fc * 1), -sinf(2 * M_PI * fc * 1));
        cfx e2 = cmk(cosf(2 * M_PI * fc * 2), -sinf(2 * M_PI * fc * 2));
        cfx hb = cadd(cadd(cmk(q->b[s][0], 0), cscale(e1, q->b[s][1])), cscale(e2, q->b[s][2]));
        cfx ha = cadd(cadd(cmk(q->a[s][0], 0), cscale(e1, q->a[s][1])), cscale(e2, q->a[s][2]));
        h = cmul(h, cdivf(hb, ha));
    }
    H->re = h.r; H->im = h.i;
}

/* ------------------------------------------------------------------------------------------
 *  iirfilt_rrrf, transfer-function form (iirfilt_execute_norm)   liquid: iirfilt.proto.c
 *  reached from: DeemphasisFilter (iirfilter.hpp:371 create, :389 per-sample execute)
 * ---------------------------------------------------------------------------------------- */
struct orc_iirfilt_rrrf_s { unsigned nb, na, n; float b[16], a[16], v[16]; };

orc_iirfilt_rrrf orc_iirfilt_rrrf_create(const float *b, unsigned nb, const float *a, unsigned na)
{
    if (nb == 0 || na == 0 || nb > 16 || na > 16) return NULL;
    orc_iirfilt_rrrf q = calloc(1, sizeof *q);
    q->nb = nb; q->na = na; q->n = nb > na ? nb : na;
    float a0 = a[0];
    for (unsigned i = 0; i < nb; i++) q->b[i] = b[i] / a0;
    for (unsigned i = 0; i < na; i++) q->a[i] = a[i] / a0;
    return q;
}
void orc_iirfilt_rrrf_destroy(orc_iirfilt_rrrf q) { free(q); }
void orc_iirfilt_rrrf_reset(orc_iirfilt_rrrf q) { memset(q->v, 0, sizeof q->v); }

void orc_iirfilt_rrrf_execute(orc_iirfilt_rrrf q, float x, float *y)
{
    unsigned i;
    for (i = q->n - 1; i > 0; i--) q->v[i] = q->v[i - 1];
    float v0 = x;
    for (i = 1; i < q->na; i++) v0 = FMA(-q->a[i], q->v[i], v0);      /* v0 -= a[i]*v[i] */
    q->v[0] = v0;
    float y0 = 0;
    for (i = 0; i < q->nb; i++) y0 = FMA(q->b[i], q->v[i], y0);       /* y0 += b[i]*v[i] */
    *y = y0;
}

void orc_iirfilt_rrrf_freqresponse(orc_iirfilt_rrrf q, float fc, orc_cf *H)
{
    cfx hb = cmk(0, 0), ha = cmk(0, 0);
    for (unsigned i = 0; i < q->nb; i++)
        hb = cadd(hb, cscale(cmk(cosf(2 * M_PI * fc * i), -sinf(2 * M_PI * fc * i)), q->b[i]));
    for (unsigned i = 0; i < q->na; i++)
        ha = cadd(ha, cscale(cmk(cosf(2 * M_PI * fc * i), -sinf(2 * M_PI * fc * i)), q->a[i]));
    cfx h = cdivf(hb, ha);
    H->re = h.r; H->im = h.i;
}

/* ==========================================================================================
 *  firfilt (real taps)   liquid: src/filter/src/firfilt.proto.c, src/buffer/src/window.proto.c,
 *                                src/dotprod/src/dotprod.proto.c (portable C order)
 *  reached from: firfilt_rrrf_execute_block (firfilter.hpp:33), BroadcastAM crcf use
 *  (demod.hpp:105,135-136), and inside ampmodem.
 *  window = double-length buffer so the n most recent samples are contiguous, oldest first.
 * ======================================================================================== */
struct orc_firfilt_s {
    unsigned n, w;          /* length, write index */
    float *hrev;            /* taps reversed (liquid stores them so for dot(h_rev, window)) */
    float *h;               /* design order */
    float *wr, *wi;         /* 2n each */
    float scale;
};

orc_firfilt orc_firfilt_create(const float *h, unsigned n)
{
    if (n == 0) return NULL;
    orc_firfilt q = calloc(1, sizeof *q);
    q->n = n; q->scale = 1.0f;
    q->h = malloc(n * sizeof(float)); q->hrev = malloc(n * sizeof(float));
    for (unsigned i = 0; i < n; i++) { q->h[i] = h[i]; q->hrev[n - 1 - i] = h[i]; }
    q->wr = calloc(2 * n, sizeof(float)); q->wi = calloc(2 * n, sizeof(float));
    return q;
}
orc_firfilt orc_firfilt_create_kaiser(unsigned n, float fc, float as, float mu)
{
    float *h = malloc(n * sizeof(float));
    if (orc_firdes_kaiser(n, fc, as, mu, h) != 0) { free(h); return NULL; }
    orc_firfilt q = orc_firfilt_create(h, n); free(h); return q;
}
orc_firfilt orc_firfilt_create_dc_blocker(unsigned m, float as)
{
    unsigned n = 2 * m + 1; float *h = malloc(n * sizeof(float));
    if (orc_firdes_notch(m, 0.0f, as, h) != 0) { free(h); return NULL; }
    orc_firfilt q = orc_firfilt_create(h, n); free(h); return q;
}
void orc_firfilt_destroy(orc_firfilt q) { if (!q) return; free(q->h); free(q->hrev); free(q->wr); free(q->wi); free(q); }
void orc_firfilt_reset(orc_firfilt q) { memset(q->wr, 0, 2 * q->n * sizeof(float)); memset(q->wi, 0, 2 * q->n * sizeof(float)); q->w = 0; }
void orc_firfilt_set_scale(orc_firfilt q, float s) { q->scale = s; }
unsigned orc_firfilt_get_taps(orc_firfilt q, float *h) { memcpy(h, q->h, q->n * sizeof(float)); return q->n; }

void orc_firfilt_crcf_push(orc_firfilt q, orc_cf x)
{
    q->wr[q->w] = q->wr[q->w + q->n] = x.re;
    q->wi[q->w] = q->wi[q->w + q->n] = x.im;
    q->w = (q->w + 1) % q->n;
}
/* dotprod_crcf_run: r += h[i]*x[i]  ->  per lane r = fma(h[i], x[i], r), i ascending (oldest first) */
void orc_firfilt_crcf_execute(orc_firfilt q, orc_cf *y)
{
    const float *r = q->wr + q->w, *im = q->wi + q->w;
    float sr = 0.0f, si = 0.0f;
    for (unsigned i = 0; i < q->n; i++) { sr = FMA(q->hrev[i], r[i], sr); si = FMA(q->hrev[i], im[i], si); }
    y->re = sr * q->scale; y->im = si * q->scale;
}
void orc_firfilt_crcf_execute_block(orc_firfilt q, const orc_cf *x, unsigned n, orc_cf *y)
{
    for (unsigned i = 0; i < n; i++) { orc_firfilt_crcf_push(q, x[i]); orc_firfilt_crcf_execute(q, &y[i]); }
}
void orc_firfilt_rrrf_push(orc_firfilt q, float x)
{
    q->wr[q->w] = q->wr[q->w + q->n] = x;
    q->w = (q->w + 1) % q->n;
}
void orc_firfilt_rrrf_execute(orc_firfilt q, float *y)
{
    const float *r = q->wr + q->w; float s = 0.0f;
    for (unsigned i = 0; i < q->n; i++) s = FMA(q->hrev[i], r[i], s);
    *y = s * q->scale;
}
void orc_firfilt_rrrf_execute_block(orc_firfilt q, const float *x, unsigned n, float *y)
{
    for (unsigned i = 0; i < n; i++) { orc_firfilt_rrrf_push(q, x[i]); orc_firfilt_rrrf_execute(q, &y[i]); }
}
void orc_firfilt_freqresponse(orc_firfilt q, float fc, orc_cf *H)
{
    cfx h = cmk(0, 0);
    for (unsigned i = 0; i < q->n; i++)
        h = cadd(h, cscale(cmk(cosf(2 * M_PI * fc * i), -sinf(2 * M_PI * fc * i)), q->h[i]));
    H->re = h.r * q->scale; H->im = h.i * q->scale;
}

/* ==========================================================================================
 *  resamp_cccf, fixed-point phase   liquid: src/filter/src/resamp.fixed.proto.c, firpfb.proto.c
 *  reached from: ComplexResampler (resampler.hpp:136 create, :153 set_rate, :165 execute)
 * ======================================================================================== */
struct orc_resamp_s {
    float rate; unsigned m, npfb, bits, sublen;
    uint32_t step, phase;
    float *bank;            /* [npfb][sublen] reversed sub-filters (real part; imaginary part is 0) */
    float *wr, *wi; unsigned w;
    int real_taps;          /* 0: resamp_cccf arithmetic, 1: resamp_crcf / resamp_rrrf */
};

static unsigned nextpow2(unsigned x) { x--; unsigned n = 0; while (x > 0) { x >>= 1; n++; } return n; }

void orc_resamp_set_rate(orc_resamp q, float rate)
{
    q->rate = rate;
    q->step = (uint32_t)round((1 << 24) / q->rate);      /* int / float -> float division */
}

orc_resamp orc_resamp_create(float rate, unsigned m, float fc, float as, unsigned npfb)
{
    if (rate <= 0 || m == 0 || npfb == 0 || fc <= 0 || fc >= 0.5f || as <= 0) return NULL;
    orc_resamp q = calloc(1, sizeof *q);
    orc_resamp_set_rate(q, rate);
    q->m = m;
    q->bits = nextpow2(npfb); q->npfb = 1u << q->bits;
    unsigned n = 2 * q->m * q->npfb + 1, i;
    float *hf = malloc(n * sizeof(float)), *h = malloc(n * sizeof(float));
    if (orc_firdes_kaiser(n, fc / ((float)(q->npfb)), as, 0.0f, hf) != 0) { free(hf); free(h); free(q); return NULL; }
    float gain = 0.0f;
    for (i = 0; i < n; i++) gain += hf[i];
    gain = (q->npfb) / (gain);
    for (i = 0; i < n; i++) h[i] = hf[i] * gain;
    /* firpfb_create(M, h, h_len = n-1): sub-filter i, reversed */
    q->sublen = (n - 1) / q->npfb;
    q->bank = malloc(q->npfb * q->sublen * sizeof(float));
    for (i = 0; i < q->npfb; i++)
        for (unsigned k = 0; k < q->sublen; k++)
            q->bank[i * q->sublen + (q->sublen - k - 1)] = h[i + k * q->npfb];
    q->wr = calloc(2 * q->sublen, sizeof(float)); q->wi = calloc(2 * q->sublen, sizeof(float));
    free(hf); free(h);
    orc_resamp_reset(q);
    return q;
}
void orc_resamp_destroy(orc_resamp q) { if (!q) return; free(q->bank); free(q->wr); free(q->wi); free(q); }
void orc_resamp_reset(orc_resamp q)
{
    memset(q->wr, 0, 2 * q->sublen * sizeof(float)); memset(q->wi, 0, 2 * q->sublen * sizeof(float));
    q->w = 0; q->phase = 0;
}
uint32_t orc_resamp_get_step(orc_resamp q)  { return q->step; }
uint32_t orc_resamp_get_phase(orc_resamp q) { return q->phase; }
unsigned orc_resamp_get_npfb(orc_resamp q)  { return q->npfb; }
unsigned orc_resamp_get_sublen(orc_resamp q){ return q->sublen; }
void orc_resamp_get_bank(orc_resamp q, float *out) { memcpy(out, q->bank, q->npfb * q->sublen * sizeof(float)); }

void orc_resamp_execute(orc_resamp q, orc_cf x, orc_cf *y, unsigned *nw)
{
    q->wr[q->w] = q->wr[q->w + q->sublen] = x.re;           /* firpfb_push */
    q->wi[q->w] = q->wi[q->w + q->sublen] = x.im;
    q->w = (q->w + 1) % q->sublen;
    unsigned n = 0;
    while (q->phase <= 0x00ffffff) {
        unsigned idx = q->phase >> (24 - q->bits);
        const float *h = q->bank + idx * q->sublen, *r = q->wr + q->w, *im = q->wi + q->w;
        /* dotprod_cccf_run with taps (h + 0j): prod = (fma(hr,xr,-(0*xi)), fma(hr,xi,0*xr)); r += prod */
        float sr = 0.0f, si = 0.0f;
        if (q->real_taps) {
            /* resamp_crcf / resamp_rrrf: dotprod_crcf / dotprod_rrrf, a fused multiply-add chain per lane */
            for (unsigned i = 0; i < q->sublen; i++) { sr = FMA(h[i], r[i], sr); si = FMA(h[i], im[i], si); }
        } else {
            for (unsigned i = 0; i < q->sublen; i++) {
                float pr = FMA(h[i], r[i], -(0.0f * im[i]));
                float pi = FMA(h[i], im[i], 0.0f * r[i]);
                sr = sr + pr; si = si + pi;
            }
        }
        y[n].re = sr; y[n].im = si; n++;
        q->phase += q->step;
    }
    q->phase -= (1 << 24);
    *nw = n;
}

/* resamp_crcf / resamp_rrrf arithmetic (CResampler, RResampler, RealResampler: resampler.hpp:4-125); real samples
 * run in the real lane.  create_default: m = 7, fc = min(0.49, rate/2), As = 60 dB, npfb = 64. */
void orc_resamp_set_real_taps(orc_resamp q, int on) { q->real_taps = on; }
orc_resamp orc_resamp_create_default(float rate)
{
    orc_resamp q = orc_resamp_create(rate, 7, 0.5f * rate > 0.49f ? 0.49f : 0.5f * rate, 60.0f, 64);
    if (q) q->real_taps = 1;
    return q;
}
unsigned orc_resamp_execute_block(orc_resamp q, const orc_cf *x, unsigned n, orc_cf *y)
{
    unsigned nw = 0, nd = 0;
    for (unsigned i = 0; i < n; i++) { orc_resamp_execute(q, x[i], &y[nw], &nd); nw += nd; }
    return nw;
}

/* ==========================================================================================
 *  nco_crcf    liquid: src/nco/src/nco.proto.c
 *  reached from: NCO (nco.hpp:18-78) and inside ampmodem
 * ======================================================================================== */
struct orc_nco_s { int type; uint32_t theta, d_theta; float alpha, beta; };
static float g_sintab[1024]; static int g_sintab_ok = 0;
const float *orc_nco_sintab(void)
{
    if (!g_sintab_ok) {
        for (unsigned i = 0; i < 1024; i++) g_sintab[i] = sinf(2.0f * M_PI * (float)(i) / 1024.0f);
        g_sintab_ok = 1;
    }
    return g_sintab;
}
orc_nco orc_nco_create(int type)
{
    orc_nco q = calloc(1, sizeof *q);
    q->type = type; orc_nco_sintab();
    orc_nco_pll_set_bandwidth(q, 0.1f);
    orc_nco_reset(q);
    return q;
}
void orc_nco_destroy(orc_nco q) { free(q); }
void orc_nco_reset(orc_nco q) { q->theta = 0; q->d_theta = 0; }

uint32_t orc_nco_constrain(float theta)
{
    float p = theta * 0.159154943091895;            /* float * double, rounded to float on store */
    float fpart = p - ((long)p);
    if (fpart < 0.) fpart += 1.;
    /* (uint32_t)(fpart * 0xffffffff): 0xffffffff converts to 4294967296.0f; when fpart rounds up to
       1.0f the product is 2^32, out of range for uint32: x86-64 converts through a 64-bit integer and
       keeps the low 32 bits (0). Done explicitly so every platform agrees. */
    float scaled = fpart * 4294967296.0f;
    return (uint32_t)(uint64_t)(int64_t)scaled;
}
void  orc_nco_set_frequency(orc_nco q, float f)     { q->d_theta  = orc_nco_constrain(f); }
void  orc_nco_adjust_frequency(orc_nco q, float df) { q->d_theta += orc_nco_constrain(df); }
void  orc_nco_set_phase(orc_nco q, float phi)       { q->theta    = orc_nco_constrain(phi); }
void  orc_nco_adjust_phase(orc_nco q, float dphi)   { q->theta   += orc_nco_constrain(dphi); }
float orc_nco_get_frequency(orc_nco q)
{
    float d = 2.0f * M_PI * (float)q->d_theta / (float)(0xffffffff);
    return d > M_PI ? d - 2 * M_PI : d;
}
float orc_nco_get_phase(orc_nco q) { return 2.0f * M_PI * (float)q->theta / (float)(0xffffffff); }
uint32_t orc_nco_get_theta_u32(orc_nco q)  { return q->theta; }
uint32_t orc_nco_get_dtheta_u32(orc_nco q) { return q->d_theta; }
void  orc_nco_set_u32(orc_nco q, uint32_t theta, uint32_t d_theta) { q->theta = theta; q->d_theta = d_theta; }
void  orc_nco_pll_set_bandwidth(orc_nco q, float bw) { q->alpha = bw; q->beta = sqrtf(q->alpha); }
void  orc_nco_pll_step(orc_nco q, float dphi)
{
    orc_nco_adjust_frequency(q, dphi * q->alpha);
    orc_nco_adjust_phase(q, dphi * q->beta);
}
void  orc_nco_step(orc_nco q) { q->theta += q->d_theta; }
void  orc_nco_sincos(orc_nco q, float *s, float *c)
{
    if (q->type == ORC_NCO) {
        unsigned idx = ((q->theta + (1u << 21)) >> 22) & 0x3ff;
        *s = g_sintab[idx]; *c = g_sintab[(idx + 256) & 0x3ff];
    } else {
        float th = orc_nco_get_phase(q);
        *s = sinf(th); *c = cosf(th);
    }
}
/* y = x * (c + js):  re = xr*c - xi*s -> fma(xr, c, -(xi*s));  im = xr*s + xi*c -> fma(xi, c, xr*s)
 * y = x * conj(.):   re = fma(xr, c, xi*s);                     im = fma(xi, c, -(xr*s))            */
void orc_nco_mix_up(orc_nco q, orc_cf x, orc_cf *y)
{
    float s, c; orc_nco_sincos(q, &s, &c);
    float re = FMA(x.re, c, -(x.im * s)), im = FMA(x.im, c, x.re * s);
    y->re = re; y->im = im;
}
void orc_nco_mix_down(orc_nco q, orc_cf x, orc_cf *y)
{
    float s, c; orc_nco_sincos(q, &s, &c);
    float re = FMA(x.re, c, x.im * s), im = FMA(x.im, c, -(x.re * s));
    y->re = re; y->im = im;
}
void orc_nco_mix_block_up(orc_nco q, const orc_cf *x, orc_cf *y, unsigned n)
{ for (unsigned i = 0; i < n; i++) { orc_nco_mix_up(q, x[i], &y[i]); orc_nco_step(q); } }
void orc_nco_mix_block_down(orc_nco q, const orc_cf *x, orc_cf *y, unsigned n)
{ for (unsigned i = 0; i < n; i++) { orc_nco_mix_down(q, x[i], &y[i]); orc_nco_step(q); } }

/* ==========================================================================================
 *  agc_crcf    liquid: src/agc/src/agc.proto.c
 *  reached from: AGC (agc.hpp:10-128)
 * ======================================================================================== */
static inline float orc_logf_cr(float x) { return (float)log((double)x); }
static inline float orc_expf_cr(float x) { return (float)exp((double)x); }

struct orc_agc_s {
    float g, scale, bandwidth, alpha, y2_prime;
    int is_locked, squelch_mode;
    float squelch_threshold; unsigned squelch_timeout, squelch_timer;
};
orc_agc orc_agc_create(void)
{
    orc_agc q = calloc(1, sizeof *q);
    orc_agc_set_bandwidth(q, 1e-2f);
    q->scale = 1.0f;
    q->squelch_mode = ORC_AGC_SQUELCH_DISABLED;
    q->squelch_threshold = 0.0f; q->squelch_timeout = 100; q->squelch_timer = 100;
    orc_agc_reset(q);
    return q;
}
void orc_agc_destroy(orc_agc q) { free(q); }
void orc_agc_reset(orc_agc q)
{
    q->g = 1.0f; q->y2_prime = 1.0f; q->is_locked = 0;
    q->squelch_mode = (q->squelch_mode == ORC_AGC_SQUELCH_DISABLED) ? ORC_AGC_SQUELCH_DISABLED : ORC_AGC_SQUELCH_ENABLED;
}
float orc_agc_get_rssi(orc_agc q) { return -20 * log10(q->g); }
static void agc_squelch_update_mode(orc_agc q)
{
    int exceeded = (orc_agc_get_rssi(q) > q->squelch_threshold);
    switch (q->squelch_mode) {
    case ORC_AGC_SQUELCH_ENABLED:  q->squelch_mode = exceeded ? ORC_AGC_SQUELCH_RISE : ORC_AGC_SQUELCH_ENABLED; break;
    case ORC_AGC_SQUELCH_RISE:     q->squelch_mode = exceeded ? ORC_AGC_SQUELCH_SIGNALHI : ORC_AGC_SQUELCH_FALL; break;
    case ORC_AGC_SQUELCH_SIGNALHI: q->squelch_mode = exceeded ? ORC_AGC_SQUELCH_SIGNALHI : ORC_AGC_SQUELCH_FALL; break;
    case ORC_AGC_SQUELCH_FALL:
        q->squelch_timer = q->squelch_timeout;
        q->squelch_mode = exceeded ? ORC_AGC_SQUELCH_SIGNALHI : ORC_AGC_SQUELCH_SIGNALLO; break;
    case ORC_AGC_SQUELCH_SIGNALLO:
        q->squelch_timer--;
        if (q->squelch_timer == 0) q->squelch_mode = ORC_AGC_SQUELCH_TIMEOUT;
        else if (exceeded)         q->squelch_mode = ORC_AGC_SQUELCH_SIGNALHI;
        break;
    case ORC_AGC_SQUELCH_TIMEOUT:  q->squelch_mode = ORC_AGC_SQUELCH_ENABLED; break;
    default: break;   /* DISABLED: stays */
    }
}
void orc_agc_execute(orc_agc q, orc_cf x, orc_cf *y)
{
    float yr = x.re * q->g, yi = x.im * q->g;                       /* y = x*g */
    float y2 = FMA(yr, yr, yi * yi);                                /* crealf(y*conjf(y)) */
    /* (1.0-alpha)*y2_prime + alpha*y2 : the literal 1.0 promotes the first product to double */
    q->y2_prime = (float)FMAD(1.0 - (double)q->alpha, (double)q->y2_prime, (double)(q->alpha * y2));
    if (q->is_locked) { y->re = yr; y->im = yi; return; }           /* returns BEFORE the output scale */
    /* liquid: g *= expf(-0.5f*alpha*logf(y2_prime)).  The last bit of logf/expf differs between C
     * libraries (glibc, Apple libm, ...), so the reference's own result is platform dependent; the
     * oracle takes the correctly rounded functions (double evaluation, one rounding to float). */
    if (q->y2_prime > 1e-6f)
        q->g *= orc_expf_cr(-0.5f * q->alpha * orc_logf_cr(q->y2_prime));
    if (q->g > 1e6f) q->g = 1e6f;
    agc_squelch_update_mode(q);
    y->re = yr * q->scale; y->im = yi * q->scale;
}
void  orc_agc_lock(orc_agc q)   { q->is_locked = 1; }
void  orc_agc_unlock(orc_agc q) { q->is_locked = 0; }
void  orc_agc_set_bandwidth(orc_agc q, float bw) { if (bw < 0 || bw > 1) return; q->bandwidth = bw; q->alpha = q->bandwidth; }
float orc_agc_get_bandwidth(orc_agc q) { return q->bandwidth; }
float orc_agc_get_signal_level(orc_agc q) { return 1.0f / q->g; }
void  orc_agc_set_signal_level(orc_agc q, float x2) { if (x2 <= 0) return; q->g = 1.0f / x2; q->y2_prime = 1.0f; }
void  orc_agc_set_rssi(orc_agc q, float rssi) { q->g = powf(10.0f, -rssi / 20.0f); if (q->g < 1e-16f) q->g = 1e-16f; q->y2_prime = 1.0f; }
float orc_agc_get_gain(orc_agc q) { return q->g; }
void  orc_agc_set_gain(orc_agc q, float g) { if (g <= 0) return; q->g = g; }
float orc_agc_get_scale(orc_agc q) { return q->scale; }
void  orc_agc_set_scale(orc_agc q, float s) { if (s <= 0) return; q->scale = s; }
void  orc_agc_squelch_enable(orc_agc q)  { q->squelch_mode = ORC_AGC_SQUELCH_ENABLED; }
void  orc_agc_squelch_disable(orc_agc q) { q->squelch_mode = ORC_AGC_SQUELCH_DISABLED; }
void  orc_agc_squelch_set_threshold(orc_agc q, float t) { q->squelch_threshold = t; }
float orc_agc_squelch_get_threshold(orc_agc q) { return q->squelch_threshold; }
void  orc_agc_squelch_set_timeout(orc_agc q, unsigned t) { q->squelch_timeout = t; }
int   orc_agc_squelch_get_status(orc_agc q) { return q->squelch_mode; }
float orc_agc_get_y2_prime(orc_agc q) { return q->y2_prime; }

/* reference wrapper loop, agc.hpp:109-128 */
unsigned orc_wrap_agc_execute(orc_agc q, const orc_cf *x, unsigned n, orc_cf *y,
                              int *state_last, unsigned *rise_idx, unsigned rise_cap)
{
    unsigned rises = 0;
    for (unsigned i = 0; i < n; i++) {
        orc_agc_execute(q, x[i], &y[i]);
        int state = orc_agc_squelch_get_status(q);
        if (state != *state_last) {
            *state_last = state;
            if (state == ORC_AGC_SQUELCH_RISE) { if (rise_idx && rises < rise_cap) rise_idx[rises] = i; rises++; }
        }
        if (state == ORC_AGC_SQUELCH_SIGNALLO || state == ORC_AGC_SQUELCH_ENABLED) { y[i].re *= 0.0f; y[i].im *= 0.0f; }
    }
    return rises;
}

/* ==========================================================================================
 *  firhilbf    liquid: src/filter/src/firhilb.proto.c (>= 1.3.2: four windows, toggle, two-output c2r)
 *  reached from: SSBDemod (demod.hpp:163 create(25, 60), :181/:184 c2r_execute), HilbertTransform
 *  (utility.hpp:79-80 create, :93 interp_execute, :101 decim_execute)
 *  Half-band Kaiser prototype of 4m+1 taps, modulated by sin(pi t / 2); the 2m odd-offset taps, reversed,
 *  form the quadrature branch; the in-phase branch is a pure delay.
 * ======================================================================================== */
struct orc_firhilbf_s { unsigned m, L; float *hq; float *w[4]; unsigned wi[4]; unsigned toggle; };

orc_firhilbf orc_firhilbf_create(unsigned m, float as)
{
    if (m < 2) return NULL;
    orc_firhilbf q = calloc(1, sizeof *q);
    unsigned h_len = 4 * m + 1, i, j = 0;
    q->m = m; q->L = 2 * m;
    float *h = malloc(h_len * sizeof(float));
    orc_firdes_kaiser(h_len, 0.25f, fabsf(as), 0.0f, h);
    for (i = 0; i < h_len; i++) {
        float t = (float)i - (float)(h_len - 1) / 2.0f;
        h[i] = h[i] * sinf((float)(0.5f * M_PI * t));          /* imag(h[i] * cexpf(j 0.5 pi t)) */
    }
    q->hq = malloc(q->L * sizeof(float));
    for (i = 1; i < h_len; i += 2) q->hq[j++] = h[h_len - i - 1];
    free(h);
    for (i = 0; i < 4; i++) q->w[i] = calloc(2 * q->L, sizeof(float));
    return q;
}
void orc_firhilbf_destroy(orc_firhilbf q) { if (!q) return; free(q->hq); for (int i = 0; i < 4; i++) free(q->w[i]); free(q); }
void orc_firhilbf_reset(orc_firhilbf q)
{
    for (int i = 0; i < 4; i++) { memset(q->w[i], 0, 2 * q->L * sizeof(float)); q->wi[i] = 0; }
    q->toggle = 0;
}
unsigned orc_firhilbf_get_hq(orc_firhilbf q, float *hq) { memcpy(hq, q->hq, q->L * sizeof(float)); return q->L; }
static void hw_push(orc_firhilbf q, int k, float x) { q->w[k][q->wi[k]] = q->w[k][q->wi[k] + q->L] = x; q->wi[k] = (q->wi[k] + 1) % q->L; }
static const float *hw_read(orc_firhilbf q, int k) { return q->w[k] + q->wi[k]; }     /* oldest first */
static float hw_dot(orc_firhilbf q, const float *r) { float s = 0.0f; for (unsigned i = 0; i < q->L; i++) s = FMA(q->hq[i], r[i], s); return s; }

void orc_firhilbf_r2c_execute(orc_firhilbf q, float x, orc_cf *y)
{
    float yi, yq;
    if (q->toggle == 0) { hw_push(q, 0, x); yi = hw_read(q, 0)[q->m - 1]; yq = hw_dot(q, hw_read(q, 1)); }
    else                { hw_push(q, 1, x); yi = hw_read(q, 1)[q->m - 1]; yq = hw_dot(q, hw_read(q, 0)); }
    q->toggle = 1 - q->toggle;
    y->re = yi; y->im = yq;
}
void orc_firhilbf_c2r_execute(orc_firhilbf q, orc_cf x, float *y_lsb, float *y_usb)
{
    float yi, yq;
    if (q->toggle == 0) { hw_push(q, 0, x.re); hw_push(q, 1, x.im); yi = hw_read(q, 0)[q->m - 1]; yq = hw_dot(q, hw_read(q, 3)); }
    else                { hw_push(q, 2, x.re); hw_push(q, 3, x.im); yi = hw_read(q, 2)[q->m - 1]; yq = hw_dot(q, hw_read(q, 1)); }
    q->toggle = 1 - q->toggle;
    *y_lsb = yi + yq; *y_usb = yi - yq;
}
void orc_firhilbf_decim_execute(orc_firhilbf q, const float *x, orc_cf *y)
{
    float yi, yq;
    hw_push(q, 1, x[0]); yq = hw_dot(q, hw_read(q, 1));
    hw_push(q, 0, x[1]); yi = hw_read(q, 0)[q->m - 1];
    if (q->toggle) { y->re = -yi; y->im = -yq; } else { y->re = yi; y->im = yq; }
    q->toggle = 1 - q->toggle;
}
void orc_firhilbf_interp_execute(orc_firhilbf q, orc_cf x, float *y)
{
    float vr = q->toggle ? -x.re : x.re, vi = q->toggle ? -x.im : x.im;
    hw_push(q, 0, vi); y[0] = hw_read(q, 0)[q->m - 1];
    hw_push(q, 1, vr); y[1] = hw_dot(q, hw_read(q, 1));
    q->toggle = 1 - q->toggle;
}

/* ==========================================================================================
 *  ampmodem (DSB)    liquid: src/modem/src/ampmodem.c, src/buffer/src/wdelay.proto.c
 *  reached from: AmpModem (demod.hpp:294 demodulate_block, :305 create)
 *  constants of ampmodem_create (liquid >= 1.4): m = 25, PLL bandwidth 0.001, lowpass
 *  kaiser(2m+1, 0.01, 40 dB), dc blocker (25, 20 dB), delay m.  The reference author's own clone
 *  of this routine (demod.hpp:101-106, :133-152) corroborates m, the PLL bandwidth and the lowpass.
 * ======================================================================================== */
struct orc_ampmodem_s {
    float mod_index; int type, suppressed; unsigned m;
    orc_nco mixer; orc_firfilt dcblock, lowpass; orc_firhilbf hilbert;     /* hilbert: USB / LSB only */
    float *dr, *di; unsigned dlen, dpos;     /* wdelaycf(m): buffer of m+1 */
};
orc_ampmodem orc_ampmodem_create(float mod_index, int type, int suppressed)
{
    if (type != ORC_AMPMODEM_DSB && type != ORC_AMPMODEM_USB && type != ORC_AMPMODEM_LSB) return NULL;
    orc_ampmodem q = calloc(1, sizeof *q);
    q->mod_index = mod_index; q->type = type; q->suppressed = suppressed; q->m = 25;
    q->mixer = orc_nco_create(ORC_NCO);
    orc_nco_pll_set_bandwidth(q->mixer, 0.001f);
    q->dcblock = orc_firfilt_create_dc_blocker(25, 20.0f);
    q->lowpass = orc_firfilt_create_kaiser(2 * q->m + 1, 0.01f, 40.0f, 0.0f);
    q->dlen = q->m + 1; q->dr = calloc(q->dlen, sizeof(float)); q->di = calloc(q->dlen, sizeof(float));
    q->hilbert = orc_firhilbf_create(q->m, 60.0f);          /* firhilbf_create(m, 60): single side-band recovery */
    return q;
}
void orc_ampmodem_destroy(orc_ampmodem q)
{
    if (!q) return;
    orc_firhilbf_destroy(q->hilbert);
    orc_nco_destroy(q->mixer); orc_firfilt_destroy(q->dcblock); orc_firfilt_destroy(q->lowpass);
    free(q->dr); free(q->di); free(q);
}
void orc_ampmodem_reset(orc_ampmodem q)
{
    orc_nco_reset(q->mixer); orc_firfilt_reset(q->dcblock); orc_firfilt_reset(q->lowpass); orc_firhilbf_reset(q->hilbert);
    memset(q->dr, 0, q->dlen * sizeof(float)); memset(q->di, 0, q->dlen * sizeof(float)); q->dpos = 0;
}
unsigned orc_ampmodem_get_lowpass_taps(orc_ampmodem q, float *h) { return orc_firfilt_get_taps(q->lowpass, h); }
unsigned orc_ampmodem_get_dcblock_taps(orc_ampmodem q, float *h) { return orc_firfilt_get_taps(q->dcblock, h); }
void orc_ampmodem_get_nco(orc_ampmodem q, uint32_t *t, uint32_t *d) { *t = q->mixer->theta; *d = q->mixer->d_theta; }

static float ampmodem_demod_dsb_pll_carrier(orc_ampmodem q, orc_cf x)
{
    orc_cf x0, x1, v0, v1; float y;
    orc_firfilt_crcf_push(q->lowpass, x); orc_firfilt_crcf_execute(q->lowpass, &x0);
    q->dr[q->dpos] = x.re; q->di[q->dpos] = x.im; q->dpos = (q->dpos + 1) % q->dlen;   /* wdelay push */
    x1.re = q->dr[q->dpos]; x1.im = q->di[q->dpos];                                     /* wdelay read: x[n-m] */
    orc_nco_mix_down(q->mixer, x0, &v0);
    orc_nco_mix_down(q->mixer, x1, &v1);
    float phase_error = v0.im;
    orc_nco_pll_step(q->mixer, phase_error);
    orc_nco_step(q->mixer);
    float m = v1.re / q->mod_index;
    orc_firfilt_rrrf_push(q->dcblock, m); orc_firfilt_rrrf_execute(q->dcblock, &y);
    return y;
}
static float ampmodem_demod_dsb_pll_costas(orc_ampmodem q, orc_cf x)
{
    orc_cf v;
    orc_nco_mix_down(q->mixer, x, &v);
    float phase_error = v.im * (v.re > 0 ? 1 : -1);
    orc_nco_pll_step(q->mixer, phase_error);
    orc_nco_step(q->mixer);
    return v.re / q->mod_index;
}
/* ampmodem_demod_ssb: Hilbert pair, then 0.5 * side-band / mod_index */
static float ampmodem_demod_ssb(orc_ampmodem q, orc_cf x)
{
    float m_lsb, m_usb;
    orc_firhilbf_c2r_execute(q->hilbert, x, &m_lsb, &m_usb);
    return 0.5f * (q->type == ORC_AMPMODEM_USB ? m_usb : m_lsb) / q->mod_index;
}
/* ampmodem_demod_ssb_pll_carrier: the DSB carrier loop, then the Hilbert pair on the mixed-down delayed branch,
 * then the DC blocker */
static float ampmodem_demod_ssb_pll_carrier(orc_ampmodem q, orc_cf x)
{
    orc_cf x0, x1, v0, v1; float y, m_lsb, m_usb;
    orc_firfilt_crcf_push(q->lowpass, x); orc_firfilt_crcf_execute(q->lowpass, &x0);
    q->dr[q->dpos] = x.re; q->di[q->dpos] = x.im; q->dpos = (q->dpos + 1) % q->dlen;
    x1.re = q->dr[q->dpos]; x1.im = q->di[q->dpos];
    orc_nco_mix_down(q->mixer, x0, &v0);
    orc_nco_mix_down(q->mixer, x1, &v1);
    orc_nco_pll_step(q->mixer, v0.im);
    orc_nco_step(q->mixer);
    orc_firhilbf_c2r_execute(q->hilbert, v1, &m_lsb, &m_usb);
    float m = 0.5f * (q->type == ORC_AMPMODEM_USB ? m_usb : m_lsb) / q->mod_index;
    orc_firfilt_rrrf_push(q->dcblock, m); orc_firfilt_rrrf_execute(q->dcblock, &y);
    return y;
}
void orc_ampmodem_demodulate_block(orc_ampmodem q, const orc_cf *x, unsigned n, float *y)
{
    if (q->type != ORC_AMPMODEM_DSB) {
        for (unsigned i = 0; i < n; i++) y[i] = q->suppressed ? ampmodem_demod_ssb(q, x[i]) : ampmodem_demod_ssb_pll_carrier(q, x[i]);
        return;
    }
    for (unsigned i = 0; i < n; i++)
        y[i] = q->suppressed ? ampmodem_demod_dsb_pll_costas(q, x[i]) : ampmodem_demod_dsb_pll_carrier(q, x[i]);
}

/* ==========================================================================================
 *  freqdem    liquid: src/modem/src/freqdem.c ; reached from FreqDem (demod.hpp:197,216)
 * ======================================================================================== */
struct orc_freqdem_s { float kf, ref, rr, ri; };
orc_freqdem orc_freqdem_create(float kf)
{
    if (kf <= 0.0f) return NULL;
    orc_freqdem q = calloc(1, sizeof *q);
    q->kf = kf; q->ref = 1.0f / (2 * M_PI * q->kf);
    return q;
}
void orc_freqdem_destroy(orc_freqdem q) { free(q); }
void orc_freqdem_reset(orc_freqdem q) { q->rr = 0; q->ri = 0; }
void orc_freqdem_demodulate_block(orc_freqdem q, const orc_cf *x, unsigned n, float *y)
{
    for (unsigned i = 0; i < n; i++) {
        /* conjf(r_prime)*r : re = a*c + b*d -> fma(a, c, b*d); im = a*d - b*c -> fma(a, d, -(b*c)) */
        float a = q->rr, b = q->ri, c = x[i].re, d = x[i].im;
        float re = FMA(a, c, b * d), im = FMA(a, d, -(b * c));
        y[i] = atan2f(im, re) * q->ref;
        q->rr = c; q->ri = d;
    }
}

/* ==========================================================================================
 *  wrapper-level arithmetic of the reference itself
 * ======================================================================================== */
/* DeemphasisFilter ctor, iirfilter.hpp:366-371: float x = exp(-1.0/(75.0E-6 * sr)) */
void orc_wrap_deemph_coeffs(float sr, float *b0, float *a1)
{
    float x = exp(-1.0 / (75.0E-6 * sr));
    *a1 = -x; *b0 = 1.0 - x;
}
orc_iirfilt_rrrf orc_wrap_deemph_create(float sr)
{
    float b[1], a[2]; a[0] = 1.0f;
    orc_wrap_deemph_coeffs(sr, &b[0], &a[1]);
    return orc_iirfilt_rrrf_create(b, 1, a, 2);
}
void orc_wrap_deemph_execute(orc_iirfilt_rrrf q, const float *x, unsigned n, float *y)
{ for (unsigned i = 0; i < n; i++) orc_iirfilt_rrrf_execute(q, x[i], &y[i]); }

/* ==========================================================================================
 *  BroadcastAM, demod.hpp:94-153 (the reference author's own demodulator; all arithmetic is liquid's)
 *  create :100-108, reset :117-122, demod_one :133-152.  arg() (std::arg -> atan2f) is taken correctly
 *  rounded, for the same reason as the AGC's logf/expf: its last bit differs between math libraries.
 *  The rrrf SOS high-pass is the real lane of the crcf section arithmetic.
 * ======================================================================================== */
struct orc_bam_s {
    unsigned m; orc_nco mixer; orc_firfilt lowpass; orc_iirfilt_crcf dcblock;
    orc_cf *d; unsigned dlen, dpos;          /* wdelaycf(m): m+1 slots */
};
orc_bam orc_wrap_bam_create(int m)
{
    if (m < 1) return NULL;
    orc_bam q = calloc(1, sizeof *q);
    q->m = (unsigned)m;
    q->mixer = orc_nco_create(ORC_NCO);
    orc_nco_pll_set_bandwidth(q->mixer, 0.001f);
    q->lowpass = orc_firfilt_create_kaiser(2 * q->m + 1, 0.01f, 40.0f, 0.0f);
    q->dcblock = orc_iirfilt_crcf_create_prototype(ORC_IIRDES_CHEBY2, ORC_IIRDES_HIGHPASS, 3, 20.0f / 48000.0f, 0.0f, 0.5f, 20.0f);
    q->dlen = q->m + 1; q->d = calloc(q->dlen, sizeof(orc_cf));
    return q;
}
void orc_wrap_bam_destroy(orc_bam q)
{
    if (!q) return;
    orc_nco_destroy(q->mixer); orc_firfilt_destroy(q->lowpass); orc_iirfilt_crcf_destroy(q->dcblock); free(q->d); free(q);
}
void orc_wrap_bam_reset(orc_bam q)
{
    orc_nco_reset(q->mixer); orc_firfilt_reset(q->lowpass); orc_iirfilt_crcf_reset(q->dcblock);
    memset(q->d, 0, q->dlen * sizeof(orc_cf)); q->dpos = 0;
}
void orc_wrap_bam_get_nco(orc_bam q, uint32_t *t, uint32_t *d) { *t = q->mixer->theta; *d = q->mixer->d_theta; }
unsigned orc_wrap_bam_get_design(orc_bam q, float *lp, float *B, float *A)
{
    orc_iirfilt_crcf_get_sos(q->dcblock, B, A);
    return orc_firfilt_get_taps(q->lowpass, lp);
}
/* test hook: run with externally supplied DC-block sections (parity tests hand over the product's design, so
 * the comparison is of the arithmetic, not of two float evaluations of a pole 2.6e-3 away from z = 1) */
void orc_wrap_bam_set_dcblock(orc_bam q, const float *B, const float *A, unsigned nsos)
{
    orc_iirfilt_crcf_destroy(q->dcblock);
    q->dcblock = orc_iirfilt_crcf_create_sos(B, A, nsos);
}
void orc_wrap_bam_execute(orc_bam q, const orc_cf *x, unsigned n, float *y)
{
    for (unsigned i = 0; i < n; i++) {
        orc_cf x0, x1, v0, v1, in, out;
        orc_firfilt_crcf_push(q->lowpass, x[i]); orc_firfilt_crcf_execute(q->lowpass, &x0);
        q->d[q->dpos] = x[i]; q->dpos = (q->dpos + 1) % q->dlen;        /* wdelaycf_push */
        x1 = q->d[q->dpos];                                             /* wdelaycf_read: x[n-m] */
        orc_nco_mix_down(q->mixer, x0, &v0);
        orc_nco_mix_down(q->mixer, x1, &v1);
        float phase_error = (float)atan2((double)v0.im, (double)v0.re);
        orc_nco_pll_step(q->mixer, phase_error);
        orc_nco_step(q->mixer);
        in.re = v1.re; in.im = 0.0f;
        orc_iirfilt_crcf_execute_block(q->dcblock, &in, 1, &out);
        y[i] = out.re;
    }
}

/* SSBDemod::execute, demod.hpp:172-186 */
void orc_wrap_ssb_execute(orc_firhilbf q, int usb, const orc_cf *x, unsigned n, float *y)
{
    float lsb_, usb_;
    for (unsigned i = 0; i < n; i++) { orc_firhilbf_c2r_execute(q, x[i], &lsb_, &usb_); y[i] = usb ? usb_ : lsb_; }
}
/* HilbertTransform::call, complex64 branch, utility.hpp:88-95: interp_execute(z[n], &y[n]) writes y[n] and y[n+1];
 * every y[n+1] is overwritten by the next iteration, so y[n] keeps the delay-branch output.  The reference's last
 * iteration writes y[len] past the end of the array; that write is dropped here. */
void orc_wrap_hilbert_c2r(orc_firhilbf q, const orc_cf *z, unsigned n, float *y)
{
    float out[2];
    for (unsigned i = 0; i < n; i++) { orc_firhilbf_interp_execute(q, z[i], out); y[i] = out[0]; }
}
/* float32 branch, utility.hpp:96-103: decim_execute(&z[n], &y[n]) reads z[n] and z[n+1] (overlapping pairs); the last
 * iteration reads z[len] past the end of the array -- undefined in the reference, taken as 0.0f here. */
void orc_wrap_hilbert_r2c(orc_firhilbf q, const float *z, unsigned n, orc_cf *y)
{
    float pair[2];
    for (unsigned i = 0; i < n; i++) { pair[0] = z[i]; pair[1] = i + 1 < n ? z[i + 1] : 0.0f; orc_firhilbf_decim_execute(q, pair, &y[i]); }
}

/* ==========================================================================================
 *  FMStereo, demod.hpp:4-85.  create :16-33, reset :35-38 (the resamplers only), demod_one :56-84, execute :40-54.
 *  As written in the reference: the mixer's frequency is never set (it starts at 0 and only the PLL moves it), the PLL
 *  keeps nco_crcf_create's default bandwidth (0.1), and phase_error is an uninitialised member -- taken as 0 here.
 *  arg() and the discriminator's cargf are taken correctly rounded (see BroadcastAM); the one-pole phase filter is
 *  the double expression 0.999*pe + 0.001*arg with GCC's contraction, narrowed to float on assignment.
 * ======================================================================================== */
struct orc_fmstereo_s {
    orc_nco mixer; float ref, rr, ri; orc_iirfilt_rrrf emphL, emphR; orc_resamp audioL, audioR; float phase_error;
};
orc_fmstereo orc_wrap_fmstereo_create(float iq_rate, float pcm_rate)
{
    orc_fmstereo q = calloc(1, sizeof *q);
    float mB[1], mA[2];
    mA[0] = 1.0;
    mA[1] = -exp(-1.0 / (75.0E-6 * iq_rate));
    mB[0] = 1.0 + mA[1];
    q->mixer = orc_nco_create(ORC_NCO);
    q->ref = 1.0f / (2 * M_PI * 4.0f);                       /* freqdem_create(4.0): ref = 1/(2 pi kf) */
    q->emphL = orc_iirfilt_rrrf_create(mB, 1, mA, 2);
    q->emphR = orc_iirfilt_rrrf_create(mB, 1, mA, 2);
    q->audioL = orc_resamp_create_default(pcm_rate / iq_rate);
    q->audioR = orc_resamp_create_default(pcm_rate / iq_rate);
    if (!q->audioL || !q->audioR) { orc_wrap_fmstereo_destroy(q); return NULL; }
    return q;
}
void orc_wrap_fmstereo_destroy(orc_fmstereo q)
{
    if (!q) return;
    orc_nco_destroy(q->mixer); orc_iirfilt_rrrf_destroy(q->emphL); orc_iirfilt_rrrf_destroy(q->emphR);
    orc_resamp_destroy(q->audioL); orc_resamp_destroy(q->audioR); free(q);
}
void orc_wrap_fmstereo_reset(orc_fmstereo q) { orc_resamp_reset(q->audioL); orc_resamp_reset(q->audioR); }
void orc_wrap_fmstereo_get_state(orc_fmstereo q, uint32_t *theta, uint32_t *d_theta, float *pe)
{ *theta = q->mixer->theta; *d_theta = q->mixer->d_theta; *pe = q->phase_error; }
void orc_wrap_fmstereo_get_deemph(orc_fmstereo q, float *b0, float *a1) { *b0 = q->emphL->b[0]; *a1 = q->emphL->a[1]; }
unsigned orc_wrap_fmstereo_execute(orc_fmstereo q, const orc_cf *x, unsigned n, float *y)
{
    unsigned nw = 0;
    for (unsigned i = 0; i < n; i++) {
        /* freqdem_demodulate: cargf(conjf(r') * r) * ref */
        float a = q->rr, b = q->ri, c = x[i].re, d = x[i].im;
        float re = FMA(a, c, b * d), im = FMA(a, d, -(b * c));
        float s = (float)atan2((double)im, (double)re) * q->ref;
        q->rr = c; q->ri = d;
        orc_cf in = { s, 0.0f }, sc;
        orc_nco_mix_down(q->mixer, in, &sc);                                  /* down by the pilot */
        float arg = (float)atan2((double)sc.im, (double)sc.re);
        q->phase_error = (float)fma(0.999, (double)q->phase_error, 0.001 * (double)arg);
        orc_nco_mix_down(q->mixer, sc, &sc);                                  /* and once more */
        orc_nco_pll_step(q->mixer, q->phase_error);
        orc_nco_step(q->mixer);
        float left, right; orc_cf ol[256], orr[256]; unsigned nl, nr;      /* (liquid bounds the rate by 250) */
        orc_iirfilt_rrrf_execute(q->emphL, s + sc.re, &left);
        orc_iirfilt_rrrf_execute(q->emphR, s - sc.re, &right);
        orc_cf li = { left, 0.0f };
        orc_resamp_execute(q->audioL, li, ol, &nl);
        /* demod_one (:79-83) hands the resamplers left = &y[nw] and right = &y[nw + 1] as both input and output: when the
         * left resampler yields two or more samples (pcm_rate > iq_rate) its second one lands on *right before the right
         * resampler reads it.  Such steps are dropped by execute() (nl + nr != 2), but the right resampler's window keeps
         * the overwritten value. */
        orc_cf ri = { nl >= 2 ? ol[1].re : right, 0.0f };
        orc_resamp_execute(q->audioR, ri, orr, &nr);
        if (nl + nr == 2) { y[nw] = ol[0].re; y[nw + 1] = orr[0].re; nw += 2; }   /* execute(), :45-48 */
    }
    return nw;
}

/* bytes_to_iq, utility.hpp:61-69 */
void orc_wrap_bytes_to_iq(const int16_t *iq, unsigned n, orc_cf *y)
{ for (unsigned i = 0; i < n; i++) { y[i].re = (float)iq[2 * i] / 32767.0f; y[i].im = (float)iq[2 * i + 1] / 32767.0f; } }

/* ==========================================================================================
 *  README AMRadio (README.md:41-58): the north-star chain as one object (CPU baseline)
 * ======================================================================================== */
struct orc_amradio_s {
    orc_iirfilt_crcf bandpass; orc_resamp resample; orc_agc agc; orc_ampmodem am; orc_iirfilt_rrrf deemph;
    int agc_state_last; orc_cf *t0, *t1; float *t2; unsigned cap;
};
orc_amradio orc_amradio_create(float bandwidth, float iq_rate, float pcm_rate)
{
    orc_amradio q = calloc(1, sizeof *q);
    q->bandpass = orc_iirfilt_crcf_create_prototype(ORC_IIRDES_CHEBY2, ORC_IIRDES_LOWPASS, 8,
                                                    bandwidth / iq_rate, 0.3f, 0.7f, 60.0f);
    q->resample = orc_resamp_create(pcm_rate / iq_rate, 20, pcm_rate / iq_rate, 60.0f, 13);
    q->am = orc_ampmodem_create(0.5f, ORC_AMPMODEM_DSB, 0);
    q->deemph = orc_wrap_deemph_create(pcm_rate);
    q->agc = orc_agc_create();
    orc_agc_unlock(q->agc); orc_agc_set_scale(q->agc, 0.01f);
    q->agc_state_last = ORC_AGC_SQUELCH_UNKNOWN;
    q->cap = 0;
    return q;
}
void orc_amradio_destroy(orc_amradio q)
{
    if (!q) return;
    orc_iirfilt_crcf_destroy(q->bandpass); orc_resamp_destroy(q->resample); orc_agc_destroy(q->agc);
    orc_ampmodem_destroy(q->am); orc_iirfilt_rrrf_destroy(q->deemph);
    free(q->t0); free(q->t1); free(q->t2); free(q);
}
unsigned orc_amradio_execute(orc_amradio q, const orc_cf *iq, unsigned n, float *pcm)
{
    if (n > q->cap) {
        free(q->t0); free(q->t1); free(q->t2);
        q->t0 = malloc(n * sizeof(orc_cf)); q->t1 = malloc(n * sizeof(orc_cf)); q->t2 = malloc(n * sizeof(float));
        q->cap = n;
    }
    orc_iirfilt_crcf_execute_block(q->bandpass, iq, n, q->t0);
    unsigned nw = orc_resamp_execute_block(q->resample, q->t0, n, q->t1);
    orc_wrap_agc_execute(q->agc, q->t1, nw, q->t0, &q->agc_state_last, NULL, 0);
    orc_ampmodem_demodulate_block(q->am, q->t0, nw, q->t2);
    orc_wrap_deemph_execute(q->deemph, q->t2, nw, pcm);
    return nw;
}
