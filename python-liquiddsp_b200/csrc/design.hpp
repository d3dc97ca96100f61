// design.hpp -- host-side filter / oscillator design for the B200 baseband chain.
//
// The reference designs nothing itself: ComplexIIRFilter's constructor calls
// iirfilt_crcf_create_prototype (/root/reference/src/iirfilter.hpp:275), ComplexResampler calls
// resamp_cccf_create (resampler.hpp:136), AmpModem calls ampmodem_create (demod.hpp:305), which
// run liquid-dsp's float-precision design routines.  Coefficient bits drive every sample that
// follows, so the design runs on the host in float arithmetic the way liquid evaluates it
// (double only where liquid's C expressions promote), and the resulting tables are uploaded once.
//
// The one formula the reference does own -- the de-emphasis pole, iirfilter.hpp:366-370 -- is
// deemph_coeffs() below.
#pragma once
#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>
#include <algorithm>

namespace lqb {
namespace design {

using cplx = std::complex<float>;
static constexpr double kPi = 3.14159265358979323846;

// ------------------------------------------------------------------ windows / FIR prototypes
inline float lngammaf(float z)
{
    if (z < 0.f) return NAN;
    // ln G(z) = ln G(z+k) - sum_i ln(z+i); the logs come off largest argument first
    std::vector<float> logs;
    while (z < 10.0f) { logs.push_back(std::log(z)); z += 1.0f; }
    float g = (float)(0.5 * (std::log((float)(2 * kPi)) - std::log((double)z)));
    g += z * (std::log(z + (1 / (12.0f * z - 0.1f / z))) - 1);
    for (size_t i = logs.size(); i-- > 0;) g = g - logs[i];
    return g;
}

inline float besseli0f(float z)
{
    if (z == 0.0f) return 1.0f;
    float y = 0.0f;
    const float lz = std::log(0.5f * z);
    for (unsigned k = 0; k < 32; k++) {
        float t = (float)k * lz - lngammaf((float)k + 1.0f);
        y += std::exp(2 * t);
    }
    return y;
}

inline float kaiser_beta(float as)
{
    as = std::fabs(as);
    if (as > 50.0f) return 0.1102f * (as - 8.7f);
    if (as > 21.0f) return (float)(0.5842 * std::pow(as - 21, 0.4f) + 0.07886f * (as - 21));
    return 0.0f;
}

inline float kaiser_window(unsigned i, unsigned wlen, float beta)
{
    float t = (float)i - (float)(wlen - 1) / 2;
    float r = 2.0f * t / (float)(wlen - 1);
    return besseli0f(beta * std::sqrt(1 - r * r)) / besseli0f(beta);
}

inline float sincf(float x)
{
    if (std::fabs(x) < 0.01f)
        return std::cos((float)(kPi * x / 2.0f)) * std::cos((float)(kPi * x / 4.0f)) * std::cos((float)(kPi * x / 8.0f));
    return (float)(std::sin((float)(kPi * x)) / (kPi * x));
}

// Kaiser-windowed sinc low-pass, cutoff fc (cycles/sample), stop-band as dB, fractional offset mu.
inline bool firdes_kaiser(unsigned n, float fc, float as, float mu, std::vector<float> &h)
{
    if (n == 0 || fc <= 0.f || fc > 0.5f || mu < -0.5f || mu > 0.5f) return false;
    const float beta = kaiser_beta(as);
    h.resize(n);
    for (unsigned i = 0; i < n; i++) {
        float t = (float)i - (float)(n - 1) / 2 + mu;
        h[i] = sincf(2.0f * fc * t) * kaiser_window(i, n, beta);
    }
    return true;
}

// Notch at f0 of semi-length m; f0 = 0 is the DC blocker ampmodem uses.
inline bool firdes_notch(unsigned m, float f0, float as, std::vector<float> &h)
{
    if (m < 1 || m > 1000 || f0 < -0.5f || f0 > 0.5f || as <= 0.f) return false;
    const float beta = kaiser_beta(as);
    const unsigned n = 2 * m + 1;
    h.resize(n);
    float scale = 0.0f;
    for (unsigned i = 0; i < n; i++) {
        float p = -std::cos((float)(2.0f * kPi * f0 * ((float)i - (float)m)));
        float w = kaiser_window(i, n, beta);
        h[i] = p * w;
        scale += h[i] * p;
    }
    for (auto &v : h) v /= scale;
    h[m] += 1.0f;
    return true;
}

// ------------------------------------------------------------------ polyphase resampler bank
struct ResampDesign {
    unsigned bits = 0, npfb = 0, sublen = 0;
    std::vector<float> bank;   // [npfb][sublen], each sub-filter reversed (oldest sample first)
};

// firhilbf_create(m, As): quadrature-branch taps.  Half-band Kaiser prototype of 4m+1 taps times sin(pi t / 2); the
// 2m taps at odd offsets, in reverse order (liquid firhilb.proto.c)
inline bool firhilb_hq(unsigned m, float as, std::vector<float> &hq)
{
    if (m < 2) return false;
    const unsigned h_len = 4 * m + 1;
    std::vector<float> h;
    if (!firdes_kaiser(h_len, 0.25f, std::fabs(as), 0.0f, h)) return false;
    for (unsigned i = 0; i < h_len; i++) {
        const float t = (float)i - (float)(h_len - 1) / 2.0f;
        h[i] = h[i] * std::sin((float)(0.5f * kPi * t));
    }
    hq.clear();
    for (unsigned i = 1; i < h_len; i += 2) hq.push_back(h[h_len - i - 1]);
    return true;
}

inline unsigned nextpow2(unsigned x) { x--; unsigned n = 0; while (x > 0) { x >>= 1; n++; } return n; }

inline uint32_t resamp_step(float rate) { return (uint32_t)std::round((float)(1 << 24) / rate); }

inline bool resamp_design(unsigned m, float fc, float as, unsigned npfb_req, ResampDesign &d)
{
    if (m == 0 || npfb_req == 0 || fc <= 0.f || fc >= 0.5f || as <= 0.f) return false;
    d.bits = nextpow2(npfb_req);
    d.npfb = 1u << d.bits;
    const unsigned n = 2 * m * d.npfb + 1;
    std::vector<float> hf;
    if (!firdes_kaiser(n, fc / (float)d.npfb, as, 0.0f, hf)) return false;
    float gain = 0.0f;
    for (float v : hf) gain += v;
    gain = (float)d.npfb / gain;
    d.sublen = (n - 1) / d.npfb;
    d.bank.assign((size_t)d.npfb * d.sublen, 0.f);
    for (unsigned i = 0; i < d.npfb; i++)
        for (unsigned k = 0; k < d.sublen; k++)
            d.bank[(size_t)i * d.sublen + (d.sublen - 1 - k)] = hf[i + k * d.npfb] * gain;
    return true;
}

// ------------------------------------------------------------------ IIR prototypes -> SOS
enum { BUTTER = 0, CHEBY1, CHEBY2, ELLIP, BESSEL };
enum { LOWPASS = 0, HIGHPASS, BANDPASS, BANDSTOP };

struct Zpk { std::vector<cplx> z, p; cplx k{1.f, 0.f}; };

inline float ellipse_axes(float eps, unsigned n, float &a, float &b)
{
    float t0 = (float)std::sqrt(1.0 + 1.0 / ((double)(eps * eps)));
    float tp = std::pow((float)(t0 + 1.0 / eps), (float)(1.0 / (float)n));
    float tm = std::pow((float)(t0 - 1.0 / eps), (float)(1.0 / (float)n));
    b = (float)(0.5 * (tp + tm));
    a = (float)(0.5 * (tp - tm));
    return t0;
}

inline float pole_angle(unsigned i, unsigned n) { return (float)((float)(2 * (i + 1) + n - 1) * kPi / (float)(2 * n)); }

// ---- elliptic prototype (liquid ellip.c, after Orfanidis): Landen sequences of 7 terms, float throughout
namespace ellip {
constexpr unsigned kIter = 7;
inline void landen(float k, float (&v)[kIter])
{
    for (unsigned i = 0; i < kIter; i++) { const float kp = std::sqrt(1 - k * k); k = (1 - kp) / (1 + kp); v[i] = k; }
}
inline float quarter_period(float k, float kc)      // K(k) given the complementary modulus kc; near k = 1 the series in kc
{
    const float kmin = 4e-4f, kmax = std::sqrt(1 - kmin * kmin);
    if (k > kmax) { const float L = -std::log(0.25f * kc); return L + 0.25f * (L - 1) * kc * kc; }
    float v[kIter]; landen(k, v);
    float K = (float)(kPi * 0.5f);
    for (float vi : v) K *= (1 + vi);
    return K;
}
inline float degree(float N, float k1)              // modulus k solving the degree equation N K'/K = K1'/K1
{
    const float k1c = std::sqrt(1 - k1 * k1);
    const float K1 = quarter_period(k1, k1c), K1p = quarter_period(k1c, k1);
    const float q1 = std::exp((float)(-kPi * K1p / K1)), q = std::pow(q1, 1.0f / N);
    float b = 0.f, a = 0.f;
    for (unsigned m = 0; m < kIter; m++) b += std::pow(q, (float)(m * (m + 1)));
    for (unsigned m = 1; m < kIter; m++) a += std::pow(q, (float)(m * m));
    const float g = b / (1.0f + 2.0f * a);
    return 4.0f * std::sqrt(q) * g * g;
}
inline cplx ascend(cplx w, float k)                 // Gauss transformation back up the Landen sequence
{
    float v[kIter]; landen(k, v);
    for (unsigned i = kIter; i > 0; i--) w = (1 + v[i - 1]) * w / (cplx(1.f, 0.f) + v[i - 1] * w * w);
    return w;
}
inline cplx cd(cplx u, float k) { return ascend(std::cos(u * (float)(kPi * 0.5)), k); }
inline cplx sn(cplx u, float k) { return ascend(std::sin(u * (float)(kPi * 0.5)), k); }
inline cplx asn(cplx w, float k)
{
    float v[kIter]; landen(k, v);
    for (unsigned i = 0; i < kIter; i++) {
        const float v1 = i == 0 ? k : v[i - 1];
        w = w / (cplx(1.f, 0.f) + std::sqrt(cplx(1.f, 0.f) - w * w * v1 * v1)) * 2.0f / (1 + v[i]);
    }
    return cplx(1.f, 0.f) - std::acos(w) * 2.0f / (float)kPi;
}
}  // namespace ellip

// reverse Bessel polynomial roots (the delay-normalised Bessel poles), Durand-Kerner in double
inline std::vector<std::complex<double>> bessel_roots(unsigned n)
{
    std::vector<double> c(n + 1);
    for (unsigned k = 0; k <= n; k++)
        c[k] = std::exp(std::lgamma(2.0 * n - k + 1) - std::lgamma((double)k + 1) - std::lgamma((double)(n - k) + 1) - (double)(n - k) * std::log(2.0));
    std::vector<std::complex<double>> z(n);
    for (unsigned i = 0; i < n; i++) z[i] = std::pow(std::complex<double>(0.4, 0.9), (double)i) * (double)n;
    for (int it = 0; it < 500; it++) {
        double worst = 0;
        for (unsigned i = 0; i < n; i++) {
            std::complex<double> num = 1.0, den = 1.0;
            for (unsigned k = n; k-- > 0;) num = num * z[i] + c[k];
            for (unsigned k = 0; k < n; k++) if (k != i) den *= (z[i] - z[k]);
            const std::complex<double> d = num / den;
            z[i] -= d; worst = std::max(worst, std::abs(d));
        }
        if (worst < 1e-14 * n) break;
    }
    return z;
}

// analog prototype; returns false for an unknown family
inline bool analog_prototype(int ftype, unsigned n, float ap, float as, Zpk &A, cplx &k0)
{
    const unsigned r = n % 2, L = (n - r) / 2;
    A.z.clear(); A.p.clear(); k0 = cplx(1.f, 0.f);
    if (ftype == BUTTER) {
        for (unsigned i = 0; i < L; i++) {
            float th = pole_angle(i, n);
            A.p.emplace_back(std::cos(th),  std::sin(th));
            A.p.emplace_back(std::cos(th), -std::sin(th));
        }
        if (r) A.p.emplace_back(-1.f, 0.f);
        return true;
    }
    if (ftype == CHEBY1) {
        float eps = std::sqrt(std::pow(10.0f, ap / 10.0f) - 1.0f), a, b;
        k0 = cplx(r ? 1.0f : 1.0f / std::sqrt(1.0f + eps * eps), 0.f);
        ellipse_axes(eps, n, a, b);
        for (unsigned i = 0; i < L; i++) {
            float th = pole_angle(i, n);
            A.p.emplace_back(a * std::cos(th), -(b * std::sin(th)));
            A.p.emplace_back(a * std::cos(th),   b * std::sin(th));
        }
        if (r) A.p.emplace_back(-a, 0.f);
        return true;
    }
    if (ftype == CHEBY2) {
        float eps = std::pow(10.0f, -as / 20.0f), a, b;
        ellipse_axes(eps, n, a, b);
        for (unsigned i = 0; i < L; i++) {
            float th = pole_angle(i, n);
            A.p.push_back(cplx(1.f, 0.f) / cplx(a * std::cos(th), -(b * std::sin(th))));
            A.p.push_back(cplx(1.f, 0.f) / cplx(a * std::cos(th),   b * std::sin(th)));
        }
        if (r) A.p.emplace_back(-1.0f / a, 0.f);
        for (unsigned i = 0; i < L; i++) {
            float th = (float)(0.5f * kPi * (2 * (i + 1) - 1) / (float)n);
            A.z.push_back(cplx(-1.f, 0.f) / cplx(0.f, std::cos(th)));
            A.z.push_back(cplx( 1.f, 0.f) / cplx(0.f, std::cos(th)));
        }
        return true;
    }
    if (ftype == ELLIP) {
        const float Gp = std::pow(10.0f, -ap / 20.0f), Gs = std::pow(10.0f, -as / 20.0f);
        const float ep = std::sqrt(1.0f / (Gp * Gp) - 1.0f), es = std::sqrt(1.0f / (Gs * Gs) - 1.0f);
        k0 = cplx(r ? 1.0f : 1.0f / std::sqrt(1.0f + ep * ep), 0.f);
        const float k1 = ep / es, N = (float)n, k = ellip::degree(N, k1);
        const cplx j(0.f, 1.f);
        const cplx v0 = -j * ellip::asn(j / ep, k1) / N;
        for (unsigned i = 0; i < L; i++) {
            const float u = (2.0f * (i + 1) - 1.0f) / N;
            const cplx z = j * 1.0f / (k * ellip::cd(cplx(u, 0.f), k));
            const cplx p = j * ellip::cd(cplx(u, 0.f) - j * v0, k);
            A.z.push_back(z); A.z.push_back(std::conj(z));
            A.p.push_back(p); A.p.push_back(std::conj(p));
        }
        if (r) A.p.push_back(j * ellip::sn(j * v0, k));
        return true;
    }
    if (ftype == BESSEL) {
        // poles of the delay-normalised prototype over the approximate 3 dB frequency sqrt((2n-1) ln 2)
        const float w3dB = std::sqrt((2 * n - 1) * std::log(2.0f));
        for (const auto &z : bessel_roots(n)) A.p.emplace_back((float)z.real() / w3dB, (float)z.imag() / w3dB);
        return true;
    }
    return false;
}

inline float prewarp(int btype, float fc, float f0)
{
    float m = 0.f;
    switch (btype) {
    case LOWPASS:  m = std::tan((float)(kPi * fc)); break;
    case HIGHPASS: m = -std::cos((float)(kPi * fc)) / std::sin((float)(kPi * fc)); break;
    case BANDPASS: m = (std::cos((float)(2 * kPi * fc)) - std::cos((float)(2 * kPi * f0))) / std::sin((float)(2 * kPi * fc)); break;
    case BANDSTOP: m = std::sin((float)(2 * kPi * fc)) / (std::cos((float)(2 * kPi * fc)) - std::cos((float)(2 * kPi * f0))); break;
    }
    return std::fabs(m);
}

// bilinear transform s -> z with frequency scale m; the digital gain starts from the nominal k0
inline void bilinear(const Zpk &A, cplx k0, float m, Zpk &D)
{
    const size_t n = std::max(A.z.size(), A.p.size());
    D.z.assign(n, cplx(-1.f, 0.f)); D.p.assign(n, cplx(-1.f, 0.f));
    cplx G = k0;
    const cplx one(1.f, 0.f);
    for (size_t i = 0; i < n; i++) {
        if (i < A.z.size()) { cplx zm = A.z[i] * m; D.z[i] = (one + zm) / (one - zm); }
        if (i < A.p.size()) { cplx pm = A.p[i] * m; D.p[i] = (one + pm) / (one - pm); }
        G *= (one - D.p[i]) / (one - D.z[i]);
    }
    D.k = G;
}

inline void lowpass_to_bandpass(Zpk &D, float f0)
{
    const float c0 = std::cos((float)(2 * kPi * f0));
    auto split = [&](const std::vector<cplx> &in) {
        std::vector<cplx> out;
        for (cplx r : in) {
            cplx t0 = cplx(1.f, 0.f) + r;
            cplx s = std::sqrt(c0 * c0 * t0 * t0 - 4.f * r);
            out.push_back(0.5f * (c0 * t0 + s));
            out.push_back(0.5f * (c0 * t0 - s));
        }
        return out;
    };
    D.z = split(D.z); D.p = split(D.p);
}

// conjugate pairing: pairs first (negative imaginary part leading, ordered by increasing real part),
// then the purely real roots in increasing order
inline std::vector<cplx> pair_conjugates(const std::vector<cplx> &z, float tol)
{
    const size_t n = z.size();
    std::vector<char> used(n, 0);
    std::vector<cplx> pairs, reals;
    for (size_t i = 0; i < n; i++) {
        if (used[i] || std::fabs(z[i].imag()) < tol) continue;
        for (size_t j = 0; j < n; j++) {
            if (j == i || used[j] || std::fabs(z[j].imag()) < tol) continue;
            if (std::fabs(z[i].imag() + z[j].imag()) < tol && std::fabs(z[i].real() - z[j].real()) < tol) {
                cplx lead = z[i].imag() < 0 ? z[i] : std::conj(z[i]);
                pairs.push_back(lead); pairs.push_back(std::conj(lead));
                used[i] = used[j] = 1;
                break;
            }
        }
    }
    for (size_t i = 0; i < n; i++) if (!used[i]) reals.push_back(z[i]);
    const size_t np = pairs.size() / 2;
    for (size_t i = 0; i < np; i++)                       // stable bubble sort keeps liquid's tie order
        for (size_t j = np - 1; j > i; j--)
            if (pairs[2 * (j - 1)].real() > pairs[2 * j].real()) {
                std::swap(pairs[2 * (j - 1)], pairs[2 * j]);
                std::swap(pairs[2 * (j - 1) + 1], pairs[2 * j + 1]);
            }
    for (size_t i = 0; i < reals.size(); i++)
        for (size_t j = reals.size() - 1; j > i; j--)
            if (reals[j - 1].real() > reals[j].real()) std::swap(reals[j - 1], reals[j]);
    pairs.insert(pairs.end(), reals.begin(), reals.end());
    return pairs;
}

// second-order sections: B[3*i..], A[3*i..]; total gain spread evenly over the sections
inline void zpk_to_sos(const Zpk &D, std::vector<float> &B, std::vector<float> &A)
{
    const size_t n = D.p.size(), r = n % 2, L = (n - r) / 2;
    std::vector<cplx> zp = pair_conjugates(D.z, 1e-6f), pp = pair_conjugates(D.p, 1e-6f);
    B.assign(3 * (L + r), 0.f); A.assign(3 * (L + r), 0.f);
    for (size_t i = 0; i < L; i++) {
        cplx p0 = -pp[2 * i], p1 = -pp[2 * i + 1], z0 = -zp[2 * i], z1 = -zp[2 * i + 1];
        A[3 * i] = 1.f; A[3 * i + 1] = (p0 + p1).real(); A[3 * i + 2] = (p0 * p1).real();
        B[3 * i] = 1.f; B[3 * i + 1] = (z0 + z1).real(); B[3 * i + 2] = (z0 * z1).real();
    }
    if (r) {
        A[3 * L] = 1.f; A[3 * L + 1] = (-pp[n - 1]).real();
        B[3 * L] = 1.f; B[3 * L + 1] = (-zp[n - 1]).real();
    }
    const float k = std::pow(D.k.real(), 1.0f / (float)(L + r));
    for (auto &b : B) b *= k;
}

// 0 ok, -1 bad arguments, -2 family not built
inline int iirdes_zpk(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as, Zpk &D)
{
    if (order == 0 || order > 16 || fc <= 0.f || fc >= 0.5f || ap <= 0.f || as <= 0.f) return -1;
    if ((btype == BANDPASS || btype == BANDSTOP) && (f0 < 0.f || f0 > 0.5f)) return -1;
    if (btype < LOWPASS || btype > BANDSTOP) return -1;
    Zpk A; cplx k0;
    if (!analog_prototype(ftype, order, ap, as, A, k0)) return -2;
    bilinear(A, k0, prewarp(btype, fc, f0), D);
    if (btype == HIGHPASS || btype == BANDSTOP) { for (auto &v : D.z) v = -v; for (auto &v : D.p) v = -v; }
    if (btype == BANDPASS || btype == BANDSTOP) lowpass_to_bandpass(D, f0);
    return 0;
}

inline int iirdes_sos(int ftype, int btype, unsigned order, float fc, float f0, float ap, float as,
                      std::vector<float> &B, std::vector<float> &A)
{
    Zpk D;
    int rc = iirdes_zpk(ftype, btype, order, fc, f0, ap, as, D);
    if (rc) return rc;
    zpk_to_sos(D, B, A);
    return 0;
}

inline cplx sos_freqresponse(const std::vector<float> &B, const std::vector<float> &A, float fc)
{
    cplx H(1.f, 0.f);
    const cplx e1 = std::polar(1.0f, (float)(-2 * kPi * fc)), e2 = std::polar(1.0f, (float)(-4 * kPi * fc));
    for (size_t s = 0; s < B.size() / 3; s++)
        H *= (B[3 * s] + B[3 * s + 1] * e1 + B[3 * s + 2] * e2) / (A[3 * s] + A[3 * s + 1] * e1 + A[3 * s + 2] * e2);
    return H;
}

inline cplx fir_freqresponse(const std::vector<float> &h, float scale, float fc)
{
    cplx H(0.f, 0.f);
    for (size_t i = 0; i < h.size(); i++) H += h[i] * std::polar(1.0f, (float)(-2 * kPi * fc * (double)i));
    return H * scale;
}

// ------------------------------------------------------------------ de-emphasis (reference's own formula)
// iirfilter.hpp:366-370: float x = exp(-1.0/(75.0E-6 * sr)); a = {1, -x}; b = {1 - x}
inline void deemph_coeffs(float sample_rate, float &b0, float &a1)
{
    float x = (float)std::exp(-1.0 / (75.0E-6 * (double)sample_rate));
    a1 = -x;
    b0 = (float)(1.0 - (double)x);
}

// ------------------------------------------------------------------ AGC logarithm table
// 128 pairs (inv_i, -ln(inv_i)), inv_i = 1/(1 + (i + 0.5)/128) rounded to double; see devmath.cuh log_rn
inline std::vector<double> log_table()
{
    std::vector<double> t(256);
    for (int i = 0; i < 128; i++) {
        const double inv = 1.0 / (1.0 + (i + 0.5) / 128.0);
        t[2 * i] = inv; t[2 * i + 1] = -std::log(inv);
    }
    return t;
}

// ------------------------------------------------------------------ oscillator
// radians -> 32-bit phase.  float product with 1/(2 pi) held in double, fractional part in float,
// scale by 2^32 in float; a fractional part that rounds up to 1.0f wraps to 0 (what the x86-64
// float -> uint32 conversion of liquid's expression yields).
// Taylor table for atan2_rn (devmath.cuh): node c_i = i/64, i = 0..64; row i holds atan(c_i) and the coefficients
// a_k = cos^k(t) sin(k (t + pi/2)) / k, t = atan(c_i), k = 1..7 -- atan(c_i + d) to 2e-18 for |d| <= 1/128
inline std::vector<double> atan_table()
{
    std::vector<double> t(65 * 8);
    for (int i = 0; i <= 64; i++) {
        const long double th = atanl((long double)i / 64.0L);
        t[8 * i] = (double)th;
        for (int k = 1; k < 8; k++) t[8 * i + k] = (double)(powl(cosl(th), k) * sinl(k * (th + 1.5707963267948966192313216916398L)) / k);
    }
    return t;
}

inline uint32_t nco_constrain(float theta)
{
    float p = (float)(theta * 0.159154943091895);
    float fpart = p - (float)((long)p);
    if (fpart < 0.) fpart = (float)(fpart + 1.);
    float scaled = fpart * 4294967296.0f;
    return (uint32_t)(uint64_t)(int64_t)scaled;
}

inline std::vector<float> nco_sintab()
{
    std::vector<float> t(1024);
    for (unsigned i = 0; i < 1024; i++) t[i] = std::sin((float)(2.0f * kPi * (float)i / 1024.0f));
    return t;
}

inline float nco_u32_to_phase(uint32_t th) { return (float)(2.0f * kPi * (float)th / (float)(0xffffffffu)); }
inline float nco_u32_to_frequency(uint32_t d)
{
    float f = nco_u32_to_phase(d);
    return f > kPi ? (float)(f - 2 * kPi) : f;
}

}  // namespace design
}  // namespace lqb
