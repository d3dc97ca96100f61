// synth.cu -- synthetic IQ generators for the measurement configs.  Not on the
// reference path: the reference takes its IQ from an SDR callback (README.md:60-63).  The signals
// are built on the device because config 5 holds 34 GB of input per block; parity tests copy a
// channel subset back and run exactly those bits through the CPU oracle.
#include <cuda_runtime.h>
#include <math.h>
#include "synth.h"

namespace lqb {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// two independent N(0,1) draws keyed by (seed, channel, sample)
__device__ __forceinline__ float2 gauss2(unsigned long long seed, unsigned long long ch, unsigned long long n)
{
    const unsigned long long h = mix64(mix64(seed ^ (ch * 0xd1342543de82ef95ull)) + n);
    const float u1 = ((float)(unsigned)(h >> 40) + 1.0f) * (1.0f / 16777217.0f);   // (0, 1)
    const float u2 = (float)(unsigned)((h >> 8) & 0xffffffu) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c; sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// exp(j 2 pi f n) with f in cycles/sample, exact phase reduction in double
__device__ __forceinline__ float2 tone(double f, unsigned long long n, double phase_cycles)
{
    double ph = f * (double)n + phase_cycles;
    ph -= floor(ph);
    double s, c; sincospi(2.0 * ph, &s, &c);
    return make_float2((float)c, (float)s);
}

__global__ void synth_kernel(int kind, float2 *x, int C, int ch0, long long n, unsigned long long n0, unsigned long long seed)
{
    const long long total = (long long)C * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long cl = i / n, k = i % n;
        const unsigned long long c = (unsigned long long)(ch0 + cl), t = n0 + (unsigned long long)k;
        const float2 w = gauss2(seed + c, c, t);
        float2 v;
        const double Fs = 2.0e6;
        if (kind == 0) {            // AM broadcast IQ: carrier offset, two audio tones, out-of-band interferer, noise
            double s1, s2, d;
            sincospi(2.0 * fmod(1000.0 / Fs * (double)t, 1.0), &s1, &d);
            sincospi(2.0 * fmod(2500.0 / Fs * (double)t, 1.0), &s2, &d);
            const float am = 0.1f * (1.0f + 0.5f * (0.6f * (float)s1 + 0.4f * (float)s2));
            const double foff = (200.0 + 10.0 * (double)(c % 32)) / Fs;
            const double phc = fmod((double)c * 0.618, 1.0);
            const float2 car = tone(foff, t, phc), itf = tone(60.0e3 / Fs, t, 0.0);
            const float sg = 0.02f * 0.70710678f;
            v = make_float2(am * car.x + 0.05f * itf.x + sg * w.x, am * car.y + 0.05f * itf.y + sg * w.y);
        } else if (kind == 1) {     // white complex Gaussian, sigma 1
            v = w;
        } else if (kind == 2) {     // tone just off the channel's mixer frequency, plus noise
            const double f = 0.05 + 0.4 * (double)(c % 4096) / 4096.0 + 0.002;
            const float2 tn = tone(f, t, 0.0);
            v = make_float2(tn.x + 0.05f * w.x, tn.y + 0.05f * w.y);
        } else {                    // FM IQ: exp(j 2 pi kf sum m), m = 1 kHz tone, amplitude ramp over channels
            const double om = 1000.0 / Fs;                      // cycles/sample
            double sa, sb, sc_, d;
            sincospi(fmod(om * (double)t, 2.0), &sa, &d);       // sin(w t / 2), w = 2 pi om
            sincospi(fmod(om * (double)(t + 1), 2.0), &sb, &d);
            sincospi(om, &sc_, &d);
            const double S = sa * sb / sc_;                     // sum_{i<=t} sin(w i)
            const float amp = 0.01f * powf(100.0f, (float)(c % 1024) / 1023.0f);
            const float2 tn = tone(0.1, 0, S - floor(S));
            double ph = 0.1 * S; ph -= floor(ph);
            double s, cc; sincospi(2.0 * ph, &s, &cc);
            (void)tn;
            v = make_float2(amp * (float)cc + 0.01f * w.x, amp * (float)s + 0.01f * w.y);
        }
        x[i] = v;
    }
}

}  // namespace

cudaError_t synth_launch(int kind, float2 *x, int C, int ch0, long long n, unsigned long long n0,
                         unsigned long long seed, cudaStream_t stream)
{
    if (C <= 0 || n <= 0) return cudaSuccess;
    synth_kernel<<<148 * 16, 256, 0, stream>>>(kind, x, C, ch0, n, n0, seed);
    return cudaGetLastError();
}

}  // namespace lqb
