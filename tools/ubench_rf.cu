// Microbenchmarks behind the lane-split front kernel (round 2): what the FP32 pipe of sm_100a issues per cycle for
// the operand forms the biquad cascade can be written in, and the cascade itself (compute only, tiles from shared memory)
// as  P: packed FFMA2, one channel per thread (re, im in a register pair)  vs  S: scalar FFMA, one COMPONENT per thread
// (re / im lanes), coefficients read from the constant bank.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_rf ubench_rf.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

struct Coef { float na1[4], na2[4], b0[4], b1[4], b2[4]; };

// ---------------------------------------------------------------- operand forms
// FORM 0: FFMA  d = c[k] * a + b        (two registers + constant bank)
// FORM 1: FFMA  d = r  * a + b          (three distinct registers)
// FORM 2: FFMA2 d = p  * a + b          (three distinct register pairs, rotating so no operand repeats)
// FORM 3: FFMA2 d = q  * a + b          (multiplier pair shared by consecutive instructions)
template <int FORM, int CH>
__global__ void kform(float *out, const __grid_constant__ Coef k, int iters)
{
    const float t = threadIdx.x * 1e-3f;
    float a[CH], b[CH], r[CH]; u64 A[CH], B[CH], R[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = t + i; b[i] = 0.5f - t * i; r[i] = 0.999f - 1e-4f * i * t; A[i] = pk(a[i], b[i]); B[i] = pk(b[i], a[i]); R[i] = pk(r[i], r[i]); }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++)
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (FORM == 0) { const float d = __fmaf_rn(k.na1[i & 3], a[i], b[i]); b[i] = a[i]; a[i] = d; }
                if (FORM == 1) { const float d = __fmaf_rn(r[i], a[i], b[i]); b[i] = a[i]; a[i] = d; }
                if (FORM == 2) { const u64 d = fma2(R[i], A[i], B[i]); B[i] = A[i]; A[i] = d; }
                if (FORM == 3) { const u64 d = fma2(R[0], A[i], B[i]); B[i] = A[i]; A[i] = d; }
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) { s += a[i] + b[i]; const float2 v = upk(A[i]), w = upk(B[i]); s += v.x + v.y + w.x + w.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int FORM, int CH> void runform(const char *name, int warps_per_sm, float *out)
{
    Coef k; for (int i = 0; i < 4; i++) { k.na1[i] = 0.9991f - 1e-4f * i; k.na2[i] = -0.5f; k.b0[i] = 0.1f; k.b1[i] = 0.2f; k.b2[i] = 0.1f; }
    const int iters = 8000, threads = 32 * warps_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kform<FORM, CH><<<148, threads>>>(out, k, 50);
    cudaEventRecord(e0);
    kform<FORM, CH><<<148, threads>>>(out, k, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double inst_per_sched = (double)iters * 4 * CH * warps_per_sm / 4.0;
    printf("%-34s chains %2d warps/SM %2d : %.2f cycles per instruction per scheduler\n", name, CH, warps_per_sm, cycles / inst_per_sched);
}

// ---------------------------------------------------------------- the cascade, compute only
constexpr int TS = 16, NS = 4;
// P: one channel per thread, packed.  Tile rows [32 channels][16 samples] complex in shared memory (dense 128-byte rows,
// chunk-swizzled like the TMA boxes), tap stream of (tap, keep) pairs.  Same instruction stream as seq_kernel's body.
__global__ void __launch_bounds__(256) kcasc_packed(float *out, const __grid_constant__ Coef k, int tiles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    unsigned char *tile = smem + wid * (32 * 128);
    float2 *taps = (float2 *)(smem + (blockDim.x >> 5) * (32 * 128)) + wid * TS;
    for (int i = lane; i < 32 * 16; i += 32) ((float2 *)tile)[i] = make_float2(1e-3f * i, -1e-3f * i);
    if (lane < TS) taps[lane] = make_float2(0.01f * lane, 1.f);
    __syncwarp();
    u64 ca1[NS], ca2[NS], cb0[NS], cb1[NS], cb2[NS], v1[NS], v2[NS], acc = 0;
#pragma unroll
    for (int s = 0; s < NS; s++) { ca1[s] = pk(k.na1[s], k.na1[s]); ca2[s] = pk(k.na2[s], k.na2[s]); cb0[s] = pk(k.b0[s], k.b0[s]); cb1[s] = pk(k.b1[s], k.b1[s]); cb2[s] = pk(k.b2[s], k.b2[s]); v1[s] = 0; v2[s] = 0; }
    const unsigned swz = lane & 7;
    const unsigned char *row = tile + lane * 128;
#pragma unroll 1
    for (int t = 0; t < tiles; t++) {
        u64 xs[TS], yy[NS];
#pragma unroll
        for (int j = 0; j < TS; j += 2) { const float4 v = *(const float4 *)(row + ((((unsigned)j >> 1) ^ swz) << 4)); xs[j] = pk(v.x, v.y); xs[j + 1] = pk(v.z, v.w); }
#pragma unroll
        for (int kk = 0; kk < TS + NS - 1; kk++) {
#pragma unroll
            for (int s = NS - 1; s >= 0; s--) {
                const int j = kk - s;
                if (j >= 0 && j < TS) {
                    const u64 in = s == 0 ? xs[j] : yy[s - 1];
                    const u64 tt = fma2(ca1[s], v1[s], in);
                    const u64 v0 = fma2(ca2[s], v2[s], tt);
                    u64 y = mul2(cb1[s], v1[s]);
                    y = fma2(cb0[s], v0, y);
                    y = fma2(cb2[s], v2[s], y);
                    v2[s] = v1[s]; v1[s] = v0; yy[s] = y;
                    if (s == NS - 1) { const float2 tk = taps[j]; acc = fma2(acc, pk(tk.y, tk.y), mul2(pk(tk.x, tk.x), y)); }
                }
            }
        }
    }
    const float2 r = upk(acc), q = upk(v1[0]);
    out[blockIdx.x * blockDim.x + tid] = r.x + r.y + q.x;
}

// S: one component per thread (lane pair = one channel), scalar FFMA, coefficients from the constant bank, resampler as
// FMUL + FADD (tiles without a restart), taps loaded four at a time.
__global__ void __launch_bounds__(256) kcasc_scalar(float *out, const __grid_constant__ Coef k, int tiles)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    unsigned char *tile = smem + wid * (16 * 128);               // 16 channels per warp
    float *taps = (float *)(smem + (blockDim.x >> 5) * (16 * 128)) + wid * TS;
    for (int i = lane; i < 16 * 16; i += 32) ((float2 *)tile)[i] = make_float2(1e-3f * i, -1e-3f * i);
    if (lane < TS) taps[lane] = 0.01f * lane;
    __syncwarp();
    float v1[NS], v2[NS], acc = 0.f;
#pragma unroll
    for (int s = 0; s < NS; s++) { v1[s] = 0.f; v2[s] = 0.f; }
    const int ch = lane >> 1, comp = lane & 1;
    const unsigned swz = ch & 7;
    const unsigned char *row = tile + ch * 128 + comp * 4;
#pragma unroll 1
    for (int t = 0; t < tiles; t++) {
        float xs[TS], yy[NS], tp[TS];
#pragma unroll
        for (int j = 0; j < TS; j++) xs[j] = *(const float *)(row + ((((unsigned)j >> 1) ^ swz) << 4) + (j & 1) * 8);
#pragma unroll
        for (int j = 0; j < TS; j += 4) { const float4 v = *(const float4 *)(taps + j); tp[j] = v.x; tp[j + 1] = v.y; tp[j + 2] = v.z; tp[j + 3] = v.w; }
#pragma unroll
        for (int kk = 0; kk < TS + NS - 1; kk++) {
#pragma unroll
            for (int s = NS - 1; s >= 0; s--) {
                const int j = kk - s;
                if (j >= 0 && j < TS) {
                    const float in = s == 0 ? xs[j] : yy[s - 1];
                    const float tt = __fmaf_rn(k.na1[s], v1[s], in);
                    const float v0 = __fmaf_rn(k.na2[s], v2[s], tt);
                    float y = __fmul_rn(k.b1[s], v1[s]);
                    y = __fmaf_rn(k.b0[s], v0, y);
                    y = __fmaf_rn(k.b2[s], v2[s], y);
                    v2[s] = v1[s]; v1[s] = v0; yy[s] = y;
                    if (s == NS - 1) acc = __fadd_rn(acc, __fmul_rn(tp[j], y));
                }
            }
        }
    }
    out[blockIdx.x * blockDim.x + tid] = acc + v1[0];
}

// systolic: lane (channel, section, component): 8 lanes per channel, one section per lane, hand-off by shuffle with a
// lag of D steps between sections.  Reports cycles per sample step of one warp (the single-channel latency).
template <int D>
__global__ void ksyst(float *out, const __grid_constant__ Coef k, int steps, long long *cyc)
{
    const int lane = threadIdx.x & 31, sec = (lane >> 1) & 3;
    const float na1 = k.na1[sec], na2 = k.na2[sec], b0 = k.b0[sec], b1 = k.b1[sec], b2 = k.b2[sec];
    float v1 = 0.f, v2 = 0.f, acc = 0.f, hand[D];
#pragma unroll
    for (int d = 0; d < D; d++) hand[d] = 0.f;
    float x = 1e-3f * lane;
    const long long t0 = clock64();
#pragma unroll 1
    for (int n = 0; n < steps; n += 16) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            // the value produced D steps ago by the previous section's lane
            const float up = __shfl_up_sync(0xffffffffu, hand[(j + 0) % D], 2);
            const float in = sec == 0 ? x : up;
            const float tt = __fmaf_rn(na1, v1, in);
            const float v0 = __fmaf_rn(na2, v2, tt);
            float y = __fmul_rn(b1, v1);
            y = __fmaf_rn(b0, v0, y);
            y = __fmaf_rn(b2, v2, y);
            v2 = v1; v1 = v0; hand[(j + 0) % D] = y;
            acc = __fadd_rn(acc, __fmul_rn(0.01f, y));
            x = __fadd_rn(x, 1e-4f);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v1;
}

template <class F> void runcasc(const char *name, F fn, int chan_per_warp, int warps_per_cta, int ctas_per_sm, float *out)
{
    Coef k; for (int i = 0; i < 4; i++) { k.na1[i] = 1.9f - 0.01f * i; k.na2[i] = -0.95f; k.b0[i] = 0.1f; k.b1[i] = 0.19f; k.b2[i] = 0.1f; }
    const int tiles = 4096, threads = 32 * warps_per_cta;
    const size_t smem = (size_t)warps_per_cta * (chan_per_warp * 128 + TS * 8);
    cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn<<<148 * ctas_per_sm, threads, smem>>>(out, k, 16);
    cudaEventRecord(e0);
    fn<<<148 * ctas_per_sm, threads, smem>>>(out, k, tiles);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double wps = warps_per_cta * ctas_per_sm / 4.0;                       // warps per scheduler
    const double chan_samples_per_sm = (double)tiles * TS * chan_per_warp * warps_per_cta * ctas_per_sm;
    // 22 complex FMA-class operations per channel-sample = 44 lane operations; an SM retires 128 per cycle
    printf("%-14s %4.1f warps/scheduler: %.1f cycles per sample per warp, FMA lanes %.1f %% busy, %.0f GS/s chip-wide at this clock (err %s)\n",
           name, wps, cycles / ((double)tiles * TS), 100.0 * chan_samples_per_sm * 44.0 / (cycles * 128.0),
           chan_samples_per_sm * 148.0 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    float *out; cudaMalloc(&out, 148 * 16 * 1024 * sizeof(float));
    long long *cyc; cudaMalloc(&cyc, 8);
    for (int w : {8, 16, 32}) {
        runform<0, 8>("FFMA  reg, const, reg", w, out);
        runform<1, 8>("FFMA  three registers", w, out);
        runform<2, 8>("FFMA2 three pairs (distinct)", w, out);
        runform<3, 8>("FFMA2 shared multiplier pair", w, out);
    }
    runform<0, 4>("FFMA  reg, const, reg", 4, out); runform<2, 4>("FFMA2 three pairs (distinct)", 4, out);
    for (int c : {1, 2, 4, 7, 8, 14}) runcasc("packed x1w", kcasc_packed, 32, 1, c, out);      // c CTAs of one warp per SM
    for (int c : {4, 8, 14, 16, 28}) runcasc("scalar x1w", kcasc_scalar, 16, 1, c, out);
    for (int c : {1, 2, 4, 7}) runcasc("scalar x4w", kcasc_scalar, 16, 4, c, out);
    {
        Coef k; for (int i = 0; i < 4; i++) { k.na1[i] = 1.9f - 0.01f * i; k.na2[i] = -0.95f; k.b0[i] = 0.1f; k.b1[i] = 0.19f; k.b2[i] = 0.1f; }
        long long h;
        ksyst<1><<<1, 32>>>(out, k, 16000, cyc); ksyst<1><<<1, 32>>>(out, k, 16000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("systolic, lag 1: %.1f cycles per sample (one warp)\n", h / 16000.0);
        ksyst<2><<<1, 32>>>(out, k, 16000, cyc); ksyst<2><<<1, 32>>>(out, k, 16000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("systolic, lag 2: %.1f cycles per sample (one warp)\n", h / 16000.0);
        ksyst<4><<<1, 32>>>(out, k, 16000, cyc); ksyst<4><<<1, 32>>>(out, k, 16000, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("systolic, lag 4: %.1f cycles per sample (one warp)\n", h / 16000.0);
    }
    return 0;
}
