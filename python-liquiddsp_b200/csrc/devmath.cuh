// devmath.cuh -- device helpers: packed f32x2 arithmetic (FFMA2/FMUL2/FADD2 on sm_100a),
// cp.async staging, and the integer phase arithmetic of the oscillator.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lqb {

typedef unsigned long long u64;

// A complex sample rides in one 64-bit register pair (re, im).  The filters on this path have real
// coefficients, so both lanes run the same recurrence and one FFMA2 does the work of two FFMAs.
__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ u64 pk(float2 v) { return pk(v.x, v.y); }
__device__ __forceinline__ float2 upk(u64 v)
{
    float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c)
{
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b)
{
    u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b)
{
    u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}

// 16-byte async copy global -> shared; bytes beyond src_bytes are zero-filled.
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, int src_bytes)
{
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) completing on an mbarrier ---------------------------------
__device__ __forceinline__ void mbar_init(void *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tLQB_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra LQB_WAIT;\n\t}" :: "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// bytes: multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(unsigned smem_dst, const void *gsrc, unsigned bytes, void *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// one [32 rows x 128 B] box of a 2-D tensor map into shared memory (SASS UTMALDG), completing on an mbarrier
__device__ __forceinline__ void tma_load_2d(unsigned smem_dst, const void *tmap, int c0, int c1, void *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// L2 prefetch of a tensor-map box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_2d(const void *tmap, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" :: "l"(tmap), "r"(c0), "r"(c1) : "memory");
}

// TMA store of a shared-memory box to a tensor-map box (bulk async group) and its bookkeeping
__device__ __forceinline__ void tma_store_2d(const void *tmap, int c0, int c1, unsigned smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" :: "l"(tmap), "r"(c0), "r"(c1), "r"(smem_src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// radians -> uint32 phase, the oscillator's own arithmetic (see design.hpp nco_constrain): liquid takes the 1/(2 pi)
// product in double and rounds it to float, everything after is float; a fractional part that rounds up to 1.0f wraps
// to phase 0.  The carrier loops call this twice per sample on their critical path, and the double product costs two
// format conversions (19 cycles each on sm_100a) around the DMUL, a FRND and a 64-bit F2I: ~110 cycles of the loop's
// ~170.  Here the same value in single precision: K = 1/(2 pi) as the double literal liquid uses splits exactly into
// K1 + K2 + K3 (24 + 24 + 3 bits); theta K1 as head + exact tail (FMUL, FFMA), plus theta K2, summed small to large.
// tools/check_nco_constrain.c compares the result with (float)((double)theta * K) -- and the whole function with the
// double / int64 formulation -- for EVERY finite float theta: 4,278,190,080 inputs, no mismatch (tests run a sample).
__device__ __forceinline__ uint32_t nco_constrain_dev(float theta)
{
    const float K1 = 0x1.45f306p-3f, K2 = 0x1.b9390ep-28f;
    const float h1 = __fmul_rn(theta, K1), l1 = __fmaf_rn(theta, K1, -h1), h2 = __fmul_rn(theta, K2);
    const float p = __fadd_rn(h1, __fadd_rn(l1, h2));
    float fpart = p;
    if (fabsf(p) >= 1.0f) fpart = __fsub_rn(p, truncf(p));          // (loop corrections are far below one turn)
    if (fpart < 0.f) fpart = __fadd_rn(fpart, 1.0f);
    const float scaled = __fmul_rn(fpart, 4294967296.0f);
    return scaled >= 4294967296.0f ? 0u : __float2uint_rz(scaled);   // (cvt.rzi.u32 saturates; liquid's cast wraps 2^32 to 0)
}

// y = x * exp(+j theta) or x * exp(-j theta) with (s, c) = (sin, cos):
//   up  : re = fma(xr, c, -(xi*s))   im = fma(xi, c,   xr*s )
//   down: re = fma(xr, c,   xi*s )   im = fma(xi, c, -(xr*s))
__device__ __forceinline__ float2 mix_up(float2 x, float2 sc)
{
    return make_float2(__fmaf_rn(x.x, sc.y, -__fmul_rn(x.y, sc.x)), __fmaf_rn(x.y, sc.y, __fmul_rn(x.x, sc.x)));
}
__device__ __forceinline__ float2 mix_down(float2 x, float2 sc)
{
    return make_float2(__fmaf_rn(x.x, sc.y, __fmul_rn(x.y, sc.x)), __fmaf_rn(x.y, sc.y, -__fmul_rn(x.x, sc.x)));
}
// atan2f for the feed-forward frequency discriminator (FreqDem): branch-free, ~2 ulp.  The octant is folded to
// t = min/max in [0, 1]; atan(t) = t + t s P(s), s = t^2, P a degree-7 fit (1.3 ulp on [0, 1]); the quotient is the
// hardware reciprocal times the numerator.  (The PLL demodulators do not use this: their arg() feeds a loop and is
// taken correctly rounded.)
__device__ __forceinline__ float atan2_fast(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    float t = __fdividef(mn, mx);
    t = mx == 0.f ? 0.f : t;
    const float s = __fmul_rn(t, t);
    float r = 0x1.7ed1eap-9f;
    r = __fmaf_rn(r, s, -0x1.0c2bf2p-6f);
    r = __fmaf_rn(r, s, 0x1.61fdcp-5f);
    r = __fmaf_rn(r, s, -0x1.3556acp-4f);
    r = __fmaf_rn(r, s, 0x1.b4e126p-4f);
    r = __fmaf_rn(r, s, -0x1.230adcp-3f);
    r = __fmaf_rn(r, s, 0x1.9978f4p-3f);
    r = __fmaf_rn(r, s, -0x1.5554dcp-2f);
    r = __fmul_rn(r, s);
    r = __fmaf_rn(r, t, t);
    r = ay > ax ? __fsub_rn(1.57079637f, r) : r;
    r = __float_as_int(x) < 0 ? __fsub_rn(3.14159274f, r) : r;      // sign bit, so that atan2(+-0, -0) = +-pi as in libm
    return copysignf(r, y);
}

// arg() for the PLL demodulators: atan2 evaluated in double to ~1 ulp(double) and rounded once to float -- the same
// number as (float)atan2((double)y, (double)x) except when the double result sits within an ulp of a float tie
// (probability ~1e-8).  The quotient of the smaller by the larger magnitude lands in [0, 1]; the nearest of 65
// nodes c_i = i/64 supplies atan(c_i) and a degree-7 Taylor expansion around it (table in shared memory, 4160 B).
// Inline, so independent evaluations overlap; the library routine is a call.
// (the shared-memory copy of the table keeps its 65 nodes kAtanPitch = 9 doubles apart: with 8 -- 64 bytes -- every node
// starts on bank 0 or 16 and the 32 lanes' loads of one coefficient collide 16 ways; 9 spreads the nodes over 16 bank pairs)
constexpr int kAtanPitch = 9;
__device__ __forceinline__ float atan2_rn(float y, float x, const double *__restrict__ tab)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const bool swap = ay > ax;
    const double mx = (double)(swap ? ay : ax), mn = (double)(swap ? ax : ay);
    double q = mn / mx;
    q = mx == 0.0 ? 0.0 : q;
    const int i = __double2int_rn(q * 64.0);
    const double d = fma((double)i, -0.015625, q);
    const double *t = tab + kAtanPitch * i;
    double p = t[7];
    p = fma(p, d, t[6]); p = fma(p, d, t[5]); p = fma(p, d, t[4]); p = fma(p, d, t[3]);
    p = fma(p, d, t[2]); p = fma(p, d, t[1]); p = fma(p, d, t[0]);
    p = swap ? 1.5707963267948966 - p : p;
    p = __float_as_int(x) < 0 ? 3.141592653589793 - p : p;
    return copysignf((float)p, y);
}

// exp(a) rounded once to float.  The AGC multiplies its gain by exp(-alpha/2 * ln(y2')) every sample, a
// factor within a few ulp of 1; a one-ulp bias there (CUDA's expf is allowed two) accumulates to
// bias/alpha = 1e-5 in the gain.  For the small arguments the loop produces a degree-9 Taylor series in
// double is exact far below float resolution; larger arguments take the double-precision library exp.
__device__ __forceinline__ float exp_rn_small(float a)
{
    const double x = (double)a;
    const float aa = fabsf(a);
    if (aa <= 0.0078125f) {
        // the settled loop: |a| <= 2^-7, six terms leave 2^-42/720 -- evaluated as two short chains
        const double x2 = x * x;
        const double lo = fma(x, fma(x, 0.5, 1.0), 1.0);                          // 1 + x + x^2/2
        const double hi = fma(x, fma(x, 1.0 / 120.0, 1.0 / 24.0), 1.0 / 6.0);     // 1/6 + x/24 + x^2/120
        return (float)fma(x2 * x, hi, lo);
    }
    if (aa > 0.125f) return (float)exp(x);
    double p = 1.0 / 362880.0;
    p = fma(p, x, 1.0 / 40320.0); p = fma(p, x, 1.0 / 5040.0); p = fma(p, x, 1.0 / 720.0);
    p = fma(p, x, 1.0 / 120.0);   p = fma(p, x, 1.0 / 24.0);   p = fma(p, x, 1.0 / 6.0);
    p = fma(p, x, 0.5);           p = fma(p, x, 1.0);          p = fma(p, x, 1.0);
    return (float)p;
}

// ln(x) rounded once to float.  x = 2^e * m, m in [1,2); the top 7 mantissa bits pick c_i = 1 + (i+.5)/128
// from a table of (inv_i = rounded 1/c_i, -ln(inv_i)); r = m*inv_i - 1 is one exact-product FMA, |r| <= 2^-8,
// and ln(1+r) is a degree-5 series (next term 2^-48/6).  All in double, so the value rounded to float is the
// correctly rounded logf except within ~1e-15 of a tie -- the same as the library log at a third of the
// instructions and a fifth of the dependent chain.  Valid for normal positive x (the AGC calls it for x > 1e-6).
// exp_rn_small with the settled-loop test taken once per warp: when every lane's argument is within 2^-7 (the steady
// state of the gain loop) the warp runs the short series with no divergence bookkeeping; otherwise the general routine
__device__ __forceinline__ float exp_rn_warp(float a)
{
    if (__all_sync(__activemask(), fabsf(a) <= 0.0078125f)) {
        const double x = (double)a, x2 = x * x;
        const double lo = fma(x, fma(x, 0.5, 1.0), 1.0);
        const double hi = fma(x, fma(x, 1.0 / 120.0, 1.0 / 24.0), 1.0 / 6.0);
        return (float)fma(x2 * x, hi, lo);
    }
    return exp_rn_small(a);
}
__device__ __forceinline__ float log_rn(float x, const double2 *__restrict__ tab)
{
    // exponent, table index and the mantissa as a double in [1, 2) straight from the float's bits (x is a normal
    // float here: the gain loop only takes the logarithm above 1e-6) -- the same numbers as widening first, minus
    // one conversion on the loop's critical path
    const unsigned fb = __float_as_uint(x);
    const int e = (int)((fb >> 23) & 0xffu) - 127;
    const int idx = (int)(fb >> 16) & 127;
    const double m = __hiloint2double((int)(0x3ff00000u | ((fb & 0x007fffffu) >> 3)), (int)((fb & 7u) << 29));
    const double2 t = tab[idx];
    const double r = fma(m, t.x, -1.0);
    // r - r^2/2 + r^3/3 - r^4/4 + r^5/5 as two short chains
    const double r2 = r * r;
    const double lo = fma(r, -0.5, 1.0);                                          // 1 - r/2
    const double hi = fma(r, fma(r, 0.2, -0.25), 1.0 / 3.0);                      // 1/3 - r/4 + r^2/5
    const double p = fma(r2, hi, lo);                                             // ln(1+r) / r
    return (float)fma(p, r, fma((double)e, 0.6931471805599453094, t.y));
}

// ---- the gain loop in single precision (agc_crcf_execute, unlocked, squelch disabled) ----------------------------
// liquid: y = x g;  y2' = (1.0 - alpha) y2' + alpha |y|^2 (double product, one rounding);  g *= expf(-alpha/2 logf(y2'))
// The loop is one dependent chain per sample and the kernel's speed IS that chain's latency, so it is kept to
// FP32 pipe operations and one MUFU:
//   * (1 - alpha) is split into two floats chi + clo = 1 - alpha exactly; chi y2' + clo y2' is formed as a
//     head + tail pair while the previous sample's logarithm is still in flight; the chain adds alpha |y|^2 into the
//     tail and then the head, so y2' is rounded once at full magnitude as liquid's double expression is;
//   * a = -alpha/2 ln(y2') = (-alpha/2 ln 2) lg2.approx(y2'): absolute error 2^-22 ln2 alpha/2 = 8e-10 at alpha = 0.01,
//     against the 3e-8 granularity of the factor exp(a) near 1;
//   * exp(a) = 1 + (a + a^2 q(a)), q of degree 5: the sum is rounded ONCE, which is expf correctly rounded except
//     within ~ulp(a) of a tie, so the loop has liquid's own dead zone around y2' = 1 and no bias (a one-ulp bias in
//     the factor would shift the settled gain by bias / alpha).  |a| > 1/4 (the first samples of an acquisition at
//     large alpha) takes ex2.approx instead.
// Measured against the oracle's correctly rounded double evaluation: relative L2 1e-7 (tests/test_parity_gpu.py).
struct AgcFast {
    float alpha, chi, clo, cl2, c, scale;      // cl2 = -alpha/2 ln 2, c = -alpha/2
};
// BIG = false: the caller guarantees |a| <= 1/2 for every finite y2' (alpha <= 0.0112, liquid's default bandwidth is
// 0.01: |a| <= alpha/2 * 88.8).  The degree-8 series then covers every argument and the MUFU.EX2 of the general case --
// which ptxas merges with the series result through a predicated write, putting its latency ON the chain -- disappears.
template <bool BIG>
__device__ __forceinline__ float2 agc_step_fast(float2 z, float &g, float &y2p, const AgcFast &k)
{
    // off the chain (y2' is a sample old): chi y2' as head + tail, clo y2' folded into the tail
    const float h = __fmul_rn(k.chi, y2p), l = __fmaf_rn(k.clo, y2p, __fmaf_rn(k.chi, y2p, -h));
    const float yr = __fmul_rn(z.x, g), yi = __fmul_rn(z.y, g);
    const float y2 = __fmaf_rn(yr, yr, __fmul_rn(yi, yi));
    y2p = __fadd_rn(h, __fmaf_rn(k.alpha, y2, l));
    float l2; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(y2p));
    const float a = __fmul_rn(k.cl2, l2);
    const float a2 = __fmul_rn(a, a);
    const bool upd = y2p > 1e-6f;
    float e;
    if constexpr (BIG) {
        float eb; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(__fmul_rn(k.c, l2)));
        const float q0 = __fmaf_rn(a, 1.0f / 6.0f, 0.5f), q1 = __fmaf_rn(a, 1.0f / 120.0f, 1.0f / 24.0f), q2 = __fmaf_rn(a, 1.0f / 5040.0f, 1.0f / 720.0f);
        const float q = __fmaf_rn(__fmaf_rn(q2, a2, q1), a2, q0);
        const float ep = __fadd_rn(1.0f, __fmaf_rn(a2, q, a));
        const float ebs = upd ? eb : 1.0f;
        e = (upd && fabsf(a) <= 0.25f) ? ep : ebs;
    } else {
        // exp(a) - 1 = a + a^2 (1/2 + a/6 + a^2 (1/24 + a/120 + a^2 (1/720 + a/5040 + a^2/40320))): |a| <= 1/2 leaves 3e-10
        const float q0 = __fmaf_rn(a, 1.0f / 6.0f, 0.5f), q1 = __fmaf_rn(a, 1.0f / 120.0f, 1.0f / 24.0f);
        const float q2 = __fmaf_rn(a2, 1.0f / 40320.0f, __fmaf_rn(a, 1.0f / 5040.0f, 1.0f / 720.0f));
        const float q = __fmaf_rn(__fmaf_rn(q2, a2, q1), a2, q0);
        const float t = __fmaf_rn(a2, q, a);
        e = __fadd_rn(1.0f, upd ? t : 0.0f);              // y2' <= 1e-6 (or NaN): the gain keeps its value
        // y2' = +inf (overflowed input): liquid's expf(-inf) = 0 zeroes the gain; here t is NaN and fminf(NaN, 0) = 0
        g = fminf(__fmul_rn(g, e), y2p > 3.0e38f ? 0.0f : 1e6f);
        return make_float2(__fmul_rn(yr, k.scale), __fmul_rn(yi, k.scale));
    }
    g = fminf(__fmul_rn(g, e), 1e6f);
    return make_float2(__fmul_rn(yr, k.scale), __fmul_rn(yi, k.scale));
}

// bytes_to_iq (reference utility.hpp:61-69): (float)s / 32767.0f for an int16 pair.  One Newton step on s * fl(1/32767)
// gives the correctly rounded quotient for every one of the 65536 possible inputs (checked exhaustively in
// tests/test_oracle_kat.py), without an IEEE division per sample.
__device__ __forceinline__ float2 i16_to_iq(unsigned packed)
{
    const float r = (float)(1.0 / 32767.0);
    const float a = (float)(short)(packed & 0xffffu), b = (float)(short)(packed >> 16);
    const float qa = __fmul_rn(a, r), qb = __fmul_rn(b, r);
    return make_float2(__fmaf_rn(__fmaf_rn(-qa, 32767.0f, a), r, qa), __fmaf_rn(__fmaf_rn(-qb, 32767.0f, b), r, qb));
}

__device__ __forceinline__ unsigned nco_index(uint32_t theta) { return ((theta + (1u << 21)) >> 22) & 0x3ffu; }

}  // namespace lqb
