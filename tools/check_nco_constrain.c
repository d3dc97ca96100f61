// Exhaustive proof behind devmath.cuh nco_constrain_dev: liquid's nco_constrain (double product rounded to float, truncf,
// int64 conversion) against the single-precision evaluation the kernels run, for EVERY finite float theta.
//   gcc -O2 -mfma -ffp-contract=off -fopenmp -o check_nco_constrain tools/check_nco_constrain.c -lm && ./check_nco_constrain
// (45 s on 8 cores; prints "tested 4278190080 mismatches 0").  Optional argument: stride (tests every stride-th bit pattern).
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <stdlib.h>
#include <omp.h>
static const double K = 0.159154943091895;
static inline uint32_t ref_constrain(float theta)
{
    float p = (float)((double)theta * K);
    float fpart = p - truncf(p);
    if (fpart < 0.f) fpart = fpart + 1.0f;
    float scaled = fpart * 4294967296.0f;
    return (uint32_t)(unsigned long long)(long long)scaled;
}
static inline uint32_t fast_constrain(float th, float K1, float K2)
{
    const float h1 = th * K1, l1 = fmaf(th, K1, -h1), h2 = th * K2;
    const float p = h1 + (l1 + h2);
    float fpart = fabsf(p) < 1.0f ? p : p - truncf(p);
    if (fpart < 0.f) fpart = fpart + 1.0f;
    const float scaled = fpart * 4294967296.0f;
    // float -> uint32 by truncation, 2^32 wraps to 0 (cvt.rzi.u32.f32 saturates, so the wrap is a select)
    return scaled >= 4294967296.0f ? 0u : (uint32_t)scaled;
}
int main(int argc, char **argv)
{
    const long long stride = argc > 1 ? atoll(argv[1]) : 1;
    const float K1 = (float)K; const float K2 = (float)(K - (double)K1);
    long long bad = 0, tested = 0, badbig = 0;
    #pragma omp parallel for reduction(+:bad,tested,badbig) schedule(static)
    for (long long b = 0; b < (1LL << 32); b += stride) {
        uint32_t u = (uint32_t)b; float th; memcpy(&th, &u, 4);
        if (!isfinite(th)) continue;      // (|p| < 2^63 for the reference's int64 cast to be defined)
        const uint32_t r = ref_constrain(th), f = fast_constrain(th, K1, K2);
        tested++;
        if (r != f) { bad++; if (fabsf(th) > 1e-30f) badbig++; if (bad < 6) printf("theta %a ref %u fast %u\n", th, r, f); }
    }
    printf("tested %lld mismatches %lld (|theta| > 1e-30: %lld)\n", tested, bad, badbig);
    return bad ? 1 : 0;
}
