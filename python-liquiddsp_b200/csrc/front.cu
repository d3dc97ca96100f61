// front.cu -- the full-rate front of the receiver (biquad cascade -> decimating resampler) with TWO channels per
// thread.  Same arithmetic, operand for operand, as seq_kernel<F_IIR | F_RS> (seq.cu); what changes is how much
// independent work one warp carries.
//
// The cascade's dependent chain is four packed operations deep per section and sample; skewing the sections gives a
// thread NS independent chains, and with 65536 channels resident in one wave there are only 3.5 warps per scheduler
// to fill the gaps (ncu: fixed-latency dependency waits lead the stall samples, FMA pipe 61 % busy).  A second channel
// in the same thread doubles the independent chains per warp, and everything that is the same for every channel --
// tap stream, output positions, tile bookkeeping, TMA issue, loop control -- is paid once per two channels.
//
// Staging: one elected lane per warp issues one cp.async.bulk.tensor.2d per tile for the warp's [64 rows x 128 B] box
// (128-byte swizzle, per-warp mbarrier, 3-stage ring); lane l works on rows l and l + 32 of the box.
#include <cuda_runtime.h>
#include <cuda.h>
#include <type_traits>
#include "params.h"
#include "devmath.cuh"
#include "front.h"

namespace lqb {
namespace {

constexpr int BT = 32, TS = 16, CPT = 2;     // one warp per CTA: 1024 CTAs for 65536 channels, 6.9 per SM -- an even single wave
constexpr int ROWS_W = 32 * CPT, ROWS_CTA = (BT / 32) * ROWS_W;       // 64 rows per warp, 128 per CTA
constexpr int ROWB = TS * 8;                                          // bytes per staged row: dense, swizzled

// NST: depth of the staging ring.  3 when the kernel has the SM to itself; 2 (20 KB per CTA instead of 28) leaves room for
// one CTA of the decimated-rate tail kernel on every SM when a chain overlaps call k's tail with call k+1's front
template <int NS, int NST>
__global__ void __launch_bounds__(BT, 7) front2_kernel(const __grid_constant__ SeqArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    smem += (1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u;     // swizzled boxes: 1024-byte aligned
    unsigned char *s_in = smem;                                       // NST stages of [ROWS_CTA][ROWB]
    float2 *s_tap = (float2 *)(s_in + NST * ROWS_CTA * ROWB);         // [warp][NST][TS] (tap, keep)
    int *s_emit = (int *)(s_tap + (BT / 32) * NST * TS);              // [warp][NST] sample an output falls on, or -1
    float *s_bank = (float *)(s_emit + 8);                            // [npfb][sublen]
    __shared__ unsigned long long s_bar[(BT / 32) * NST];

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * ROWS_CTA + wid * ROWS_W;
    const long long CT = a.Ctot, N = a.n;
    long long chl[CPT], gch[CPT]; bool act[CPT];
#pragma unroll
    for (int c = 0; c < CPT; c++) { chl[c] = row0 + c * 32 + lane; act[c] = chl[c] < a.C; gch[c] = a.ch0 + (act[c] ? chl[c] : 0); }

    // ---- coefficients (shared by the two channels) and per-channel state ----
    u64 ca1[NS], ca2[NS], cb0[NS], cb1[NS], cb2[NS], iv1[CPT][NS], iv2[CPT][NS], rs_acc[CPT];
#pragma unroll
    for (int s = 0; s < NS; s++) {
        ca1[s] = pk(-a.iir.a[s][1], -a.iir.a[s][1]); ca2[s] = pk(-a.iir.a[s][2], -a.iir.a[s][2]);
        cb0[s] = pk(a.iir.b[s][0], a.iir.b[s][0]);   cb1[s] = pk(a.iir.b[s][1], a.iir.b[s][1]);
        cb2[s] = pk(a.iir.b[s][2], a.iir.b[s][2]);
#pragma unroll
        for (int c = 0; c < CPT; c++) { iv1[c][s] = pk(a.iir.v[(2 * s + 0) * CT + gch[c]]); iv2[c][s] = pk(a.iir.v[(2 * s + 1) * CT + gch[c]]); }
    }
    const int L = a.rs.sublen;
    for (int i = tid; i < a.rs.npfb * L; i += BT) s_bank[i] = a.rs.bank[i];
    if (tid == 0) { for (int i = 0; i < (BT / 32) * NST; i++) mbar_init(&s_bar[i], 1); mbar_init_fence(); }
    __syncthreads();

    // the first output's window may start before this call; that part comes from the ring (dotprod_cccf arithmetic:
    // each product is rounded, then added -- oldest sample first)
    {
        const long long nnext = a.rs.phase >> 24;
        const unsigned f = (a.rs.phase & 0xffffffu) >> (24 - a.rs.bits);
#pragma unroll
        for (int c = 0; c < CPT; c++) {
            float ar = 0.f, ai = 0.f;
            for (long long j = nnext - (L - 1); j < 0; j++) {
                const int slot = (int)(((long long)a.rs.count + j + 4LL * L) % L);
                const float h = s_bank[f * L + (int)(j - nnext + L - 1)];
                const float2 w = a.rs.ring[slot * CT + gch[c]];
                ar = __fadd_rn(ar, __fmul_rn(h, w.x)); ai = __fadd_rn(ai, __fmul_rn(h, w.y));
            }
            rs_acc[c] = pk(ar, ai);
        }
    }

    // ---- the tap stream: lanes 0..TS-1 of each warp follow one sample position of every tile (see seq.cu) ----
    uint32_t gP = a.rs.phase;
    if (lane < TS) for (int i = 0; i < lane; i++) { if (gP <= 0x00ffffffu) gP += a.rs.step; gP -= (1u << 24); }
    auto gen_taps = [&](int stage, bool first_tile) {
        const bool lane_on = lane < TS;
        const bool emit = lane_on && gP <= 0x00ffffffu;
        if (lane_on) {
            const unsigned cnt = gP >> 24;
            const unsigned f = (gP & 0xffffffu) >> (24 - a.rs.bits);
            const float h = cnt < (unsigned)L ? s_bank[f * L + (L - 1 - (int)cnt)] : 0.f;
            const float keep = (gP < a.rs.step - (1u << 24) || (first_tile && lane == 0)) ? 1.f : 0.f;
            s_tap[(wid * NST + stage) * TS + lane] = make_float2(h, keep);
            if (gP < ((unsigned)TS << 24)) gP += a.rs.step;
            gP -= ((unsigned)TS << 24);
        }
        const unsigned m = __ballot_sync(0xffffffffu, emit);
        if (lane == 0) s_emit[wid * NST + stage] = m ? (__ffs(m) - 1) : -1;
    };
    const unsigned s_in_sh = (unsigned)__cvta_generic_to_shared(s_in);
    auto load_tile = [&](long long t, int stage) {
        if (lane == 0) {
            unsigned long long *bar = &s_bar[wid * NST + stage];
            mbar_arrive_expect_tx(bar, ROWS_W * ROWB);
            tma_load_2d(s_in_sh + stage * (ROWS_CTA * ROWB) + wid * (ROWS_W * ROWB), &a.tmap, (int)(t * (TS * 2)), (int)row0, bar);
        }
    };

    long long kout = 0;
    auto emit_out = [&](int c, u64 v) {
        if (act[c]) {
            const long long o = a.out_tmajor ? kout * a.out_pitch + chl[c] : chl[c] * a.out_pitch + kout;
            ((float2 *)a.y)[o] = upk(v);
        }
    };
    const unsigned swz = (unsigned)(lane & 7);                       // rows l and l + 32 share their chunk permutation
    auto ld_row = [&](const unsigned char *rw, u64 (&xs)[TS]) {
#pragma unroll
        for (int j = 0; j < TS; j += 2) {
            const float4 v = *(const float4 *)(rw + ((((unsigned)j >> 1) ^ swz) << 4));
            xs[j] = pk(v.x, v.y); xs[j + 1] = pk(v.z, v.w);
        }
    };
    auto ld1 = [&](const unsigned char *rw, int j) -> u64 {
        return pk(*(const float2 *)(rw + ((((unsigned)j >> 1) ^ swz) << 4) + (j & 1) * 8));
    };

    // ---- stream the tiles ----
    const long long ntiles = (N + TS - 1) / TS;
    const long long nfast = (N - L) > 0 ? (N - L) / TS : 0;         // complete tiles that need no ring save
    for (int p = 0; p < NST - 1; p++) if (p < ntiles) { load_tile(p, p); gen_taps(p, p == 0); }
    int stage = 0; unsigned phases = 0;
#pragma unroll 1
    for (long long t = 0; t < ntiles; t++) {
        mbar_wait(&s_bar[wid * NST + stage], (phases >> stage) & 1u);
        phases ^= 1u << stage;
        __syncwarp();                          // tile t has landed; every lane is done with tile t-1
        {
            const int sn = stage == 0 ? NST - 1 : stage - 1;
            if (t + NST - 1 < ntiles) { load_tile(t + NST - 1, sn); gen_taps(sn, false); }
        }
        const unsigned char *rowp[CPT];
#pragma unroll
        for (int c = 0; c < CPT; c++) rowp[c] = s_in + stage * (ROWS_CTA * ROWB) + (wid * ROWS_W + c * 32 + lane) * ROWB;
        const float2 *tk = s_tap + (wid * NST + stage) * TS;
        const int e = s_emit[wid * NST + stage];
        if (t < nfast) {
            // Skewed cascade over both channels: at step k section s works on sample k-s of channel 0 and of channel 1 --
            // 2 NS independent chains.  Tiles without an output (62 %) run a copy without the output capture.
            auto body = [&](auto with_emit) {
                u64 xs[CPT][TS];
#pragma unroll
                for (int c = 0; c < CPT; c++) ld_row(rowp[c], xs[c]);
                u64 yy[CPT][NS], outv[CPT];
#pragma unroll
                for (int c = 0; c < CPT; c++) outv[c] = 0ull;
#pragma unroll
                for (int k = 0; k < TS + NS - 1; k++) {
                    // the 2 NS section updates of a step are independent; they are written operation by operation across
                    // all of them (not update by update), so dependent operations sit 2 NS instructions apart
                    u64 tt[NS][CPT], v0[NS][CPT], y[NS][CPT];
#pragma unroll
                    for (int sct = NS - 1; sct >= 0; sct--) {
                        const int j = k - sct;
                        if (j >= 0 && j < TS) {
#pragma unroll
                            for (int c = 0; c < CPT; c++) tt[sct][c] = fma2(ca1[sct], iv1[c][sct], sct == 0 ? xs[c][j] : yy[c][sct - 1]);
                        }
                    }
#pragma unroll
                    for (int sct = NS - 1; sct >= 0; sct--) {
                        const int j = k - sct;
                        if (j >= 0 && j < TS) {
#pragma unroll
                            for (int c = 0; c < CPT; c++) { v0[sct][c] = fma2(ca2[sct], iv2[c][sct], tt[sct][c]); y[sct][c] = mul2(cb1[sct], iv1[c][sct]); }
                        }
                    }
#pragma unroll
                    for (int sct = NS - 1; sct >= 0; sct--) {
                        const int j = k - sct;
                        if (j >= 0 && j < TS) {
#pragma unroll
                            for (int c = 0; c < CPT; c++) y[sct][c] = fma2(cb0[sct], v0[sct][c], y[sct][c]);
                        }
                    }
#pragma unroll
                    for (int sct = NS - 1; sct >= 0; sct--) {
                        const int j = k - sct;
                        if (j >= 0 && j < TS) {
#pragma unroll
                            for (int c = 0; c < CPT; c++) {
                                y[sct][c] = fma2(cb2[sct], iv2[c][sct], y[sct][c]);
                                iv2[c][sct] = iv1[c][sct]; iv1[c][sct] = v0[sct][c]; yy[c][sct] = y[sct][c];
                            }
                        }
                    }
                    {
                        const int j = k - (NS - 1);
                        if (j >= 0 && j < TS) {
                            // acc = acc*keep + round(tap*y): the two roundings of liquid's complex-tap dot product
                            const float2 tkj = tk[j];
                            const u64 tap2 = pk(tkj.x, tkj.x), keep2 = pk(tkj.y, tkj.y);
#pragma unroll
                            for (int c = 0; c < CPT; c++) {
                                rs_acc[c] = fma2(rs_acc[c], keep2, mul2(tap2, yy[c][NS - 1]));
                                if constexpr (decltype(with_emit)::value) { if (j == e) outv[c] = rs_acc[c]; }
                            }
                        }
                    }
                }
                if constexpr (decltype(with_emit)::value) {
#pragma unroll
                    for (int c = 0; c < CPT; c++) emit_out(c, outv[c]);
                    kout++;
                }
            };
            if (e < 0) body(std::false_type{}); else body(std::true_type{});
        } else {
            // the call's last tiles: sample by sample, saving the newest L filtered samples to the history ring
            const long long n0 = t * TS;
            const int nv = (int)((N - n0) < TS ? (N - n0) : TS);
#pragma unroll 1
            for (int j = 0; j < nv; j++) {
                const float2 tkj = tk[j];
#pragma unroll
                for (int c = 0; c < CPT; c++) {
                    u64 x = ld1(rowp[c], j);
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        const u64 tt = fma2(ca1[s], iv1[c][s], x);
                        const u64 v0 = fma2(ca2[s], iv2[c][s], tt);
                        u64 y = mul2(cb1[s], iv1[c][s]);
                        y = fma2(cb0[s], v0, y);
                        y = fma2(cb2[s], iv2[c][s], y);
                        iv2[c][s] = iv1[c][s]; iv1[c][s] = v0; x = y;
                    }
                    rs_acc[c] = fma2(rs_acc[c], pk(tkj.y, tkj.y), mul2(pk(tkj.x, tkj.x), x));
                    if (n0 + j >= N - L && act[c]) a.rs.ring[(int)((a.rs.count + n0 + j) % L) * CT + gch[c]] = upk(x);
                    if (j == e) emit_out(c, rs_acc[c]);
                }
                if (j == e) kout++;
            }
        }
        stage = stage + 1 == NST ? 0 : stage + 1;
    }

    // ---- carried state back to HBM ----
#pragma unroll
    for (int c = 0; c < CPT; c++) {
        if (!act[c]) continue;
#pragma unroll
        for (int s = 0; s < NS; s++) { a.iir.v[(2 * s + 0) * CT + gch[c]] = upk(iv1[c][s]); a.iir.v[(2 * s + 1) * CT + gch[c]] = upk(iv2[c][s]); }
    }
}

typedef void (*FrontFn)(const SeqArgs);
FrontFn pick(int nsos, int nst)
{
    if (nst == 2) {
        switch (nsos) {
        case 1: return front2_kernel<1, 2>; case 2: return front2_kernel<2, 2>; case 3: return front2_kernel<3, 2>; case 4: return front2_kernel<4, 2>;
        default: return nullptr;
        }
    }
    switch (nsos) {
    case 1: return front2_kernel<1, 3>; case 2: return front2_kernel<2, 3>; case 3: return front2_kernel<3, 3>; case 4: return front2_kernel<4, 3>;
    default: return nullptr;
    }
}

size_t smem_bytes(const SeqArgs &a, int NST)
{
    return 1024 + (size_t)NST * ROWS_CTA * ROWB + (size_t)(BT / 32) * NST * TS * sizeof(float2) + 32
         + (((size_t)a.rs.npfb * a.rs.sublen * sizeof(float) + 15) & ~(size_t)15);
}

}  // namespace

bool front2_supported(unsigned mask, int nsos) { return mask == (F_IIR | F_RS) && pick(nsos, 3) != nullptr; }

cudaError_t front2_launch(int nsos, const SeqArgs &a, cudaStream_t stream, int ring_depth)
{
    const int nst = ring_depth == 2 ? 2 : 3;
    FrontFn fn = pick(nsos, nst);
    if (!fn) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const size_t smem = smem_bytes(a, nst);
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const unsigned grid = (unsigned)((a.C + ROWS_CTA - 1) / ROWS_CTA);
    fn<<<grid, BT, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
