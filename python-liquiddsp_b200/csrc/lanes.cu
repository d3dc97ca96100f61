// lanes.cu -- the full-rate front of the receiver (biquad cascade -> decimating resampler) with the REAL and the
// IMAGINARY part of a channel on two adjacent lanes.
//
// Replaces the per-sample loops of ComplexIIRFilter::execute (iirfilter.hpp:292-298, iirfilt_crcf_execute_block) and
// ComplexResampler::execute (resampler.hpp:160-172, resamp_cccf_execute) for the README chain; same operations on the
// same operands in the same order as seq_kernel<F_IIR | F_RS> / front2_kernel, so the results are bit-identical.
//
// Why lanes instead of register pairs.  The filters have real coefficients, so the two components of a complex sample
// never meet: a channel is two independent real recurrences.  The packed form (one FFMA2 per complex operation,
// front.cu) reads three 64-bit operands per instruction, and the register file of sm_100a delivers them in 2.7 cycles
// (tools/ubench_rf.cu: 2.66 cycles per FFMA2 with three distinct register pairs, 2.0 only when operands repeat) --
// 75 % of the FP32 pipe at best, and 65536 channels make only 1024 such warps (432 schedulers hold two, 160 hold one).
// One component per lane turns every operation into a scalar FFMA with the coefficient taken from the constant bank
// (two register operands: 1.08 cycles per instruction), makes 4096 warps -- seven on every scheduler, an even load --
// and gives each scheduler seven instruction streams to cover tile boundaries with.
//
// Staging: one elected lane per warp issues one cp.async.bulk.tensor.2d per tile for the warp's [16 rows x 128 B] box
// (128-byte swizzle, per-warp mbarrier, 3-stage ring); lane 2c + p reads component p of row c.  Warps never wait for
// each other: no CTA barrier anywhere.
#include <cuda_runtime.h>
#include <cuda.h>
#include <type_traits>
#include "params.h"
#include "devmath.cuh"
#include "lanes.h"

namespace lqb {
namespace {

constexpr int TS = 16;
constexpr int ROWB = TS * 8;                       // bytes per staged row (dense, swizzled)
constexpr int WARP_TILE = 16 * ROWB;               // a warp's stage: 16 rows
// What is the same for every channel -- which polyphase tap multiplies a sample, whether the dot product restarts on it,
// where in a tile an output falls -- is worked out once per call by tapstream_kernel (closed form of liquid's uint32 phase
// recurrence) and reaches each warp as one 144-byte record per tile, copied by the same mbarrier transaction as the tile.
struct TileRec {
    float tap[TS];                                 // tap that multiplies sample j of the tile (0 outside every window)
    float keep[TS];                                // 0 on the sample after an output (the accumulator restarts), else 1
    int emit;                                      // sample of the tile an output falls on, or -1 (step >= TS * 2^24: at most one)
    int gen;                                       // 1 when the tile holds an output or a restart, 0: plain accumulation
    int pad[2];
};
static_assert(sizeof(TileRec) == 144, "tile records are copied 16 bytes at a time");
// shared memory: every warp's NST tiles first (each a 2 KB swizzled TMA box, 1024-byte aligned), then per warp the
// tile records and the mbarriers
// NST: depth of the staging ring.  3 when every scheduler holds seven warps (the machine is full and shared memory is what
// limits residency); 8 when few channels leave most of an SM empty -- then the bytes a warp keeps in flight are what
// covers the HBM latency (512 warps x 2 tiles x 2 KB = 2 MB in flight sustains only ~2 TB/s)
template <int NST> struct WarpAux {
    TileRec rec[NST];
    unsigned long long bar[NST];
    unsigned long long pad[NST & 1 ? 1 : 2];
};

__device__ __forceinline__ bool elect_one()
{
    unsigned p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}

// phase of the resampler at sample n of the call: P(n) = p0 + k(n) step - n 2^24 with k(n) = #{j >= 0 : p0 + j step < n 2^24}
// outputs before sample n (output j falls on sample (p0 + j step) >> 24) -- the integers resamp_cccf_execute reaches
__global__ void tapstream_kernel(ResampP rs, long long N, TileRec *out)
{
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long ntiles = (N + TS - 1) / TS;
    if (n >= ntiles * TS) return;
    const unsigned long long pos = (unsigned long long)n << 24, p0 = rs.phase, step = rs.step;
    const unsigned long long k = pos > p0 ? (pos - p0 + step - 1) / step : 0ull;
    const unsigned long long P = p0 + k * step - pos;                       // in [0, step)
    const bool emit = P <= 0x00ffffffull && n < N;
    const unsigned cnt = (unsigned)(P >> 24), f = ((unsigned)P & 0xffffffu) >> (24 - rs.bits);
    const float h = cnt < (unsigned)rs.sublen ? rs.bank[f * rs.sublen + (rs.sublen - 1 - (int)cnt)] : 0.f;
    const bool restart = !(P < step - (1ull << 24) || n == 0);
    TileRec &r = out[n / TS];
    const int j = (int)(n % TS);
    r.tap[j] = h; r.keep[j] = restart ? 0.f : 1.f;
    // (TS == 16: one half-warp per tile)
    const unsigned half = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;
    const unsigned m = __ballot_sync(0xffffffffu, emit) & half, rr = __ballot_sync(0xffffffffu, restart) & half;
    if (j == 0) { r.emit = m ? (__ffs(m) - 1) & 15 : -1; r.gen = (m | rr) ? 1 : 0; r.pad[0] = r.pad[1] = 0; }
}

template <int NS, int NST>
__global__ void __launch_bounds__(128, NST == 3 ? 7 : 1) lane2_kernel(const __grid_constant__ SeqArgs a)
{
    static_assert(sizeof(WarpAux<NST>) % 16 == 0, "records are read 16 bytes at a time");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    smem += (1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u;     // swizzled boxes: 1024-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);              // (warp-uniform for the compiler: TMA operands stay in uniform registers)
    unsigned char *tiles = smem + (size_t)wid * (NST * WARP_TILE);
    WarpAux<NST> &ws = *(WarpAux<NST> *)(smem + (size_t)nw * (NST * WARP_TILE) + (size_t)wid * sizeof(WarpAux<NST>));

    const int c = lane >> 1, comp = lane & 1;
    const long long row0 = ((long long)blockIdx.x * nw + wid) * 16;
    if (row0 >= a.C) return;                                            // (whole warp; no CTA-wide barrier below)
    const long long chl = row0 + c, CT = a.Ctot, N = a.n;
    const bool act = chl < a.C;
    const long long gch = a.ch0 + (act ? chl : 0);

    // ---- per-channel state: this lane's component of v1, v2 of every section ----
    float v1[NS], v2[NS], acc;
#pragma unroll
    for (int s = 0; s < NS; s++) {
        v1[s] = ((const float *)(a.iir.v + (2 * s + 0) * CT + gch))[comp];
        v2[s] = ((const float *)(a.iir.v + (2 * s + 1) * CT + gch))[comp];
    }
    const int L = a.rs.sublen;
    if (lane == 0) { for (int i = 0; i < NST; i++) mbar_init(&ws.bar[i], 1); mbar_init_fence(); }     // (each warp owns its barriers)
    __syncwarp();

    // the first output's window may start before this call; that part comes from the ring (dotprod_cccf arithmetic:
    // each product is rounded, then added -- oldest sample first)
    {
        const long long nnext = a.rs.phase >> 24;
        const unsigned f = (a.rs.phase & 0xffffffu) >> (24 - a.rs.bits);
        float ar = 0.f;
        for (long long j = nnext - (L - 1); j < 0; j++) {
            const int slot = (int)(((long long)a.rs.count + j + 4LL * L) % L);
            const float h = __ldg(a.rs.bank + f * L + (int)(j - nnext + L - 1));
            const float w = ((const float *)(a.rs.ring + slot * CT + gch))[comp];
            ar = __fadd_rn(ar, __fmul_rn(h, w));
        }
        acc = ar;
    }

    const unsigned tile_sh = (unsigned)__cvta_generic_to_shared(tiles);
    const unsigned rec_sh = (unsigned)__cvta_generic_to_shared(&ws.rec[0]);
    const TileRec *recs = (const TileRec *)a.tapstream;
    const int irow0 = (int)row0;
    auto load_tile = [&](int t, int stage) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&ws.bar[stage], WARP_TILE + (unsigned)sizeof(TileRec));
            tma_load_2d(tile_sh + stage * WARP_TILE, &a.tmap, t * (TS * 2), irow0, &ws.bar[stage]);
            bulk_g2s(rec_sh + stage * (unsigned)sizeof(TileRec), recs + t, (unsigned)sizeof(TileRec), &ws.bar[stage]);
        }
    };

    long long kout = 0;
    float *yf = (float *)a.y;
    auto emit_out = [&](float v) {
        if (act) {
            const long long o = a.out_tmajor ? kout * a.out_pitch + chl : chl * a.out_pitch + kout;
            yf[2 * o + comp] = v;
        }
    };
    // byte offset of this lane's component of sample j within its row: 16-byte chunks are permuted by the row's address
    // bits 7..9 under the 128-byte swizzle (tiles are 1024-byte aligned, so that is row & 7)
    // (rows are 128-byte aligned, so the permuted chunk is an XOR on address bits 4..6: one LOP3 per chunk and tile)
    const unsigned swz = (unsigned)(c & 7);
    unsigned rowbase = (unsigned)__cvta_generic_to_shared(tiles) + c * ROWB + comp * 4 + (swz << 4);
    asm volatile("" : "+r"(rowbase));                                   // (kept in its register: ptxas otherwise rebuilds it from %tid every tile)

    // ---- stream the tiles ----
    const int ntiles = (int)((N + TS - 1) / TS);                        // (2 N fits an int32 tensor-map coordinate)
    const int nfast = (N - L) > 0 ? (int)((N - L) / TS) : 0;            // complete tiles that need no ring save
    for (int p = 0; p < NST - 1; p++) if (p < ntiles) load_tile(p, p);
    int stage = 0; unsigned parity = 0;                                 // every barrier of the ring completes once per lap
#pragma unroll 1
    for (int t = 0; t < ntiles; t++) {
        mbar_wait(&ws.bar[stage], parity);
        __syncwarp();                              // tile t has landed; every lane is done with tile t-1
        {
            const int sn = stage == 0 ? NST - 1 : stage - 1;
            if (t + NST - 1 < ntiles) load_tile(t + NST - 1, sn);
        }
        const unsigned rowp = rowbase + stage * WARP_TILE;
        const TileRec &rec = ws.rec[stage];
        const int e = rec.emit;
        if (t < nfast) {
            // Skewed cascade: at step k section s works on sample k-s -- NS independent chains, written operation by
            // operation across the sections so that dependent operations sit NS instructions apart.  Tiles in which
            // no output falls and no dot product restarts (most of them) accumulate with FMUL + FADD and plain taps.
            auto body = [&](auto general) {
                constexpr bool GEN = decltype(general)::value;
                float xs[TS];
#pragma unroll
                for (int j = 0; j < TS; j += 2) {
                    const unsigned p = rowp ^ (((unsigned)j >> 1) << 4);
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(xs[j]) : "r"(p) : "memory");
                    asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(xs[j + 1]) : "r"(p) : "memory");
                }
                float tp[TS], kp[TS];
#pragma unroll
                for (int j = 0; j < TS; j += 4) {
                    const float4 q = *(const float4 *)&rec.tap[j];
                    tp[j] = q.x; tp[j + 1] = q.y; tp[j + 2] = q.z; tp[j + 3] = q.w;
                    if constexpr (GEN) {
                        const float4 w = *(const float4 *)&rec.keep[j];
                        kp[j] = w.x; kp[j + 1] = w.y; kp[j + 2] = w.z; kp[j + 3] = w.w;
                    }
                }
                float yy[NS], outv = 0.f;
#pragma unroll
                for (int k = 0; k < TS + NS - 1; k++) {
                    float tt[NS], v0[NS], y[NS];
#pragma unroll
                    for (int s = NS - 1; s >= 0; s--) {
                        const int j = k - s;
                        if (j >= 0 && j < TS) tt[s] = __fmaf_rn(-a.iir.a[s][1], v1[s], s == 0 ? xs[j] : yy[s - 1]);
                    }
#pragma unroll
                    for (int s = NS - 1; s >= 0; s--) {
                        const int j = k - s;
                        if (j >= 0 && j < TS) { v0[s] = __fmaf_rn(-a.iir.a[s][2], v2[s], tt[s]); y[s] = __fmul_rn(a.iir.b[s][1], v1[s]); }
                    }
#pragma unroll
                    for (int s = NS - 1; s >= 0; s--) {
                        const int j = k - s;
                        if (j >= 0 && j < TS) y[s] = __fmaf_rn(a.iir.b[s][0], v0[s], y[s]);
                    }
#pragma unroll
                    for (int s = NS - 1; s >= 0; s--) {
                        const int j = k - s;
                        if (j >= 0 && j < TS) { y[s] = __fmaf_rn(a.iir.b[s][2], v2[s], y[s]); v2[s] = v1[s]; v1[s] = v0[s]; yy[s] = y[s]; }
                    }
                    {
                        const int j = k - (NS - 1);
                        if (j >= 0 && j < TS) {
                            // acc = acc*keep + round(tap*y): the two roundings of liquid's complex-tap dot product
                            if constexpr (GEN) {
                                acc = __fmaf_rn(acc, kp[j], __fmul_rn(tp[j], yy[NS - 1]));
                                if (j == e) outv = acc;
                            } else {
                                acc = __fadd_rn(acc, __fmul_rn(tp[j], yy[NS - 1]));
                            }
                        }
                    }
                }
                if constexpr (GEN) { if (e >= 0) { emit_out(outv); kout++; } }
            };
            if (rec.gen) body(std::true_type{}); else body(std::false_type{});
        } else {
            // the call's last tiles: sample by sample, saving the newest L filtered samples to the history ring
            const long long n0 = (long long)t * TS;
            const int nv = (int)((N - n0) < TS ? (N - n0) : TS);
#pragma unroll 1
            for (int j = 0; j < nv; j++) {
                const float2 tkj = make_float2(rec.tap[j], rec.keep[j]);
                float x;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"((rowp ^ (((unsigned)j >> 1) << 4)) + (unsigned)(j & 1) * 8u) : "memory");
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const float tt = __fmaf_rn(-a.iir.a[s][1], v1[s], x);
                    const float v0 = __fmaf_rn(-a.iir.a[s][2], v2[s], tt);
                    float y = __fmul_rn(a.iir.b[s][1], v1[s]);
                    y = __fmaf_rn(a.iir.b[s][0], v0, y);
                    y = __fmaf_rn(a.iir.b[s][2], v2[s], y);
                    v2[s] = v1[s]; v1[s] = v0; x = y;
                }
                acc = __fmaf_rn(acc, tkj.y, __fmul_rn(tkj.x, x));
                if (n0 + j >= N - L && act) ((float *)(a.rs.ring + (int)((a.rs.count + n0 + j) % L) * CT + gch))[comp] = x;
                if (j == e) { emit_out(acc); kout++; }
            }
        }
        if (++stage == NST) { stage = 0; parity ^= 1u; }
    }

    // ---- carried state back to HBM ----
    if (act) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            ((float *)(a.iir.v + (2 * s + 0) * CT + gch))[comp] = v1[s];
            ((float *)(a.iir.v + (2 * s + 1) * CT + gch))[comp] = v2[s];
        }
    }
}

typedef void (*LaneFn)(const SeqArgs);
template <int NST> LaneFn pick2(int nsos)
{
    switch (nsos) {
    case 1: return lane2_kernel<1, NST>; case 2: return lane2_kernel<2, NST>; case 3: return lane2_kernel<3, NST>; case 4: return lane2_kernel<4, NST>;
    default: return nullptr;
    }
}
constexpr int kDeepRing = 8;

}  // namespace

int lanes_per_channel(unsigned mask, int nsos, long long nch)
{
    (void)nch;
    if (mask != (F_IIR | F_RS) || !pick2<3>(nsos)) return 0;
    return 2;
}

const char *lanes_kernel_name(int nsos, int lanes)
{
    static const char *n2[] = { "lane2_kernel<1>", "lane2_kernel<2>", "lane2_kernel<3>", "lane2_kernel<4>" };
    if (lanes == 2 && nsos >= 1 && nsos <= 4) return n2[nsos - 1];
    return "?";
}

size_t lanes_tapstream_bytes(long long n) { return (size_t)((n + TS - 1) / TS) * sizeof(TileRec); }

cudaError_t lanes_tapstream_launch(const ResampP &rs, long long n, void *buf, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const long long slots = (n + TS - 1) / TS * TS;
    tapstream_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(rs, n, (TileRec *)buf);
    return cudaGetLastError();
}

cudaError_t lanes_launch(int nsos, int lanes, const SeqArgs &a, cudaStream_t stream)
{
    if (lanes != 2 || !a.tapstream) return cudaErrorInvalidValue;
    if (nsos < 1 || nsos > 4) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const long long warps = (a.C + 15) / 16;
    // four warps per CTA (one per scheduler) and the shallow ring once the warps cover every scheduler several times;
    // otherwise single-warp CTAs spread the channels over as many SMs as possible, each with a deep ring
    const bool full = warps >= 148 * 4 * 5;
    const int nw = full ? 4 : 1;
    LaneFn fn = full ? pick2<3>(nsos) : pick2<kDeepRing>(nsos);
    const size_t smem = 1024 + (size_t)nw * (full ? 3 * WARP_TILE + sizeof(WarpAux<3>) : kDeepRing * WARP_TILE + sizeof(WarpAux<kDeepRing>));
    cudaError_t rc = cudaFuncSetAttribute((const void *)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const unsigned grid = (unsigned)((warps + nw - 1) / nw);
    fn<<<grid, 32 * nw, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
