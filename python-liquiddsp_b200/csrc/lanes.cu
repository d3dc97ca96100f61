// lanes.cu -- the full-rate front of the receiver (biquad cascade -> decimating resampler) with the REAL and the
// IMAGINARY part of a channel on two adjacent lanes and, for few channels, the cascade's sections spread over further
// lane pairs as a systolic pipeline.
//
// Replaces the per-sample loops of ComplexIIRFilter::execute (iirfilter.hpp:292-298, iirfilt_crcf_execute_block) and
// ComplexResampler::execute (resampler.hpp:160-172, resamp_cccf_execute) for the README chain; same operations on the
// same operands in the same order as seq_kernel<F_IIR | F_RS> / front2_kernel, so the results are bit-identical.
//
// Why lanes instead of register pairs.  The filters have real coefficients, so the two components of a complex sample
// never meet: a channel is two independent real recurrences.  One component per lane makes every operation a scalar
// FFMA whose coefficient comes from the constant bank (two register operands: 1.08 cycles per instruction,
// tools/ubench_rf.cu), makes 4096 warps out of 65536 channels -- seven on every scheduler, an even load, where the
// packed two-channels-per-thread kernel (front.cu) puts two warps on 432 schedulers and one on 160 -- and gives each
// scheduler seven instruction streams to cover tile boundaries with (ncu: issue slots 90 % busy against 44 %).
//
// The cascade as a software pipeline.  A biquad is a two-operation recurrence (t = x - a1 v1, v0 = t - a2 v2) followed by
// three feed-forward operations (y = b1 v1, += b0 v0, += b2 v2), and the warp issues in order: written sample by sample
// every operation waits for the one before it (ncu: 1.2 stalled cycles per instruction on fixed-latency dependencies).
// So a step is split into stages that run in DIFFERENT steps: stage A of section s (the recurrence) works on sample
// k - lag(s) at step k, stage B (the output) finishes that sample one step later, the next section's stage A picks it up
// one step after that, and the resampler's multiply-accumulate (stage C) follows the last section's stage B by one
// step.  Within a step the stages of all sections are independent chains, two or three operations deep, and they are
// written level by level across each other -- same operations, same operands, same results, no tile-boundary ramp.
// With few channels the sections are spread over G = 2 or 4 lane pairs (G * 2 lanes per channel): a lane pair hands its
// output to the next one by warp shuffle through a four-slot ring, so the shuffle's latency is off the per-step path.
// The first and the last tiles of a call run the same step with validity predicates (a stage commits only for samples
// 0 <= k - lag < N).
//
// Staging: one elected lane per warp issues one cp.async.bulk.tensor.2d per tile for the warp's [rows x 128 B] box
// (128-byte swizzle, per-warp mbarrier ring: 3 deep when the machine is full, 8 deep when few warps must keep the HBM
// pipe busy) plus one 208-byte record of everything that is the same for all channels (tap stream, below).  Warps never
// wait for each other: no CTA barrier anywhere.
#include <cuda_runtime.h>
#include <cuda.h>
#include <cstdlib>
#include <type_traits>
#include <utility>
#include "params.h"
#include "devmath.cuh"
#include "lanes.h"

namespace lqb {
namespace {

constexpr int TS = 16;
constexpr int ROWB = TS * 8;                       // bytes per staged row (dense, swizzled)
constexpr int DX = 4;                              // slots of the hand-off ring between lane pairs (a value is read DX steps after it was written)

// What is the same for every channel -- which polyphase tap multiplies a sample, whether the dot product restarts on it,
// where in a tile an output falls -- is worked out once per call by tapstream_kernel (closed form of liquid's uint32 phase
// recurrence) and reaches each warp as one 208-byte record per tile, copied by the same mbarrier transaction as the tile.
// Records are in STEP space: slot j of record t belongs to step k = 16 t + j, i.e. to sample k - lag of the last section.
struct TileRec {
    float tap[TS];                                 // tap that multiplies the sample (0 outside every window and outside the call)
    float keep[TS];                                // 0 on the sample after an output (the accumulator restarts), else 1
    float cap[TS];                                 // 1 on the step an output falls on, else 0 (out = cap * acc + out: one FFMA instead of compare + select)
    int emit;                                      // step of the tile an output falls on, or -1 (step >= TS * 2^24: at most one)
    int gen;                                       // 1 when the tile holds an output or a restart, 0: plain accumulation
    int pad[2];
};
static_assert(sizeof(TileRec) == 208, "tile records are copied 16 bytes at a time");
// a stage of the ring = TPS consecutive tiles under ONE mbarrier (TPS tensor boxes + one copy of their TPS records): with
// few channels the per-stage bookkeeping (barrier wait, TMA issue) is what a lone warp cannot hide, so it is paid once
// per 64 samples there; the full machine stages single tiles (shared memory limits it to three 2 KB stages per warp)
template <int NST, int TPS> struct WarpAux {
    TileRec rec[NST][TPS];
    unsigned long long bar[NST];
    unsigned long long pad[NST & 1 ? 1 : 2];
};

__device__ __forceinline__ bool elect_one()
{
    unsigned p;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(p));
    return p != 0;
}
// one non-blocking probe of an mbarrier phase (the result is consumed a tile later, so its latency hides under the tile)
__device__ __forceinline__ bool mbar_try(void *bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}"
                 : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// phase of the resampler at sample n of the call: P(n) = p0 + k(n) step - n 2^24 with k(n) = #{j >= 0 : p0 + j step < n 2^24}
// outputs before sample n (output j falls on sample (p0 + j step) >> 24) -- the integers resamp_cccf_execute reaches
__global__ void tapstream_kernel(ResampP rs, long long N, int lag, TileRec *out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // step
    const long long ntiles = ((N + lag + TS - 1) / TS + 3) / 4 * 4;             // padded to whole stages (records past the call: tap 0)
    if (k >= ntiles * TS) return;
    const long long n = k - lag;
    const bool in_call = n >= 0 && n < N;
    const unsigned long long pos = (unsigned long long)(in_call ? n : 0) << 24, p0 = rs.phase, step = rs.step;
    const unsigned long long kk = pos > p0 ? (pos - p0 + step - 1) / step : 0ull;
    const unsigned long long P = p0 + kk * step - pos;                      // in [0, step)
    const bool emit = in_call && P <= 0x00ffffffull;
    const unsigned cnt = (unsigned)(P >> 24), f = ((unsigned)P & 0xffffffu) >> (24 - rs.bits);
    const float h = (in_call && cnt < (unsigned)rs.sublen) ? rs.bank[f * rs.sublen + (rs.sublen - 1 - (int)cnt)] : 0.f;
    // the sample after an output starts a new dot product (never the first sample of a call: there the accumulator holds
    // the history ring's contribution)
    const bool restart = in_call && !(P < step - (1ull << 24) || n == 0);
    TileRec &r = out[k / TS];
    const int j = (int)(k % TS);
    r.tap[j] = h; r.keep[j] = restart ? 0.f : 1.f; r.cap[j] = emit ? 1.f : 0.f;
    const unsigned half = (threadIdx.x & 16) ? 0xffff0000u : 0x0000ffffu;     // (TS == 16: one half-warp per tile)
    const unsigned m = __ballot_sync(0xffffffffu, emit) & half, rr = __ballot_sync(0xffffffffu, restart) & half;
    if (j == 0) { r.emit = m ? (__ffs(m) - 1) & 15 : -1; r.gen = (m | rr) ? 1 : 0; r.pad[0] = r.pad[1] = 0; }
}

template <class F, int... I> __device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>)
{
    (f(std::integral_constant<int, I>{}), ...);
}
template <int N, class F> __device__ __forceinline__ void static_for(F &&f) { static_for_impl(f, std::make_integer_sequence<int, N>{}); }

// SPL sections per lane, G lane pairs per channel (NS = SPL * G sections), NST stages of TPS tiles in the staging ring
template <int SPL, int G, int NST, int TPS>
__global__ void __launch_bounds__(128, TPS == 1 ? 7 : 1) lanes_kernel(const __grid_constant__ SeqArgs a)
{
    static_assert(sizeof(WarpAux<NST, TPS>) % 16 == 0, "records are read 16 bytes at a time");
    constexpr int LPC = 2 * G, CPW = 32 / LPC, WT = CPW * ROWB;
    // stage A of local section i of lane pair g runs lagA = g * LAGG + 2 i steps behind the input; the resampler's
    // multiply-accumulate LAG_RS steps behind it
    constexpr int LAGG = 2 * (SPL - 1) + DX + 1;
    constexpr int LAG_RS = (G - 1) * LAGG + 2 * (SPL - 1) + 2;
    static_assert(LAG_RS <= 2 * TS && TS % DX == 0, "ring slots are addressed by the step's position in its tile");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw;
    smem += (1024u - ((unsigned)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u;     // swizzled boxes: 1024-byte aligned
    const int tid = threadIdx.x, lane = tid & 31, nw = blockDim.x >> 5;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);              // (warp-uniform for the compiler: TMA operands stay in uniform registers)
    // shared memory: every warp's NST tiles first (swizzled TMA boxes), then per warp the tile records and the mbarriers
    constexpr int WTA = (WT + 1023) / 1024 * 1024;                      // a stage, padded so that every box starts 1024-byte aligned
    constexpr int STB = TPS * WTA;                                      // bytes of a stage's tiles
    unsigned char *tiles = smem + (size_t)wid * (NST * STB);
    WarpAux<NST, TPS> &ws = *(WarpAux<NST, TPS> *)(smem + (size_t)nw * (NST * STB) + (size_t)wid * sizeof(WarpAux<NST, TPS>));

    const int c = lane / LPC, g = (lane >> 1) % G, comp = lane & 1;
    const bool first_grp = g == 0, last_grp = g == G - 1;
    const long long row0 = ((long long)blockIdx.x * nw + wid) * CPW;
    if (row0 >= a.C) return;                                            // (whole warp; no CTA-wide barrier below)
    const long long chl = row0 + c, CT = a.Ctot, N = a.n;
    const bool act = chl < a.C;
    const long long gch = a.ch0 + (act ? chl : 0);

    // ---- this lane's sections: coefficients (registers only when the section depends on the lane) and state ----
    float v1[SPL], v2[SPL], cna1[SPL], cna2[SPL], cb0[SPL], cb1[SPL], cb2[SPL];
#pragma unroll
    for (int i = 0; i < SPL; i++) {
        const int s = g * SPL + i;
        v1[i] = ((const float *)(a.iir.v + (2 * s + 0) * CT + gch))[comp];
        v2[i] = ((const float *)(a.iir.v + (2 * s + 1) * CT + gch))[comp];
        if constexpr (G > 1) { cna1[i] = -a.iir.a[s][1]; cna2[i] = -a.iir.a[s][2]; cb0[i] = a.iir.b[s][0]; cb1[i] = a.iir.b[s][1]; cb2[i] = a.iir.b[s][2]; }
    }
    auto NA1 = [&](int i) -> float { if constexpr (G == 1) return a.lc[i]; else return cna1[i]; };
    auto NA2 = [&](int i) -> float { if constexpr (G == 1) return a.lc[4 + i]; else return cna2[i]; };
    auto B1 = [&](int i) -> float { if constexpr (G == 1) return a.lc[8 + i]; else return cb1[i]; };
    auto B0 = [&](int i) -> float { if constexpr (G == 1) return a.lc[12 + i]; else return cb0[i]; };
    auto B2 = [&](int i) -> float { if constexpr (G == 1) return a.lc[16 + i]; else return cb2[i]; };
    const int L = a.rs.sublen;
    if (lane == 0) { for (int i = 0; i < NST; i++) mbar_init(&ws.bar[i], 1); mbar_init_fence(); }     // (each warp owns its barriers)
    __syncwarp();

    // the first output's window may start before this call; that part comes from the ring (dotprod_cccf arithmetic:
    // each product is rounded, then added -- oldest sample first)
    float acc = 0.f;
    if (last_grp) {
        const long long nnext = a.rs.phase >> 24;
        const unsigned f = (a.rs.phase & 0xffffffu) >> (24 - a.rs.bits);
        // (eight loads in flight at a time, one modulo for the whole walk: as a plain loop this was ~40 dependent global-memory
        // round trips, 20 us per launch -- as much as 1000 samples of a lone warp's work)
        const long long j0 = nnext - (L - 1);
        const int cnt = j0 < 0 ? (int)(-j0) : 0;                        // ring samples in the window (at most L - 1)
        int slot = (int)(((long long)a.rs.count + j0 + 4LL * L) % L);
        const float *hb = a.rs.bank + f * L;
        for (int i0 = 0; i0 < cnt; i0 += 8) {
            float h[8], w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const bool on = i0 + i < cnt;
                h[i] = on ? __ldg(hb + i0 + i) : 0.f;
                w[i] = on ? ((const float *)(a.rs.ring + slot * CT + gch))[comp] : 0.f;
                slot = slot + 1 == L ? 0 : slot + 1;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) if (i0 + i < cnt) acc = __fadd_rn(acc, __fmul_rn(h[i], w[i]));
        }
    }

    const unsigned tile_sh = (unsigned)__cvta_generic_to_shared(tiles);
    const unsigned rec_sh = (unsigned)__cvta_generic_to_shared(&ws.rec[0][0]);
    const TileRec *recs = (const TileRec *)a.tapstream;
    const int irow0 = (int)row0;
    const int ntx = (int)((N + TS - 1) / TS);                           // tiles that hold samples (2 N fits an int32 tensor-map coordinate)
    const int ntiles = (int)((N + LAG_RS + TS - 1) / TS);               // tiles of steps: the pipeline drains LAG_RS steps past the call
    const int nstg = (ntiles + TPS - 1) / TPS;                          // stages of the call (the tap stream is padded to whole stages)
    auto load_stage = [&](int sg, int stage, unsigned soff) {
        if (elect_one()) {
            const int t0 = sg * TPS;
            const int nb = ntx - t0 < TPS ? (ntx - t0 > 0 ? ntx - t0 : 0) : TPS;       // tiles of the stage that hold samples
            mbar_arrive_expect_tx(&ws.bar[stage], (unsigned)nb * WT + (unsigned)(TPS * sizeof(TileRec)));
#pragma unroll
            for (int q = 0; q < TPS; q++)
                if (q < nb) tma_load_2d(tile_sh + soff + q * WTA, &a.tmap, (t0 + q) * (TS * 2), irow0, &ws.bar[stage]);
            bulk_g2s(rec_sh + stage * (unsigned)(TPS * sizeof(TileRec)), recs + t0, (unsigned)(TPS * sizeof(TileRec)), &ws.bar[stage]);
        }
    };

    // where this lane's next output goes: a running pointer (one 64-bit add per output instead of an index product); the
    // stride and the write flag are pinned in registers (ptxas otherwise rebuilds both from the constant bank at every output)
    float *yp = (float *)a.y + 2 * (a.out_tmajor ? chl : chl * a.out_pitch) + comp;
    unsigned ystride = (unsigned)(a.out_tmajor ? 8 * a.out_pitch : 8);          // bytes (a time-major row is C * 8 bytes < 4 GB)
    unsigned writer = (act && last_grp) ? 1u : 0u;
    asm volatile("" : "+r"(ystride), "+r"(writer));
    auto emit_out = [&](float v) {
        if (writer) *yp = v;
        yp = (float *)((char *)yp + ystride);
    };
    // byte address of this lane's component of sample j within its row: 16-byte chunks are permuted by the row's address
    // bits 7..9 under the 128-byte swizzle (rows are 128-byte aligned, so the permuted chunk is an XOR on address bits 4..6)
    unsigned rowbase = tile_sh + c * ROWB + comp * 4;
    rowbase += ((rowbase >> 7) & 7u) << 4;
    // the eight chunk addresses of the row, kept in registers for the whole call: a load is then [register + stage offset
    // (uniform) + immediate] with no per-tile address arithmetic (stage offsets are multiples of 1 KB, so they commute with the XOR)
    // (they come back through a volatile shared-memory load of the warp's own, still empty, ring: an XOR with an immediate
    // is something ptxas rematerialises at every use to save a register -- six LOP3 per tile again -- a volatile load is not)
    unsigned xr[8];
    {
        unsigned *scr = (unsigned *)tiles + lane * 8;
#pragma unroll
        for (int m = 0; m < 8; m++) scr[m] = rowbase ^ ((unsigned)m << 4);
        __syncwarp();
        const unsigned sa = (unsigned)__cvta_generic_to_shared(scr);
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(xr[0]), "=r"(xr[1]), "=r"(xr[2]), "=r"(xr[3]) : "r"(sa) : "memory");
        asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(xr[4]), "=r"(xr[5]), "=r"(xr[6]), "=r"(xr[7]) : "r"(sa + 16u) : "memory");
        __syncwarp();
        fence_async_smem();                        // the ring is written by the TMA engine next
    }
    auto ldx = [&](unsigned soff, auto jc) -> float {
        constexpr int j = decltype(jc)::value;
        float v;
        if (j & 1) asm volatile("ld.shared.f32 %0, [%1+8];" : "=f"(v) : "r"(xr[j >> 1] + soff) : "memory");
        else       asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(xr[j >> 1] + soff) : "memory");
        return v;
    };
    auto ldx_dyn = [&](unsigned soff, int j) -> float {                 // (the call's first and last tiles)
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(((rowbase + soff) ^ (((unsigned)j >> 1) << 4)) + (unsigned)(j & 1) * 8u) : "memory");
        return v;
    };

    // pipeline registers: sv* = what stage A left for stage B of the same section; h[i] = output of local section i for
    // section i + 1; hx[slot] = ring towards the next lane pair; ylp = last section's output for the resampler
    // upf[slot] = the previous lane pair's output fetched two steps ahead of its use (the shuffle's latency off the step)
    // (stage B reads v1 / v2 themselves for the two newer delay elements: after a committed stage A they are exactly what it
    // left behind -- v0 and the old v1 -- and after an uncommitted one the sample is outside the call and its output unused)
    float sv2[SPL], h[SPL], hx[DX], upf[2] = { 0.f, 0.f }, ylp = 0.f;
#pragma unroll
    for (int i = 0; i < SPL; i++) { sv2[i] = h[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < DX; i++) hx[i] = 0.f;

    // One step.  SLOT: the step's position in its tile (static: ring slots, emit test); x: the sample section 0 reads;
    // tap / keep: the tile record's entries of this step; PRED: stages commit only inside the call (first and last
    // tiles), GEN: the accumulator may restart and an output may fall on this step.
    float outv = 0.f;
    auto step = [&](auto slotc, auto predc, auto genc, float x, float tap, float keep, float cap, int e, long long k) {
        [[maybe_unused]] constexpr int SLOT = decltype(slotc)::value;
        constexpr bool PRED = decltype(predc)::value, GEN = decltype(genc)::value;
        float in[SPL], tt[SPL], v0[SPL], y[SPL];
        if constexpr (G > 1) {
            // what the previous lane pair's stage B wrote DX steps ago: fetched two steps back from the slot it still occupied
            in[0] = first_grp ? x : upf[SLOT % 2];
            upf[SLOT % 2] = __shfl_up_sync(0xffffffffu, hx[(SLOT + 2) % DX], 2);
        } else {
            in[0] = x;
        }
#pragma unroll
        for (int i = 1; i < SPL; i++) in[i] = h[i - 1];                            // stage B of section i - 1, previous step
        // level 1 of the three stages
#pragma unroll
        for (int i = SPL - 1; i >= 0; i--) tt[i] = __fmaf_rn(NA1(i), v1[i], in[i]);
#pragma unroll
        for (int i = SPL - 1; i >= 0; i--) y[i] = __fmul_rn(B1(i), v2[i]);
        const float p = __fmul_rn(tap, ylp);
        // level 2
#pragma unroll
        for (int i = SPL - 1; i >= 0; i--) v0[i] = __fmaf_rn(NA2(i), v2[i], tt[i]);
#pragma unroll
        for (int i = SPL - 1; i >= 0; i--) y[i] = __fmaf_rn(B0(i), v1[i], y[i]);
        // acc = acc*keep + round(tap*y): the two roundings of liquid's complex-tap dot product; tiles in which no output
        // falls and no dot product restarts (most of them) accumulate with FMUL + FADD
        bool rs_on = true;
        if constexpr (PRED) { const long long nl = k - LAG_RS; rs_on = nl >= 0 && nl < N; }
        if (rs_on) {
            if constexpr (GEN || PRED) acc = __fmaf_rn(acc, keep, p); else acc = __fadd_rn(acc, p);
        }
        if constexpr (PRED) {
            const long long nl = k - LAG_RS;
            if (rs_on && nl >= N - L && act && last_grp) ((float *)(a.rs.ring + (int)((a.rs.count + nl) % L) * CT + gch))[comp] = ylp;
            if (rs_on && (int)(k & (TS - 1)) == e) emit_out(acc);                 // (tiles start on multiples of TS)
        } else if constexpr (GEN) {
            // the output falls on the step whose cap is 1: 1 * acc + 0 is acc exactly (acc is never -0: sums start from +0)
            outv = __fmaf_rn(cap, acc, outv);
        }
        // level 3, then the hand-overs
#pragma unroll
        for (int i = SPL - 1; i >= 0; i--) {
            y[i] = __fmaf_rn(B2(i), sv2[i], y[i]);
            sv2[i] = v2[i];
            bool commit = true;
            if constexpr (PRED) { const long long n = k - (g * LAGG + 2 * i); commit = n >= 0 && n < N; }
            if (commit) { v2[i] = v1[i]; v1[i] = v0[i]; }
            h[i] = y[i];
        }
        if constexpr (G > 1) hx[SLOT % DX] = y[SPL - 1];
        ylp = y[SPL - 1];
    };

    // ---- stream the tiles ----
    // fast tiles: every stage inside the call for all 16 steps and no history-ring save
    const long long lim = (N - L + LAG_RS) < N ? (N - L + LAG_RS) : N;
    const int tfast0 = (LAG_RS + TS - 1) / TS, tfast1 = lim > 0 ? (int)(lim / TS) : 0;
    const unsigned tfast_n = tfast1 > tfast0 ? (unsigned)(tfast1 - tfast0) : 0u;       // fast tiles: one unsigned compare per tile
    const bool no_stage_body = a.lanes_flags & 1;                       // (A/B: LQB_NO_STAGEBODY=1)
    for (int p = 0; p < NST; p++) if (p < nstg) load_stage(p, p, (unsigned)(p * STB));
    int stage = 0; unsigned parity = 0;                                 // every barrier of the ring completes once per lap
    unsigned soff = 0;                                                  // byte offset of the stage within the warp's ring (running: no multiply per tile)
    bool ready = false;
#pragma unroll 1
    for (int sg = 0; sg < nstg; sg++) {
        if (!ready) mbar_wait(&ws.bar[stage], parity);                  // stage sg has landed (each lane sees the phase flip itself)
        // probe the next stage's barrier now: the answer is needed only after this stage's arithmetic (one vote makes it
        // warp-uniform, so the loop's bookkeeping stays in uniform registers)
        const int stage_n = stage + 1 == NST ? 0 : stage + 1;
        const unsigned parity_n = stage + 1 == NST ? parity ^ 1u : parity;
        // (only where a lone warp runs a scheduler: with seven warps per scheduler the wait costs nothing and the probe does)
        bool probe = false;
        if constexpr (TPS > 1) probe = (sg + 1 < nstg) ? mbar_try(&ws.bar[stage_n], parity_n) : false;
        // A lone warp pays for everything around a tile body in full (flags, fast-tile test, branches, the loads' latency at the
        // head of the body: 60 % of a one-channel call's cycles, ncu), so a stage whose TPS tiles all lie inside the call runs as
        // ONE unrolled body: the next tile's samples and taps are fetched before the current tile's steps, every tile takes the
        // general step (capture stream, restart multiplier), outputs are stored after the body in order.
        bool whole = false;
        if constexpr (TPS > 1) {
            const int t0 = sg * TPS;
            whole = (unsigned)(t0 - tfast0) < tfast_n && (unsigned)(t0 + TPS - 1 - tfast0) < tfast_n && !no_stage_body;
        }
        if (whole) {
            if constexpr (TPS > 1) {
                float xs[2][TS], tp[2][TS], kp[2][TS], cp[2][TS], outs[TPS]; int es[TPS];
                auto fetch = [&](auto qc) {
                    constexpr int q = decltype(qc)::value, b = q & 1;
                    const unsigned rowq = soff + (unsigned)(q * WTA);
                    const TileRec &rq = ws.rec[stage][q];
                    if (G == 1 || first_grp) {
                        static_for<TS>([&](auto jc) { xs[b][decltype(jc)::value] = ldx(rowq, jc); });
                    } else {
#pragma unroll
                        for (int j = 0; j < TS; j++) xs[b][j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < TS; j += 4) {
                        const float4 q4 = *(const float4 *)&rq.tap[j], w = *(const float4 *)&rq.keep[j], u = *(const float4 *)&rq.cap[j];
                        tp[b][j] = q4.x; tp[b][j + 1] = q4.y; tp[b][j + 2] = q4.z; tp[b][j + 3] = q4.w;
                        kp[b][j] = w.x; kp[b][j + 1] = w.y; kp[b][j + 2] = w.z; kp[b][j + 3] = w.w;
                        cp[b][j] = u.x; cp[b][j + 1] = u.y; cp[b][j + 2] = u.z; cp[b][j + 3] = u.w;
                    }
                    es[q] = rq.emit;
                };
                fetch(std::integral_constant<int, 0>{});
                static_for<TPS>([&](auto qc) {
                    constexpr int q = decltype(qc)::value, b = q & 1;
                    if constexpr (q + 1 < TPS) fetch(std::integral_constant<int, q + 1>{});
                    outv = 0.f;
                    static_for<TS>([&](auto jc) {
                        constexpr int j = decltype(jc)::value;
                        step(jc, std::false_type{}, std::true_type{}, xs[b][j], tp[b][j], kp[b][j], cp[b][j], 0, 0);
                    });
                    outs[q] = outv;
                });
                static_for<TPS>([&](auto qc) { constexpr int q = decltype(qc)::value; if (es[q] >= 0) emit_out(outs[q]); });
            }
        } else {
        int e_nx = -1, gen_nx = 0; bool flags_nx = false;      // the next tile's flags, read from its record one tile ahead
#pragma unroll 1
        for (int q = 0; q < TPS; q++) {
        const int t = sg * TPS + q;
        if (t >= ntiles) break;
        const unsigned rowp = soff + (unsigned)(q * WTA);                   // (uniform) offset of the tile within the warp's ring
        const TileRec &rec = ws.rec[stage][q];
        const int e = flags_nx ? e_nx : rec.emit, gen = flags_nx ? gen_nx : rec.gen;
        if constexpr (TPS > 1) {
            flags_nx = q + 1 < TPS;
            if (flags_nx) { const int2 f = *(const int2 *)&ws.rec[stage][q + 1].emit; e_nx = f.x; gen_nx = f.y; }
        }
        if ((unsigned)(t - tfast0) < tfast_n) {
            auto body = [&](auto general) {
                constexpr bool GEN = decltype(general)::value;
                float xs[TS];
                if (G == 1 || first_grp) {
                    static_for<TS>([&](auto jc) { xs[decltype(jc)::value] = ldx(rowp, jc); });
                } else {
#pragma unroll
                    for (int j = 0; j < TS; j++) xs[j] = 0.f;
                }
                float tp[TS], kp[TS], cp[TS];
#pragma unroll
                for (int j = 0; j < TS; j += 4) {
                    const float4 q4 = *(const float4 *)&rec.tap[j];
                    tp[j] = q4.x; tp[j + 1] = q4.y; tp[j + 2] = q4.z; tp[j + 3] = q4.w;
                    if constexpr (GEN) {
                        const float4 w = *(const float4 *)&rec.keep[j], u = *(const float4 *)&rec.cap[j];
                        kp[j] = w.x; kp[j + 1] = w.y; kp[j + 2] = w.z; kp[j + 3] = w.w;
                        cp[j] = u.x; cp[j + 1] = u.y; cp[j + 2] = u.z; cp[j + 3] = u.w;
                    } else {
                        kp[j] = kp[j + 1] = kp[j + 2] = kp[j + 3] = 1.f;
                        cp[j] = cp[j + 1] = cp[j + 2] = cp[j + 3] = 0.f;
                    }
                }
                if constexpr (GEN) outv = 0.f;
                static_for<TS>([&](auto jc) {
                    constexpr int j = decltype(jc)::value;
                    step(jc, std::false_type{}, general, xs[j], tp[j], kp[j], cp[j], e, 0);
                });
                if constexpr (GEN) { if (e >= 0) emit_out(outv); }
            };
            if (gen) body(std::true_type{}); else body(std::false_type{});
        } else {
            // the call's first and last tiles: the same step with validity predicates, saving the newest L filtered samples
            // to the history ring
            const long long k0 = (long long)t * TS;
#pragma unroll 1
            for (int jo = 0; jo < TS; jo += DX) {
                static_for<DX>([&](auto jc) {
                    const int j = jo + decltype(jc)::value;
                    const long long k = k0 + j;
                    const float x = (k < N && (G == 1 || first_grp)) ? ldx_dyn(rowp, j) : 0.f;
                    step(jc, std::true_type{}, std::true_type{}, x, rec.tap[j], rec.keep[j], 0.f, e, k);
                });
            }
        }
        }
        }
        // (the vote makes the probe's answer warp-uniform; taken here, after the stage's arithmetic, so that nothing waits for it)
        if constexpr (TPS > 1) ready = __all_sync(0xffffffffu, probe) != 0;
        // every lane is done with this stage: refill it with the stage NST ahead (the stage's own addresses are at hand, so the
        // refill costs no index arithmetic; NST - 1 stages stay in flight while the next one is worked on)
        __syncwarp();
        if (sg + NST < nstg) load_stage(sg + NST, stage, soff);
        soff += STB;
        if (++stage == NST) { stage = 0; parity ^= 1u; soff = 0; }
    }

    // ---- carried state back to HBM ----
    if (act) {
#pragma unroll
        for (int i = 0; i < SPL; i++) {
            const int s = g * SPL + i;
            ((float *)(a.iir.v + (2 * s + 0) * CT + gch))[comp] = v1[i];
            ((float *)(a.iir.v + (2 * s + 1) * CT + gch))[comp] = v2[i];
        }
    }
}

typedef void (*LaneFn)(const SeqArgs);
struct Pick { LaneFn fn; int lag_rs; size_t smem_per_warp; };
template <int SPL, int G, int NST, int TPS> Pick mk()
{
    constexpr int WT = (32 / (2 * G)) * ROWB, WTA = (WT + 1023) / 1024 * 1024;
    return Pick{ lanes_kernel<SPL, G, NST, TPS>, (G - 1) * (2 * (SPL - 1) + DX + 1) + 2 * (SPL - 1) + 2, (size_t)NST * TPS * WTA + sizeof(WarpAux<NST, TPS>) };
}
constexpr int kFewNST = 4, kFewTPS = 4;            // few channels: four stages of four tiles (64 samples) per warp
Pick pick(int nsos, int lanes, bool full)
{
    if (lanes == 2) {
        switch (nsos) {
        case 1: return full ? mk<1, 1, 3, 1>() : mk<1, 1, kFewNST, kFewTPS>();
        case 2: return full ? mk<2, 1, 3, 1>() : mk<2, 1, kFewNST, kFewTPS>();
        case 3: return full ? mk<3, 1, 3, 1>() : mk<3, 1, kFewNST, kFewTPS>();
        case 4: return full ? mk<4, 1, 3, 1>() : mk<4, 1, kFewNST, kFewTPS>();
        }
    } else if (lanes == 4) {
        if (nsos == 2) return mk<1, 2, kFewNST, kFewTPS>();
        if (nsos == 4) return mk<2, 2, kFewNST, kFewTPS>();
    } else if (lanes == 8) {
        if (nsos == 4) return mk<1, 4, kFewNST, kFewTPS>();
    }
    return Pick{ nullptr, 0, 0 };
}
// four warps per CTA (one per scheduler) and the shallow ring once the warps cover every scheduler;
// otherwise single-warp CTAs spread the channels over as many SMs as possible, each with a deep ring
// (the deep variant holds 32 KB of shared memory per warp: six warps per SM.  From 888 warps on -- 14208 channels -- it would
// need a second wave, so everything above runs the shallow variant, 6 KB per warp)
bool full_machine(long long nch, int lanes) { return lanes == 2 && (nch + 15) / 16 > 148 * 6; }

}  // namespace

int lanes_per_channel(unsigned mask, int nsos, long long nch)
{
    if (mask != (F_IIR | F_RS) || nsos < 1 || nsos > 4) return 0;
    int lanes = 2;
    // few channels: spread the sections over lane pairs so that one channel's recurrence is not a single instruction stream
    if (nsos == 4) lanes = nch <= 1184 ? 8 : (nch <= 2368 ? 4 : 2);
    else if (nsos == 2) lanes = nch <= 2368 ? 4 : 2;
    if (const char *e = getenv("LQB_LANES")) { const int v = atoi(e); if ((v == 2 || v == 4 || v == 8) && pick(nsos, v, false).fn) lanes = v; }   // tuning override
    return lanes;
}

int lanes_lag(int nsos, int lanes) { return pick(nsos, lanes, false).lag_rs; }

const char *lanes_kernel_name(int nsos, int lanes)
{
    static const char *n2[] = { "lanes_kernel<1,1>", "lanes_kernel<2,1>", "lanes_kernel<3,1>", "lanes_kernel<4,1>" };
    if (lanes == 2 && nsos >= 1 && nsos <= 4) return n2[nsos - 1];
    if (lanes == 4) return nsos == 2 ? "lanes_kernel<1,2>" : "lanes_kernel<2,2>";
    if (lanes == 8) return "lanes_kernel<1,4>";
    return "?";
}

size_t lanes_tapstream_bytes(long long n) { return (size_t)(((n + 2 * TS + TS - 1) / TS + 3) / 4 * 4) * sizeof(TileRec); }      // (lag <= 2 TS; whole stages of 4 tiles)

cudaError_t lanes_tapstream_launch(const ResampP &rs, long long n, int lag, void *buf, cudaStream_t stream)
{
    if (n <= 0) return cudaSuccess;
    const long long slots = ((n + lag + TS - 1) / TS + 3) / 4 * 4 * TS;
    tapstream_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(rs, n, lag, (TileRec *)buf);
    return cudaGetLastError();
}

cudaError_t lanes_launch(int nsos, int lanes, const SeqArgs &a0, cudaStream_t stream)
{
    SeqArgs a = a0;
    for (int i = 0; i < 4; i++) {
        const bool on = i < nsos;
        a.lc[i] = on ? -a.iir.a[i][1] : 0.f; a.lc[4 + i] = on ? -a.iir.a[i][2] : 0.f;
        a.lc[8 + i] = on ? a.iir.b[i][1] : 0.f; a.lc[12 + i] = on ? a.iir.b[i][0] : 0.f; a.lc[16 + i] = on ? a.iir.b[i][2] : 0.f;
    }
    a.lanes_flags = getenv("LQB_NO_STAGEBODY") ? 1 : 0;
    if (!a.tapstream) return cudaErrorInvalidValue;
    if (a.C <= 0 || a.n <= 0) return cudaSuccess;
    const bool full = full_machine(a.C, lanes);
    const Pick p = pick(nsos, lanes, full);
    if (!p.fn) return cudaErrorInvalidValue;
    const int cpw = 32 / lanes;
    const long long warps = (a.C + cpw - 1) / cpw;
    const int nw = full ? 4 : 1;
    const size_t smem = 1024 + (size_t)nw * p.smem_per_warp;
    cudaError_t rc = cudaFuncSetAttribute((const void *)p.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    const unsigned grid = (unsigned)((warps + nw - 1) / nw);
    p.fn<<<grid, 32 * nw, smem, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace lqb
