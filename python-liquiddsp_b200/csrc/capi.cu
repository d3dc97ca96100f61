// capi.cu -- the C ABI of include/liquiddsp_b200.h: stage handles, carried state in HBM, the chain
// planner and the host / device execute paths.
//
// This file is what the reference's L2 layer (one C++ class per liquid object, src/*.hpp) becomes:
// each handle owns device state arrays instead of a liquid handle, `execute` launches kernels
// instead of calling liquid's per-sample loops, and there is no CPU path -- without a CUDA device
// every execute fails with LQB_ECUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <set>
#include <string>
#include <vector>

#include "../../include/liquiddsp_b200.h"
#include "design.hpp"
#include "params.h"
#include "seq.h"
#include "fir.h"
#include "am.h"
#include "bam.h"
#include "fmst.h"
#include "front.h"
#include "lanes.h"
#include "pipe.h"
#include "par.h"
#include "scan.h"
#include "synth.h"

namespace lqb {

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define LQB_CUDA(call)                                                                              \
    do { cudaError_t e_ = (call);                                                                   \
         if (e_ != cudaSuccess) return ::lqb::fail(LQB_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define LQB_TRY(call) do { int rc_ = (call); if (rc_ != LQB_OK) return rc_; } while (0)

// ------------------------------------------------------------------------------------ device arrays
template <class T> struct DevArr {
    T *p = nullptr; size_t n = 0;
    DevArr() {}
    DevArr(const DevArr &) = delete; DevArr &operator=(const DevArr &) = delete;
    ~DevArr() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    // State arrays are zeroed; scratch (reserve) is not.  cudaMemset / cudaMemcpy run on the legacy default stream,
    // which orders nothing against the non-blocking streams callers execute on (lqb_stream_create, torch side streams),
    // and both may return before the device has finished: settle() waits, so that whatever stream runs next sees the
    // initialised state.  State is set up once per object, so the wait is off every hot path.
    static int settle() { LQB_CUDA(cudaStreamSynchronize(cudaStreamLegacy)); return LQB_OK; }
    int alloc(size_t count, bool zeroed = true)
    {
        release();
        if (count == 0) return LQB_OK;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e != cudaSuccess) return fail(e == cudaErrorMemoryAllocation ? LQB_ENOMEM : LQB_ECUDA, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
        n = count;
        return zeroed ? zero() : LQB_OK;
    }
    int reserve(size_t count) { return count <= n ? LQB_OK : alloc(count, false); }      // scratch: contents undefined
    int zero() { if (p) { LQB_CUDA(cudaMemset(p, 0, n * sizeof(T))); return settle(); } return LQB_OK; }
    int fill(const T &v)
    {
        if (!p) return LQB_OK;
        std::vector<T> h(n, v);
        LQB_CUDA(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
        return settle();
    }
    int upload(const T *h, size_t count) { LQB_CUDA(cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice)); return settle(); }
    int download(T *h, size_t count) const { LQB_CUDA(cudaMemcpy(h, p, count * sizeof(T), cudaMemcpyDeviceToHost)); return LQB_OK; }
};

// the oscillator's 1024-entry (sin, cos) table, one copy per device
static std::mutex g_tab_mu;
static std::map<int, float2 *> g_sincos;
static int sincos_table(const float2 **out)
{
    int dev = 0; LQB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto it = g_sincos.find(dev);
    if (it == g_sincos.end()) {
        std::vector<float> s = design::nco_sintab();
        std::vector<float2> t(1024);
        for (int i = 0; i < 1024; i++) t[i] = make_float2(s[i], s[(i + 256) & 0x3ff]);
        float2 *d = nullptr;
        LQB_CUDA(cudaMalloc((void **)&d, 1024 * sizeof(float2)));
        LQB_CUDA(cudaMemcpy(d, t.data(), 1024 * sizeof(float2), cudaMemcpyHostToDevice));
        it = g_sincos.emplace(dev, d).first;
    }
    *out = it->second;
    return LQB_OK;
}

static std::map<int, double2 *> g_logtab;
static int log_table_dev(const double2 **out)
{
    int dev = 0; LQB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto it = g_logtab.find(dev);
    if (it == g_logtab.end()) {
        std::vector<double> t = design::log_table();
        double2 *d = nullptr;
        LQB_CUDA(cudaMalloc((void **)&d, 128 * sizeof(double2)));
        LQB_CUDA(cudaMemcpy(d, t.data(), 128 * sizeof(double2), cudaMemcpyHostToDevice));
        it = g_logtab.emplace(dev, d).first;
    }
    *out = it->second;
    return LQB_OK;
}

static std::map<int, double *> g_atantab;
static int atan_table_dev(const double **out)
{
    int dev = 0; LQB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto it = g_atantab.find(dev);
    if (it == g_atantab.end()) {
        std::vector<double> t = design::atan_table();
        double *d = nullptr;
        LQB_CUDA(cudaMalloc((void **)&d, t.size() * sizeof(double)));
        LQB_CUDA(cudaMemcpy(d, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
        it = g_atantab.emplace(dev, d).first;
    }
    *out = it->second;
    return LQB_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr; cudaDriverEntryPointQueryResult q;
        if (getenv("LQB_NO_TMA")) return (EncodeTiledFn) nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// rows of n complex64 samples as a 2-D float tensor [rows][2n]; box = one warp's tile: 32 rows x 16 samples (128 B), 128B swizzle
// (pitch: samples between the starts of two rows when the call works on a time slice of longer rows; 0 = n)
static bool make_input_tmap(CUtensorMap *tm, const void *x, size_t n, size_t rows, unsigned box_rows = 32, unsigned box_floats = 32, size_t pitch = 0)
{
    EncodeTiledFn enc = encode_tiled();
    if (pitch == 0) pitch = n;
    if (!enc || (((size_t)x) & 15) || (n & 1) || (pitch & 1) || n * 2 > 0x7fffffffull || rows > 0x7fffffffull) return false;   // box coordinates are int32
    const cuuint64_t gdim[2] = { (cuuint64_t)n * 2, (cuuint64_t)rows }, gstride[1] = { (cuuint64_t)pitch * 8 };
    const cuuint32_t box[2] = { box_floats, box_rows }, estr[2] = { 1, 1 };
    // the 32-float box is the staged tile (128-byte swizzle); wider boxes are L2-prefetch shapes and carry no swizzle
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               box_floats == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ------------------------------------------------------------------------------------ stages
enum Kind { K_NCO = 0, K_IIR, K_RESAMP, K_AGC, K_AM, K_FM, K_DEEMPH, K_FIR, K_TF, K_BAM, K_DELAY, K_FMST };

}  // namespace lqb

struct lqb_chain_s;
struct lqb_stage_s;

namespace lqb {
// Stage handles that exist.  A chain borrows raw stage pointers; an object layer that destroys and recreates a stage
// (AmpModem's property setters, demod.hpp:250-276) leaves such a pointer dangling, so chains check their stages against
// this set before touching them instead of dereferencing freed memory.
static std::mutex g_live_mu;
static std::set<const lqb_stage_s *> g_live;
static thread_local bool t_no_tail_ahead = false;      // run_timepipe inside an SM partition: the tail's own side stream would leave it
// an SM partition of the device (green contexts): a stream for the front's SMs, one for the rest (sm_partition_create)
struct SmPartition {
    CUgreenCtx ga = nullptr, gb = nullptr; cudaStream_t front = nullptr, tail = nullptr, tail2 = nullptr; int sm_a = 0, sm_b = 0; bool ok = false;
};
// entry points that may run from a destructor / garbage collector put the caller's current device back
struct DevGuard {
    int prev = -1;
    DevGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace lqb

struct lqb_stage_s {
    lqb::Kind kind; int C = 1; int device = 0;
    lqb_chain_s *self_chain = nullptr;         // one-stage chain used by lqb_stage_execute*
    bool ready = false;                        // device state allocated (first execute / state access)
    lqb_stage_s(lqb::Kind k, int c) : kind(k), C(c)
    {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
        std::lock_guard<std::mutex> lk(lqb::g_live_mu); lqb::g_live.insert(this);
    }
    virtual ~lqb_stage_s();
    virtual int materialize() = 0;             // allocate and initialise the carried state in HBM
    virtual int clear() = 0;                   // liquid *_reset() on materialised state
    int ensure() { LQB_TRY(bind()); if (!ready) { LQB_TRY(materialize()); ready = true; } return LQB_OK; }
    virtual int reset() { host_reset(); if (!ready) return LQB_OK; LQB_TRY(bind()); LQB_CUDA(cudaDeviceSynchronize()); return clear(); }
    virtual void host_reset() {}
    virtual bool in_real() const { return false; }
    virtual bool out_real() const { return false; }
    virtual size_t out_len(size_t n) const { return n; }
    virtual void advance(size_t n) { (void)n; }    // host bookkeeping after n samples per channel
    int bind() const
    {
        cudaError_t e = cudaSetDevice(device);
        return e == cudaSuccess ? LQB_OK : lqb::fail(LQB_ECUDA, "cudaSetDevice(%d): %s -- this library has no CPU path", device, cudaGetErrorString(e));
    }
};

namespace lqb {

struct IirStage : lqb_stage_s {
    std::vector<float> B, A; int nsos = 0, mode = 0;
    bool real_io = false;                               // iirfilt_rrrf: real samples (RealIIRFilter and the R* classes)
    bool in_real() const override { return real_io; }
    bool out_real() const override { return real_io; }
    DevArr<float2> v;                                   // [nsos][2][C]
    IirStage(int c) : lqb_stage_s(K_IIR, c) {}
    int init(const std::vector<float> &b, const std::vector<float> &a)
    {
        nsos = (int)b.size() / 3; B.resize(3 * nsos); A.resize(3 * nsos);
        for (int s = 0; s < nsos; s++) {               // iirfiltsos_create normalises by a0
            const float a0 = a[3 * s];
            for (int k = 0; k < 3; k++) { B[3 * s + k] = b[3 * s + k] / a0; A[3 * s + k] = a[3 * s + k] / a0; }
        }
        return LQB_OK;
    }
    // blocked-scan evaluation (mode 2): block responses in double, scratch for the per-block states
    int scanB = 0; DevArr<double> scanH, scanM; DevArr<float2> vblk; DevArr<double2> sin;
    int materialize() override { return v.alloc((size_t)nsos * 2 * C); }
    int clear() override { return v.zero(); }
    // H[k][j]: output at step k of a block started from unit state j with zero input; M[i][j]: state i after B steps
    int prepare_scan(int B)
    {
        if (scanB == B) return LQB_OK;
        const int S = 2 * nsos;
        std::vector<double> H((size_t)B * S), M((size_t)S * S);
        for (int j = 0; j < S; j++) {
            std::vector<double> st(S, 0.0); st[j] = 1.0;
            for (int k = 0; k < B; k++) {
                double in = 0.0;
                for (int q = 0; q < nsos; q++) {
                    const double v1 = st[2 * q], v2 = st[2 * q + 1];
                    const double v0 = in - (double)A[3 * q + 1] * v1 - (double)A[3 * q + 2] * v2;
                    in = (double)B_(q, 0) * v0 + (double)B_(q, 1) * v1 + (double)B_(q, 2) * v2;
                    st[2 * q + 1] = v1; st[2 * q] = v0;
                }
                H[(size_t)k * S + j] = in;
            }
            for (int i = 0; i < S; i++) M[(size_t)i * S + j] = st[i];
        }
        LQB_TRY(scanH.alloc(H.size())); LQB_TRY(scanH.upload(H.data(), H.size()));
        LQB_TRY(scanM.alloc(M.size())); LQB_TRY(scanM.upload(M.data(), M.size()));
        scanB = B;
        return LQB_OK;
    }
    double B_(int q, int k) const { return this->B[3 * q + k]; }
    void fill(IirP &p, int s0, int ns) const
    {
        p.nsos = ns;
        for (int s = 0; s < ns; s++) for (int k = 0; k < 3; k++) { p.b[s][k] = B[3 * (s0 + s) + k]; p.a[s][k] = A[3 * (s0 + s) + k]; }
        p.v = v.p + (size_t)s0 * 2 * C;
    }
};

struct DeemphStage : lqb_stage_s {
    float b0 = 0, a1 = 0; DevArr<float> v1;
    DeemphStage(int c) : lqb_stage_s(K_DEEMPH, c) {}
    int materialize() override { return v1.alloc(C); }
    int clear() override { return v1.zero(); }
    bool in_real() const override { return true; }
    bool out_real() const override { return true; }
    void fill(DeP &p) const { p.b0 = b0; p.a1 = a1; p.v1 = v1.p; }
};

// iirfilt_{crcf,rrrf}_create(b, nb, a, na): transfer-function form, coefficients normalised by a[0] in float
struct TfStage : lqb_stage_s {
    std::vector<float> b, a; bool real_io = false; DevArr<float2> v;
    TfStage(int c) : lqb_stage_s(K_TF, c) {}
    bool in_real() const override { return real_io; }
    bool out_real() const override { return real_io; }
    int delay() const { return (int)std::max(b.size(), a.size()) - 1; }
    // delay elements the kernel instantiation keeps in registers
    int padded() const { const int d = std::max(1, delay()); return d <= 1 ? 1 : d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : 15; }
    int materialize() override { return v.alloc((size_t)(kMaxTf - 1) * C); }
    int clear() override { return v.zero(); }
    void fill(TfP &p) const
    {
        for (int i = 0; i < kMaxTf; i++) { p.b[i] = i < (int)b.size() ? b[i] : 0.f; p.na[i] = i < (int)a.size() ? -a[i] : 0.f; }
        p.nb = (int)b.size(); p.nna = (int)a.size(); p.v = v.p;
    }
};

struct FirStage : lqb_stage_s {
    std::vector<float> h; float scale = 1.f; DevArr<float> taps; DevArr<float2> hist[2]; int cur = 0;
    bool real_io = false;                               // firfilt_rrrf
    // firhilbf users: the imaginary lane has its own taps and the epilogue combines the lanes (fir.cu)
    std::vector<float> hlane_q, hq; DevArr<float> taps_q; int mode = FIR_PLAIN, delay = 0; bool in_r = false, out_r = false;
    float post_div = 0.f;                                                   // ampmodem SSB: (0.5 * side-band) / mod_index
    unsigned long long count = 0; std::vector<unsigned long long> ends;     // samples since reset; ends of recent calls (R2C)
    bool in_real() const override { return real_io || in_r; }
    bool out_real() const override { return real_io || out_r; }
    FirStage(int c) : lqb_stage_s(K_FIR, c) {}
    void host_reset() override { count = 0; ends.clear(); if (mode == FIR_R2C) ends.push_back(0); }   // the empty window behaves like a call boundary at 0
    int materialize() override
    {
        LQB_TRY(taps.alloc(h.size())); LQB_TRY(taps.upload(h.data(), h.size()));
        if (!hlane_q.empty()) { LQB_TRY(taps_q.alloc(hlane_q.size())); LQB_TRY(taps_q.upload(hlane_q.data(), hlane_q.size())); }
        for (int k = 0; k < 2; k++) LQB_TRY(hist[k].alloc((size_t)std::max<size_t>(1, h.size() - 1) * C));
        return LQB_OK;
    }
    int clear() override { LQB_TRY(hist[0].zero()); return hist[1].zero(); }
    void advance(size_t n) override
    {
        if (!n) return;
        cur ^= 1; count += n;
        if (mode == FIR_R2C) {              // the reference's pairwise read runs one sample past each call (taken as zero)
            ends.push_back(count);
            while (!ends.empty() && ends.front() + (unsigned long long)delay - 1 < count) ends.erase(ends.begin());
        }
    }
};

// wdelayf / wdelaycf as the reference's Delay uses them (read, then push): y[k] = x[k - (nd + 1)]
struct DelayStage : lqb_stage_s {
    long long D = 1; bool real_io = false; int cur = 0; DevArr<float2> hist[2];      // float2 slots also when real
    DelayStage(int c) : lqb_stage_s(K_DELAY, c) {}
    bool in_real() const override { return real_io; }
    bool out_real() const override { return real_io; }
    int materialize() override { for (int k = 0; k < 2; k++) LQB_TRY(hist[k].alloc((size_t)D * C)); return LQB_OK; }
    int clear() override { LQB_TRY(hist[0].zero()); return hist[1].zero(); }
    void advance(size_t n) override { if (n) cur ^= 1; }
};

struct ResampStage : lqb_stage_s {
    int variant = 0;                                    // 0 cccf (ComplexResampler), 1 crcf (CResampler), 2 rrrf (RResampler, RealResampler)
    bool in_real() const override { return variant == 2; }
    bool out_real() const override { return variant == 2; }
    float rate = 1.f; design::ResampDesign d; uint32_t step = 0, phase = 0, count = 0;
    DevArr<float> bank; DevArr<float2> ring;            // ring [sublen][C]
    // lane-split front kernels: the call's tap stream (lanes.cu), one scratch buffer per stream that runs the stage
    std::map<cudaStream_t, DevArr<char>> tapbuf;
    ResampStage(int c) : lqb_stage_s(K_RESAMP, c) {}
    int materialize() override
    {
        LQB_TRY(bank.alloc(d.bank.size())); LQB_TRY(bank.upload(d.bank.data(), d.bank.size()));
        return ring.alloc((size_t)d.sublen * C);
    }
    void host_reset() override { phase = 0; count = 0; }
    int clear() override { return ring.zero(); }
    // the fused sequential path needs output windows that do not overlap and at most one output per
    // staged tile, with the 32-bit phase arithmetic of the tap stream free of wrap-around
    bool decimating() const
    {
        const uint64_t lo = (uint64_t)std::max<unsigned>(d.sublen, kSeqTS) << 24;
        return variant == 0 && (uint64_t)step >= lo && (uint64_t)step < (1ull << 32) - (1ull << 28) && d.sublen <= (unsigned)kMaxResampSub;
    }
    size_t out_len(size_t n) const override
    {
        const uint64_t lim = ((uint64_t)n << 24);           // outputs while phase + k*step <= n*2^24 - 1
        if (n == 0 || (uint64_t)phase > lim - 1) return 0;
        return (size_t)((lim - 1 - phase) / step + 1);
    }
    void advance(size_t n) override
    {
        const uint64_t k = out_len(n);
        phase = (uint32_t)((uint64_t)phase + k * step - ((uint64_t)n << 24));
        count = (uint32_t)((count + n) % d.sublen);
    }
    void fill(ResampP &p) const
    {
        p.step = step; p.phase = phase; p.bits = (int)d.bits; p.sublen = (int)d.sublen; p.npfb = (int)d.npfb;
        p.bank = bank.p; p.ring = ring.p; p.count = count; p.variant = variant;
    }
};

struct NcoStage : lqb_stage_s {
    int type = 0, dir = LQB_MIX_UP; float alpha = 0.1f, beta = 0.f;
    DevArr<uint32_t> theta, dtheta;
    NcoStage(int c) : lqb_stage_s(K_NCO, c) { beta = std::sqrt(alpha); }
    int materialize() override { LQB_TRY(theta.alloc(C)); return dtheta.alloc(C); }
    int clear() override { LQB_TRY(theta.zero()); return dtheta.zero(); }
    int fill(NcoP &p) const
    {
        p.theta = theta.p; p.dtheta = dtheta.p; p.type = type; p.dir = dir;
        return sincos_table(&p.sincos);
    }
    int rmw(DevArr<uint32_t> &arr, uint32_t add, bool set)
    {
        LQB_TRY(ensure());
        LQB_CUDA(cudaDeviceSynchronize());
        if (set) return arr.fill(add);
        std::vector<uint32_t> h(C);
        LQB_TRY(arr.download(h.data(), C));
        for (auto &x : h) x += add;
        return arr.upload(h.data(), C);
    }
};

struct AgcStage : lqb_stage_s {
    float alpha = 1e-2f, scale = 1.f, threshold = 0.f; int locked = 0; bool squelch = false; unsigned timeout = 100;
    // Which gain loop runs (devmath.cuh): the general one evaluates the oracle's operations bit for bit (double-precision
    // smoothing, correctly rounded log / exp); the single-precision one is 1e-7 from it and 3-4x faster.  A carrier PLL
    // behind the AGC amplifies a last-bit difference to 1e-4..1e-3, so AUTO takes the single-precision loop
    // only where the planner sees a feed-forward discriminator (FreqDem) and no PLL demodulator downstream.
    int precision = LQB_AGC_AUTO; bool plan_fast = false;
    DevArr<float> g, y2p; DevArr<int> mode; DevArr<unsigned> timer, rise;
    // few channels: the gain loop runs on this stream, a few hundred samples ahead of the demodulator (run_segment)
    cudaStream_t ahead = nullptr; cudaEvent_t ev_begin = nullptr, ev_chunk[8] = {};
    AgcStage(int c) : lqb_stage_s(K_AGC, c) {}
    ~AgcStage() override
    {
        if (ahead) { cudaStreamSynchronize(ahead); cudaStreamDestroy(ahead); }
        if (ev_begin) cudaEventDestroy(ev_begin);
        for (auto &e : ev_chunk) if (e) cudaEventDestroy(e);
    }
    int pipeline_resources()
    {
        if (ahead) return LQB_OK;
        LQB_CUDA(cudaStreamCreateWithFlags(&ahead, cudaStreamNonBlocking));
        LQB_CUDA(cudaEventCreateWithFlags(&ev_begin, cudaEventDisableTiming));
        for (auto &e : ev_chunk) LQB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return LQB_OK;
    }
    int materialize() override
    {
        LQB_TRY(g.alloc(C)); LQB_TRY(y2p.alloc(C)); LQB_TRY(mode.alloc(C)); LQB_TRY(timer.alloc(C)); LQB_TRY(rise.alloc(1));
        LQB_TRY(timer.fill(timeout));
        return clear();
    }
    void host_reset() override { locked = 0; }
    int clear() override
    {
        LQB_TRY(g.fill(1.0f)); LQB_TRY(y2p.fill(1.0f));
        return mode.fill(squelch ? 1 : 7);
    }
    int fill(AgcP &p) const
    {
        LQB_TRY(log_table_dev(&p.logtab));
        p.alpha = alpha; p.scale = scale; p.threshold = threshold; p.one_minus_alpha = 1.0 - (double)alpha;
        p.chi = (float)p.one_minus_alpha; p.clo = (float)(p.one_minus_alpha - (double)p.chi);
        p.chalf = -0.5f * alpha; p.cl2 = (float)((double)p.chalf * 0.6931471805599453094);
        p.big = alpha > 0.0112f ? 1 : 0;
        p.fast = (!locked && !squelch && (precision == LQB_AGC_FAST || (precision == LQB_AGC_AUTO && plan_fast))) ? 1 : 0;
        p.locked = locked; p.timeout = timeout; p.g = g.p; p.y2p = y2p.p; p.mode = mode.p; p.timer = timer.p;
        p.rise_count = rise.p;
        return LQB_OK;
    }
};

struct AmStage : lqb_stage_s {
    float mod = 0.75f; int type = 0, suppressed = 1; std::vector<float> lp, dc; uint32_t count = 0;
    DevArr<float2> lp_ring; DevArr<float> dc_ring; DevArr<uint32_t> theta, dtheta;
    // USB / LSB: the Hilbert pair (and, with carrier, the DC blocker behind it) are feed-forward filters run by the FIR
    // kernel; they live here as sub-stages the planner expands into their own segments
    FirStage *hil = nullptr, *dcb = nullptr;
    AmStage(int c) : lqb_stage_s(K_AM, c) {}
    ~AmStage() override { delete hil; delete dcb; }
    bool ssb() const { return type != LQB_AMPMODEM_DSB; }
    int materialize() override
    {
        if (hil) LQB_TRY(hil->ensure());
        if (dcb) LQB_TRY(dcb->ensure());
        LQB_TRY(lp_ring.alloc((size_t)kAmRing * C)); LQB_TRY(dc_ring.alloc((size_t)kAmRing * C));
        LQB_TRY(theta.alloc(C)); return dtheta.alloc(C);
    }
    void host_reset() override { count = 0; if (hil) hil->host_reset(); if (dcb) dcb->host_reset(); }
    int clear() override
    {
        if (hil && hil->ready) LQB_TRY(hil->clear());
        if (dcb && dcb->ready) LQB_TRY(dcb->clear());
        LQB_TRY(lp_ring.zero()); LQB_TRY(dc_ring.zero()); LQB_TRY(theta.zero()); return dtheta.zero();
    }
    bool out_real() const override { return true; }
    void advance(size_t n) override { count = (uint32_t)((count + n) % kAmRing); if (hil) hil->advance(n); if (dcb) dcb->advance(n); }
    int fill(AmP &p) const
    {
        p.out_v1 = ssb() && !suppressed ? 1 : 0;
        p.mod_index = mod; p.pll_alpha = 0.001f; p.pll_beta = std::sqrt(0.001f); p.suppressed = suppressed;
        for (int i = 0; i < kAmTaps; i++) { p.lp[i] = lp[kAmTaps - 1 - i]; p.dc[i] = dc[kAmTaps - 1 - i]; }
        p.lp_ring = lp_ring.p; p.dc_ring = dc_ring.p; p.theta = theta.p; p.dtheta = dtheta.p; p.count = count;
        return sincos_table(&p.sincos);
    }
};

// BroadcastAM (demod.hpp:94-153)
struct BamStage : lqb_stage_s {
    int m = 25, ntaps_pad = 56; std::vector<float> lp, B, A;       // lowpass taps (design order), DC-block sections
    DevArr<float> hrev, dcv; DevArr<float2> hist; DevArr<uint32_t> theta, dtheta;
    BamStage(int c) : lqb_stage_s(K_BAM, c) {}
    bool out_real() const override { return true; }
    int materialize() override
    {
        std::vector<float> hr((size_t)ntaps_pad, 0.f);                // window order, zero taps on the oldest side
        const int nt = (int)lp.size(), pad = ntaps_pad - nt;
        for (int i = 0; i < nt; i++) hr[pad + i] = lp[nt - 1 - i];
        LQB_TRY(hrev.alloc(hr.size())); LQB_TRY(hrev.upload(hr.data(), hr.size()));
        LQB_TRY(hist.alloc((size_t)(ntaps_pad - 1) * C)); LQB_TRY(dcv.alloc((size_t)4 * C));
        LQB_TRY(theta.alloc(C)); return dtheta.alloc(C);
    }
    int clear() override { LQB_TRY(hist.zero()); LQB_TRY(dcv.zero()); LQB_TRY(theta.zero()); return dtheta.zero(); }
    int fill(BamP &p) const
    {
        p.m = m; p.ntaps_pad = ntaps_pad; p.hrev = hrev.p; p.pll_alpha = 0.001f; p.pll_beta = std::sqrt(0.001f);
        for (int s = 0; s < 2; s++) for (int k = 0; k < 3; k++) { p.b[s][k] = B[3 * s + k]; p.a[s][k] = A[3 * s + k]; }
        p.hist = hist.p; p.dcv = dcv.p; p.theta = theta.p; p.dtheta = dtheta.p;
        LQB_TRY(atan_table_dev(&p.atantab));
        return sincos_table(&p.sincos);
    }
};

// FMStereo (demod.hpp:4-85): complex in, interleaved (left, right) float pairs out -- two floats per resampler output
struct FmstStage : lqb_stage_s {
    float b0 = 0, a1 = 0; design::ResampDesign d; uint32_t step = 0, phase = 0, count = 0;
    DevArr<float> bank, pe, vL, vR, ringL, ringR; DevArr<float2> rprime; DevArr<uint32_t> theta, dtheta;
    FmstStage(int c) : lqb_stage_s(K_FMST, c) {}
    bool out_real() const override { return true; }
    int materialize() override
    {
        LQB_TRY(bank.alloc(d.bank.size())); LQB_TRY(bank.upload(d.bank.data(), d.bank.size()));
        LQB_TRY(pe.alloc(C)); LQB_TRY(vL.alloc(C)); LQB_TRY(vR.alloc(C)); LQB_TRY(rprime.alloc(C)); LQB_TRY(theta.alloc(C)); LQB_TRY(dtheta.alloc(C));
        LQB_TRY(ringL.alloc((size_t)d.sublen * C)); return ringR.alloc((size_t)d.sublen * C);
    }
    // FMStereo::reset (demod.hpp:35-38) resets the two resamplers and nothing else
    void host_reset() override { phase = 0; count = 0; }
    int clear() override { LQB_TRY(ringL.zero()); return ringR.zero(); }
    // resampler outputs over n input samples: all of them (rate <= 1: at most one per input, every one a kept pair), and for
    // pcm_rate > iq_rate the inputs that yield exactly one -- the only pairs FMStereo::execute keeps (demod.hpp:45-48)
    void walk(size_t n, uint64_t *total, uint64_t *kept) const
    {
        const uint64_t lim = ((uint64_t)n << 24);
        *total = (n == 0 || (uint64_t)phase > lim - 1) ? 0 : (lim - 1 - phase) / step + 1;
        *kept = *total;
        if (step < (1u << 24)) {
            uint64_t k = 0; uint32_t ph = phase;
            for (size_t i = 0; i < n; i++) {
                unsigned nl = 0;
                while (ph <= 0x00ffffffu) { nl++; ph += step; }
                ph -= (1u << 24);
                if (nl == 1) k++;
            }
            *kept = k;
        }
    }
    size_t out_len(size_t n) const override { uint64_t t, k; walk(n, &t, &k); return (size_t)(2 * k); }
    void advance(size_t n) override
    {
        uint64_t t, k; walk(n, &t, &k);
        phase = (uint32_t)((uint64_t)phase + t * step - ((uint64_t)n << 24));
        count = (uint32_t)((count + n) % d.sublen);
    }
    int fill(FmstP &p) const
    {
        p.ref = (float)(1.0f / (2 * design::kPi * 4.0f)); p.pll_alpha = 0.1f; p.pll_beta = std::sqrt(0.1f); p.b0 = b0; p.a1 = a1;
        p.rs.step = step; p.rs.phase = phase; p.rs.bits = (int)d.bits; p.rs.sublen = (int)d.sublen; p.rs.npfb = (int)d.npfb;
        p.rs.bank = bank.p; p.rs.ring = nullptr; p.rs.count = count; p.rs.variant = 2;
        p.rprime = rprime.p; p.theta = theta.p; p.dtheta = dtheta.p; p.pe = pe.p; p.vL = vL.p; p.vR = vR.p; p.ringL = ringL.p; p.ringR = ringR.p;
        LQB_TRY(atan_table_dev(&p.atantab));
        return sincos_table(&p.sincos);
    }
};

struct FmStage : lqb_stage_s {
    float kf = 0.1f, ref = 0.f; DevArr<float2> rprime;
    FmStage(int c) : lqb_stage_s(K_FM, c) {}
    int materialize() override { return rprime.alloc(C); }
    int clear() override { return rprime.zero(); }
    bool out_real() const override { return true; }
    void fill(FmP &p) const { p.ref = ref; p.rprime = rprime.p; }
};

// channel count below which closed-form stages run time-parallel instead of one thread per channel
constexpr int kParChannels = 16384;

// ------------------------------------------------------------------------------------ chain
struct Segment {
    enum Type { SEQ, FIR, RESAMP_PAR, AMTAIL, NCO_PAR, BAM, DELAY, FMST } type = SEQ;
    unsigned mask = 0; int nsos = 0, sos0 = 0;
    int out_real = -1;                         // element type of this segment's output when it is not its last stage's (0 / 1)
    std::vector<lqb_stage_s *> st;
    std::string name;
};

}  // namespace lqb

struct lqb_chain_s {
    std::vector<lqb_stage_s *> stages;
    int fuse = 1;                              // 0 one kernel per stage, 1 front/tail split, 2 longest runs
    int last_launches = 0;
    std::string plan;
    std::string last_kernels;                  // kernels the last execute launched, in order (lqb_chain_last_kernels)
    // a call that failed after its first launch leaves device state ahead of the host bookkeeping: refuse further calls
    bool poisoned = false; std::string poison_why;
    // host-execute resources: per-stream input/output staging and ping-pong scratch
    static constexpr int kStreams = 3;
    cudaStream_t streams[kStreams] = { nullptr, nullptr, nullptr };
    lqb::DevArr<char> h_in[kStreams], h_out[kStreams], h_tmp[kStreams][2], h_cvt[kStreams];
    cudaEvent_t ev_ready[2] = { nullptr, nullptr }, ev_free[2] = { nullptr, nullptr };   // time-sliced host pipeline
    // pageable callers (a plain numpy array): pinned bounce buffers the host threads fill while the previous one crosses PCIe
    static constexpr int kPin = 3;
    void *pin_buf[kPin] = { nullptr, nullptr, nullptr }; size_t pin_bytes = 0;
    cudaEvent_t ev_pin[kPin] = { nullptr, nullptr, nullptr }; bool pin_used[kPin] = { false, false, false }; unsigned pin_turn = 0;
    // device-execute scratch
    lqb::DevArr<char> d_tmp[2], d_cvt;
    // overlapped device calls (lqb_chain_set_overlap): the decimated-rate tail of call k runs on tail_stream while the
    // caller's stream already runs the full-rate front of call k+1; the hand-off buffer alternates between d_tmp[0] / [1]
    bool overlap = false;
    cudaStream_t tail_stream = nullptr;
    cudaEvent_t ev_front[2] = { nullptr, nullptr }, ev_tail[2] = { nullptr, nullptr };
    bool tail_pending[2] = { false, false };
    unsigned ov_calls = 0;
    // few channels (run_timepipe): the decimated-rate tail of time slice j runs on pipe_stream while the front works on slice j + 1
    cudaStream_t pipe_stream = nullptr, front_stream = nullptr;
    std::vector<cudaEvent_t> ev_pipe;
    // optional per-segment timing of execute_dev: one event pair per segment per call, on the stream the segment ran on
    bool timing = false;
    std::vector<std::vector<std::pair<cudaEvent_t, cudaEvent_t>>> timed_calls;
    void clear_timing()
    {
        for (auto &call : timed_calls) for (auto &p : call) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        timed_calls.clear();
    }
    ~lqb_chain_s()
    {
        if (tail_stream) { cudaStreamSynchronize(tail_stream); cudaStreamDestroy(tail_stream); }
        if (pipe_stream) { cudaStreamSynchronize(pipe_stream); cudaStreamDestroy(pipe_stream); }
        if (front_stream) { cudaStreamSynchronize(front_stream); cudaStreamDestroy(front_stream); }
        for (auto &e : ev_pipe) cudaEventDestroy(e);
        for (int b = 0; b < 2; b++) { if (ev_front[b]) cudaEventDestroy(ev_front[b]); if (ev_tail[b]) cudaEventDestroy(ev_tail[b]); }
        clear_timing();
        for (auto &s : streams) if (s) cudaStreamDestroy(s);
        for (int b = 0; b < 2; b++) { if (ev_ready[b]) cudaEventDestroy(ev_ready[b]); if (ev_free[b]) cudaEventDestroy(ev_free[b]); }
        for (int b = 0; b < kPin; b++) { if (pin_buf[b]) cudaFreeHost(pin_buf[b]); if (ev_pin[b]) cudaEventDestroy(ev_pin[b]); }
    }
};

lqb_stage_s::~lqb_stage_s()
{
    { std::lock_guard<std::mutex> lk(lqb::g_live_mu); lqb::g_live.erase(this); }
    delete self_chain;
}

namespace lqb {

static const char *kind_name(Kind k)
{
    switch (k) {
    case K_NCO: return "nco"; case K_IIR: return "iir"; case K_RESAMP: return "resamp"; case K_AGC: return "agc";
    case K_AM: return "ampmodem"; case K_FM: return "freqdem"; case K_DEEMPH: return "deemph"; case K_FIR: return "fir"; case K_TF: return "tf"; case K_BAM: return "broadcast_am"; case K_DELAY: return "delay"; case K_FMST: return "fmstereo";
    }
    return "?";
}
static unsigned kind_flag(Kind k)
{
    switch (k) {
    case K_NCO: return F_NCO; case K_IIR: return F_IIR; case K_RESAMP: return F_RS; case K_AGC: return F_AGC;
    case K_AM: return F_AM; case K_FM: return F_FM; case K_DEEMPH: return F_DE; case K_TF: return F_TF; default: return 0;
    }
}
// position in the order the sequential kernel applies its stages
static int kind_rank(Kind k)
{
    switch (k) { case K_NCO: return 0; case K_IIR: return 1; case K_RESAMP: return 2; case K_AGC: return 3;
                 case K_AM: case K_FM: return 4; case K_DEEMPH: return 5; default: return 99; }
}

static unsigned run_mask(const std::vector<lqb_stage_s *> &st, size_t i0, size_t len, int *nsos)
{
    unsigned m = 0; *nsos = 0;
    for (size_t i = i0; i < i0 + len; i++) {
        m |= kind_flag(st[i]->kind);
        if (st[i]->kind == K_IIR) *nsos = static_cast<IirStage *>(st[i])->nsos;
        if (st[i]->kind == K_TF) *nsos = static_cast<TfStage *>(st[i])->padded();
    }
    if (st[i0]->in_real()) m |= F_INREAL;
    return m;
}

static bool run_fusable(const std::vector<lqb_stage_s *> &st, size_t i0, size_t len, int level)
{
    int prev = -1; bool has_rs = false, has_am = false;
    for (size_t i = i0; i < i0 + len; i++) {
        const lqb_stage_s *s = st[i];
        const int r = kind_rank(s->kind);
        if (r <= prev || r == 99) return false;
        prev = r;
        if (s->kind == K_RESAMP && !static_cast<const ResampStage *>(s)->decimating()) return false;
        if (s->kind == K_IIR && static_cast<const IirStage *>(s)->mode == 2) return false;
        if (s->kind == K_AM && static_cast<const AmStage *>(s)->ssb()) return false;
        has_rs |= s->kind == K_RESAMP; has_am |= s->kind == K_AM;
    }
    // level 1 keeps the ampmodem's shared-memory windows out of the full-rate kernel: they would cut
    // its occupancy four-fold, while the decimated hand-off they avoid is 2.4 % of the traffic
    if (level < 2 && has_am) return false;    // below level 2 the ampmodem always runs in its own tail kernel
    (void)has_rs;
    int nsos; const unsigned m = run_mask(st, i0, len, &nsos);
    return seq_supported(m, nsos);
}

static int build_plan(lqb_chain_s *c, std::vector<Segment> &segs)
{
    const auto &st = c->stages;
    segs.clear();
    for (size_t i = 0; i < st.size();) {
        Segment g;
        if (st[i]->kind == K_FIR) { g.type = Segment::FIR; g.st = { st[i] }; g.name = "fir"; segs.push_back(g); i++; continue; }
        if (st[i]->kind == K_FMST) { g.type = Segment::FMST; g.st = { st[i] }; g.name = "fmstereo"; segs.push_back(g); i++; continue; }
        if (st[i]->kind == K_DELAY) { g.type = Segment::DELAY; g.st = { st[i] }; g.name = "delay"; segs.push_back(g); i++; continue; }
        if (st[i]->kind == K_RESAMP && static_cast<ResampStage *>(st[i])->variant != 0) {      // real-tap resamplers: time-parallel kernel
            g.type = Segment::RESAMP_PAR; g.st = { st[i] }; g.name = "par[resamp]"; segs.push_back(g); i++; continue;
        }
        // [AGC ->] BroadcastAM [-> de-emphasis] (bam.cu).  The gain loop joins only behind a decimating kernel: it then
        // runs in place on the time-major hand-off buffer
        {
            size_t j = i; std::string nm; Segment b;
            const bool after_rs = !segs.empty() && segs.back().type == Segment::SEQ && (segs.back().mask & F_RS);
            if (c->fuse >= 1 && after_rs && st[j]->kind == K_AGC && j + 1 < st.size() && st[j + 1]->kind == K_BAM) { b.st.push_back(st[j++]); nm = "agc+"; }
            if (st[j]->kind == K_BAM) {
                b.type = Segment::BAM; b.st.push_back(st[j++]); nm += "broadcast_am";
                if (c->fuse >= 1 && j < st.size() && st[j]->kind == K_DEEMPH) { b.st.push_back(st[j++]); nm += "+deemph"; }
                b.name = "bam[" + nm + "]"; segs.push_back(b); i = j; continue;
            }
        }
        // ampmodem USB / LSB: [carrier loop ->] Hilbert pair [-> DC blocker], the last two on the FIR kernel
        if (st[i]->kind == K_AM && static_cast<AmStage *>(st[i])->ssb()) {
            AmStage *am = static_cast<AmStage *>(st[i]);
            if (!am->suppressed) {
                Segment p; p.type = Segment::AMTAIL; p.st = { am }; p.name = "am[carrier-loop]"; p.out_real = 0; segs.push_back(p);
            }
            Segment h; h.type = Segment::FIR; h.st = { am->hil }; h.name = "fir[hilbert]"; segs.push_back(h);
            if (!am->suppressed) { Segment d; d.type = Segment::FIR; d.st = { am->dcb }; d.name = "fir[dcblock]"; segs.push_back(d); }
            i++; continue;
        }
        // [AGC ->] ampmodem [-> de-emphasis] : the decimated-rate tail kernel (am.cu)
        if (c->fuse < 2) {
            size_t j = i; std::string nm;
            if (c->fuse >= 1 && st[j]->kind == K_AGC && j + 1 < st.size() && st[j + 1]->kind == K_AM && !static_cast<AmStage *>(st[j + 1])->ssb()) { g.st.push_back(st[j++]); nm = "agc+"; }
            if (st[j]->kind == K_AM) {
                g.type = Segment::AMTAIL; g.st.push_back(st[j++]); nm += "ampmodem";
                if (c->fuse >= 1 && j < st.size() && st[j]->kind == K_DEEMPH) { g.st.push_back(st[j++]); nm += "+deemph"; }
                g.name = "am[" + nm + "]"; segs.push_back(g); i = j; continue;
            }
            g.st.clear();
        }
        // Few channels cannot fill the GPU one thread per channel.  The oscillator (closed-form phase) and the
        // resampler (closed-form output positions) have no loop-carried dependence, so below kParChannels they run
        // time-parallel, the mixer applied while the resampler stages its input.
        if (st[i]->C < kParChannels || (st[i]->kind == K_RESAMP && !static_cast<ResampStage *>(st[i])->decimating()) ||
            (st[i]->kind == K_NCO && i + 1 < st.size() && st[i + 1]->kind == K_RESAMP && !static_cast<ResampStage *>(st[i + 1])->decimating())) {
            if (st[i]->kind == K_NCO && c->fuse >= 1 && i + 1 < st.size() && st[i + 1]->kind == K_RESAMP) {
                g.type = Segment::RESAMP_PAR; g.st = { st[i], st[i + 1] }; g.name = "par[nco+resamp]"; segs.push_back(g); i += 2; continue;
            }
            if (st[i]->kind == K_RESAMP) { g.type = Segment::RESAMP_PAR; g.st = { st[i] }; g.name = "par[resamp]"; segs.push_back(g); i++; continue; }
            if (st[i]->kind == K_NCO && st[i]->C < kParChannels) { g.type = Segment::NCO_PAR; g.st = { st[i] }; g.name = "par[nco]"; segs.push_back(g); i++; continue; }
        }
        size_t best = 1;
        if (c->fuse) for (size_t len = std::min<size_t>(6, st.size() - i); len >= 2; len--) if (run_fusable(st, i, len, c->fuse)) { best = len; break; }
        if (best == 1 && st[i]->kind == K_AM) {      // (fusion level 2 without the one fully fused pattern: the tail kernel after all)
            Segment t; t.type = Segment::AMTAIL; t.st = { st[i] }; std::string nm = "ampmodem";
            size_t j = i + 1;
            if (j < st.size() && st[j]->kind == K_DEEMPH) { t.st.push_back(st[j++]); nm += "+deemph"; }
            t.name = "am[" + nm + "]"; segs.push_back(t); i = j; continue;
        }
        if (best == 1) {
            if (st[i]->kind == K_IIR) {              // long cascades: kMaxSos sections per launch
                IirStage *q = static_cast<IirStage *>(st[i]);
                for (int s0 = 0; s0 < q->nsos; s0 += kMaxSos) {
                    Segment h; h.type = Segment::SEQ; h.mask = F_IIR | (q->real_io ? F_INREAL : 0u); h.sos0 = s0; h.nsos = std::min(kMaxSos, q->nsos - s0);
                    h.st = { st[i] }; h.name = std::string(q->mode == 2 && q->nsos <= kMaxSos ? "scan" : "seq") + "[iir" + std::to_string(h.nsos) + "]"; segs.push_back(h);
                }
                i++; continue;
            }
        }
        g.type = Segment::SEQ; g.mask = run_mask(st, i, best, &g.nsos);
        if (!seq_supported(g.mask, g.nsos)) return fail(LQB_EINVAL, "no kernel for stage %s", kind_name(st[i]->kind));
        g.name = "seq[";
        for (size_t k = i; k < i + best; k++) {
            g.st.push_back(st[k]);
            g.name += (k > i ? "+" : "") + std::string(kind_name(st[k]->kind));
            if (st[k]->kind == K_IIR) g.name += std::to_string(g.nsos);
        }
        g.name += "]";
        segs.push_back(g); i += best;
    }
    c->plan.clear();
    for (size_t k = 0; k < segs.size(); k++) c->plan += (k ? " -> " : "") + segs[k].name;
    // gain-loop precision (AgcStage::precision): single precision where a discriminator follows and no carrier PLL does
    for (size_t i = 0; i < st.size(); i++) {
        if (st[i]->kind != K_AGC) continue;
        bool fm = false, pll = false;
        for (size_t k = i + 1; k < st.size(); k++) {
            fm  = fm  || st[k]->kind == K_FM;
            pll = pll || st[k]->kind == K_AM || st[k]->kind == K_BAM || st[k]->kind == K_FMST;
        }
        static_cast<AgcStage *>(st[i])->plan_fast = fm && !pll;
    }
    return LQB_OK;
}

static size_t seg_out_len(const Segment &g, size_t n) { for (auto *s : g.st) n = s->out_len(n); return n; }

// run one segment on channels [ch0, ch0 + nch) of the chain; x/y point at the first of those rows
static void note_kernel(std::string *kn, const std::string &name) { if (kn) { if (!kn->empty()) *kn += ";"; *kn += name; } }

// a time slice of a longer call (run_timepipe): input rows in_pitch samples apart, output rows out_pitch samples apart
struct SegIO { size_t in_pitch = 0, out_pitch = 0; int tail_part = 0; };      // tail_part: 0 the whole AM tail, 1 its gain loop only, 2 its demodulator only

static int run_segment(const Segment &g, const void *x, void *y, size_t n, size_t n_out, int ch0, int nch, cudaStream_t stream,
                       bool in_tmajor, bool out_tmajor, int *launches, unsigned extra_mask = 0, int front_ring = 3, std::string *kn = nullptr,
                       SegIO io = SegIO())
{
    (*launches)++;
    const lqb_stage_s *first = g.st.front();
    if (g.type == Segment::AMTAIL) {
        AmTailArgs a{};
        bool has_agc = false, has_de = false;
        a.x = (const float2 *)x; a.y = (float *)y; a.C = nch; a.ch0 = ch0; a.Ctot = first->C; a.in_tmajor = in_tmajor ? 1 : 0;
        a.n = (long long)n; a.in_pitch = in_tmajor ? (long long)nch : (long long)n; a.out_pitch = (long long)(io.out_pitch ? io.out_pitch : n_out);
        if (io.in_pitch && !in_tmajor) return fail(LQB_EINVAL, "internal: a time slice reaches the AM tail through the time-major hand-off only");
        for (lqb_stage_s *s : g.st) {
            if (s->kind == K_AGC) { LQB_TRY(static_cast<AgcStage *>(s)->fill(a.agc)); has_agc = true; }
            else if (s->kind == K_AM) LQB_TRY(static_cast<AmStage *>(s)->fill(a.am));
            else if (s->kind == K_DEEMPH) { static_cast<DeemphStage *>(s)->fill(a.de); has_de = true; }
        }
        // Few channels: gain loop and demodulator are each ONE dependent chain per channel with a lone warp per scheduler --
        // latency, not throughput, and the machine is mostly idle.  So the block is cut along time: the gain loop runs
        // chunk j + 1 on its own stream while the demodulator works on chunk j (state carries between chunks exactly as
        // between calls), and the tail takes about max(gain loop, demodulator) instead of their sum.
        if (io.tail_part && has_agc && in_tmajor && !a.am.suppressed) {
            // run_timepipe inside an SM partition: gain loop and demodulator of a slice on two streams of the tail's SMs
            if (io.tail_part == 1) { LQB_CUDA(agc_tmajor_launch(a, stream)); note_kernel(kn, "agc_tmajor_kernel"); }
            else { LQB_CUDA(amtail_launch(false, has_de, a, stream)); note_kernel(kn, amtail_kernel_name(false, a)); }
            return LQB_OK;
        }
        // (eight chunks up to 16384 channels; two up to 40000, where the second half of the gain loop hides behind the first
        // half of the demodulator: 32768 channels 1.06 -> 1.01 ms; at 65536 both fill the machine and nothing is gained, measured)
        int pipe_max = 40000;
        if (const char *e = getenv("LQB_TAILPIPE_MAX")) pipe_max = atoi(e);                // tuning override
        if (has_agc && in_tmajor && nch <= pipe_max && n >= 512 && !a.am.suppressed && !t_no_tail_ahead && !getenv("LQB_NO_TAILPIPE")) {
            AgcStage *ag = nullptr;
            for (lqb_stage_s *s : g.st) if (s->kind == K_AGC) ag = static_cast<AgcStage *>(s);
            LQB_TRY(ag->pipeline_resources());
            int K = nch <= 16384 ? 8 : 2;
            if (const char *e = getenv("LQB_TAILPIPE_CHUNKS")) { const int v = atoi(e); if (v >= 2 && v <= 8) K = v; }     // tuning override
            const long long step = (((long long)n + K - 1) / K + 7) / 8 * 8;      // whole groups of the demodulator's eight samples
            LQB_CUDA(cudaEventRecord(ag->ev_begin, stream));
            LQB_CUDA(cudaStreamWaitEvent(ag->ahead, ag->ev_begin, 0));
            int j = 0;
            for (long long k0 = 0; k0 < (long long)n; k0 += step, j++) {
                AmTailArgs c = a;
                c.n = std::min<long long>(step, (long long)n - k0);
                c.x = a.x + k0 * a.in_pitch; c.y = a.y + k0;
                c.am.count = (uint32_t)((a.am.count + (unsigned long long)k0) % kAmRing);
                LQB_CUDA(agc_tmajor_launch(c, ag->ahead));
                LQB_CUDA(cudaEventRecord(ag->ev_chunk[j], ag->ahead));
                LQB_CUDA(cudaStreamWaitEvent(stream, ag->ev_chunk[j], 0));
                LQB_CUDA(amtail_launch(false, has_de, c, stream));
                (*launches) += 2;
            }
            (*launches)--;
            note_kernel(kn, "agc_tmajor_kernel(x" + std::to_string(j) + ", ahead)");
            note_kernel(kn, std::string(amtail_kernel_name(false, a)) + "(x" + std::to_string(j) + ")");
            return LQB_OK;
        }
        LQB_CUDA(amtail_launch(has_agc, has_de, a, stream));
        if (amtail_launch_count(has_agc, a) > 1) note_kernel(kn, "agc_tmajor_kernel");
        note_kernel(kn, amtail_kernel_name(has_agc, a));
        *launches += amtail_launch_count(has_agc, a) - 1;
        return LQB_OK;
    }
    if (g.type == Segment::FMST) {
        FmstArgs a{};
        a.x = (const float2 *)x; a.y = (float *)y; a.C = nch; a.ch0 = ch0; a.Ctot = first->C; a.n = (long long)n; a.n_out = (long long)(n_out / 2);
        LQB_TRY(static_cast<const FmstStage *>(first)->fill(a.p));
        // enough warps to cover the loop's FP64 dependency chain: ~4 warps per scheduler when the channels allow
        a.cpw = nch >= 32768 ? 32 : (nch >= 4096 ? 16 : 8);        // (measured at 16384 channels: 23.5 / 25.0 / 13.6 GS/s at 32 / 16 / 8 -- idle lanes still cost conversion-unit slots)
        if (const char *e = getenv("LQB_CPW")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) a.cpw = v; }
        LQB_CUDA(fmstereo_launch(a, stream));
        note_kernel(kn, "fmstereo_kernel");
        return LQB_OK;
    }
    if (g.type == Segment::DELAY) {
        const DelayStage *d = static_cast<const DelayStage *>(first);
        // histories are float2-slotted arrays; a real delay line uses them as float rows of the same length
        LQB_CUDA(delay_launch(d->real_io, x, y, d->hist[d->cur].p, d->hist[d->cur ^ 1].p, nch, ch0, (long long)n, d->D, stream));
        note_kernel(kn, "delay_kernel");
        return LQB_OK;
    }
    if (g.type == Segment::BAM) {
        BamArgs a{};
        a.x = (const float2 *)x; a.y = (float *)y; a.C = nch; a.ch0 = ch0; a.Ctot = first->C; a.in_tmajor = in_tmajor ? 1 : 0;
        a.n = (long long)n; a.in_pitch = in_tmajor ? (long long)nch : (long long)n; a.out_pitch = (long long)n_out;
        for (lqb_stage_s *s : g.st) {
            if (s->kind == K_AGC) {
                if (!in_tmajor) return fail(LQB_EINVAL, "internal: in-place gain control needs the time-major hand-off");
                AmTailArgs ag{};
                ag.x = a.x; ag.C = nch; ag.ch0 = ch0; ag.Ctot = first->C; ag.in_tmajor = 1; ag.n = a.n; ag.in_pitch = a.in_pitch;
                LQB_TRY(static_cast<AgcStage *>(s)->fill(ag.agc));
                LQB_CUDA(agc_tmajor_launch(ag, stream));
                note_kernel(kn, "agc_tmajor_kernel");
                (*launches)++;
            }
            else if (s->kind == K_BAM) LQB_TRY(static_cast<BamStage *>(s)->fill(a.p));
            else if (s->kind == K_DEEMPH) { static_cast<DeemphStage *>(s)->fill(a.de); a.has_de = 1; }
        }
        LQB_CUDA(bam_launch(a, stream));
        note_kernel(kn, "bam_kernel");
        return LQB_OK;
    }
    if (g.type == Segment::FIR) {
        const FirStage *f = static_cast<const FirStage *>(first);
        FirArgs a{};
        a.x = (const float2 *)x; a.y = (float2 *)y; a.C = nch; a.ch0 = ch0; a.Ctot = f->C; a.ntaps = (int)f->h.size();
        a.real_io = f->real_io ? 1 : 0;
        a.pair = f->real_io && f->mode == FIR_PLAIN ? 1 : 0;           // two real channels share one packed lane pair
        if (a.pair && (ch0 & 1)) return fail(LQB_EINVAL, "internal: paired real FIR needs an even first channel");
        a.in_real = f->in_r ? 1 : 0; a.out_real = f->out_r ? 1 : 0; a.mode = f->mode; a.delay = f->delay; a.count = f->count;
        a.taps_q = f->hlane_q.empty() ? nullptr : f->taps_q.p;
        a.post_div = f->post_div;
        {
            // every other tap zero in both lanes (Hilbert pairs: quadrature taps on one parity, the in-phase delay the lone
            // tap on the other): the kernel skips that parity's multiply-adds
            int nz[2] = { 0, 0 }, last[2] = { -1, -1 };
            for (size_t k = 0; k < f->h.size(); k++) {
                const bool z = f->h[k] == 0.f && (f->hlane_q.empty() || f->hlane_q[k] == 0.f);
                if (!z) { nz[k & 1]++; last[k & 1] = (int)k; }
            }
            a.skip = 0; a.skip_keep = -1;
            if (f->h.size() >= 8 && !getenv("LQB_FIR_NOSKIP")) {
                if (nz[0] <= 1 && nz[1] > 1) { a.skip = 1; a.skip_keep = last[0]; }
                else if (nz[1] <= 1 && nz[0] > 1) { a.skip = 2; a.skip_keep = last[1]; }
            }
        }
        for (int k = 0; k < 4; k++) a.zero_at[k] = -1;
        if (f->mode == FIR_R2C) {
            int nz = 0;
            for (unsigned long long e : f->ends) {
                const unsigned long long kabs = e + (unsigned long long)f->delay - 1;
                if (kabs < f->count || kabs >= f->count + n) continue;
                if (nz == 4) return fail(LQB_ENOTIMPL, "HilbertTransform: more than 4 call boundaries within one filter delay");
                a.zero_at[nz++] = (long long)(kabs - f->count);
            }
        }
        a.n = (long long)n; a.scale = f->scale; a.taps = f->taps.p; a.hist_in = f->hist[f->cur].p; a.hist_out = f->hist[f->cur ^ 1].p;
        a.utap = 0;
        if (f->hlane_q.empty() && f->h.size() <= (size_t)kFirUTaps && !getenv("LQB_FIR_NOUTAP")) {
            a.utap = 1;
            for (int k = 0; k < kFirUTaps; k++) a.taps_c[k] = k < (int)f->h.size() ? f->h[k] : 0.f;
        }
        LQB_CUDA(fir_launch(a, stream));
        note_kernel(kn, "fir_kernel");
        return LQB_OK;
    }
    if (g.type == Segment::RESAMP_PAR) {
        const ResampStage *r = static_cast<const ResampStage *>(g.st.back());
        ResampP p{}; r->fill(p);
        NcoP q{}; const bool has_nco = g.st.size() == 2;
        if (has_nco) LQB_TRY(static_cast<NcoStage *>(g.st.front())->fill(q));
        LQB_CUDA(resamp_par_launch(p, has_nco ? &q : nullptr, x, y, nch, ch0, r->C, (long long)n, (long long)n_out, stream));
        *launches += resamp_par_launch_count(has_nco, (long long)n_out) - 1;
        note_kernel(kn, has_nco ? "resamp_par_kernel<nco>" : "resamp_par_kernel");
        return LQB_OK;
    }
    if (g.type == Segment::NCO_PAR) {
        NcoP q{}; LQB_TRY(static_cast<NcoStage *>(g.st.front())->fill(q));
        LQB_CUDA(nco_par_launch(q, (const float2 *)x, (float2 *)y, nch, ch0, (long long)n, stream));
        note_kernel(kn, "nco_par_kernel");
        (*launches)++;
        return LQB_OK;
    }
    if (g.type == Segment::SEQ && g.mask == F_IIR && g.st.size() == 1 && static_cast<IirStage *>(g.st[0])->mode == 2) {   // (complex data only)
        // time-parallel blocked scan when the call length allows equal blocks; otherwise the sequential kernel
        IirStage *q = static_cast<IirStage *>(g.st[0]);
        const int B = n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : (n % 64 == 0 ? 64 : 0));
        if (B && n >= (size_t)(4 * B) && q->nsos <= kMaxSos && (((size_t)x) % 16 == 0) && (((size_t)y) % 16 == 0)) {
            const long long K = (long long)n / B, rows = (long long)nch * K;
            LQB_TRY(q->prepare_scan(B));
            LQB_TRY(q->vblk.reserve((size_t)q->nsos * 2 * rows));
            LQB_TRY(q->sin.reserve((size_t)rows * 2 * q->nsos));
            LQB_CUDA(cudaMemsetAsync(q->vblk.p, 0, (size_t)q->nsos * 2 * rows * sizeof(float2), stream));
            SeqArgs a{};
            a.x = x; a.y = y; a.C = (int)rows; a.ch0 = 0; a.Ctot = (int)rows; a.n = B; a.out_pitch = B; a.vec_in = a.vec_out = 1; a.cpw = 32;
            q->fill(a.iir, 0, q->nsos); a.iir.v = q->vblk.p;
            LQB_CUDA(seq_launch(F_IIR, q->nsos, a, stream));
            IirScanArgs sa{};
            sa.y = (float2 *)y; sa.C = nch; sa.ch0 = ch0; sa.Ctot = q->C; sa.nsos = q->nsos; sa.B = B; sa.n = (long long)n;
            sa.H = q->scanH.p; sa.M = q->scanM.p; sa.vblk = q->vblk.p; sa.v = q->v.p; sa.sin = q->sin.p;
            LQB_CUDA(iir_scan_launch(sa, stream));
            note_kernel(kn, "seq_kernel[scan blocks];iir_scan_kernels");
            *launches += 2;
            return LQB_OK;
        }
    }
    SeqArgs a{};
    const bool in_real = ((g.mask | extra_mask) & (F_INREAL | F_INI16)) != 0, out_real = (g.mask & (F_AM | F_FM | F_INREAL)) != 0;
    a.x = x; a.y = y; a.C = nch; a.ch0 = ch0; a.Ctot = first->C; a.n = (long long)n;
    // enough warps to cover the recurrences' latency: aim for >= 12 warps per SM (148 SMs)
    // (only where the per-sample work is light: with the AGC / discriminator in the loop the kernel is issue-bound and
    // idle lanes would only cost slots)
    const bool light = (g.mask & (F_AGC | F_FM | F_AM)) == 0;
    a.cpw = !light ? 32 : (nch >= 148 * 12 * 32 ? 32 : (nch >= 148 * 12 * 16 ? 16 : 8));
    if (const char *e = getenv("LQB_CPW")) { const int v = atoi(e); if (v == 8 || v == 16 || v == 32) a.cpw = (g.mask & F_AM) ? 32 : v; }   // tuning override
    a.out_tmajor = out_tmajor ? 1 : 0; a.out_pitch = out_tmajor ? (long long)nch : (long long)n_out;
    a.vec_in  = ((n * (in_real ? 4 : 8)) % 16 == 0) && (((size_t)x) % 16 == 0);
    a.vec_out = ((n_out * (out_real ? 4 : 8)) % 16 == 0) && (((size_t)y) % 16 == 0);
    for (lqb_stage_s *s : g.st) {
        switch (s->kind) {
        case K_NCO:    LQB_TRY(static_cast<NcoStage *>(s)->fill(a.nco)); break;
        case K_IIR:    static_cast<IirStage *>(s)->fill(a.iir, g.sos0, g.nsos); break;
        case K_RESAMP: static_cast<ResampStage *>(s)->fill(a.rs); break;
        case K_AGC:    LQB_TRY(static_cast<AgcStage *>(s)->fill(a.agc)); break;
        case K_AM:     LQB_TRY(static_cast<AmStage *>(s)->fill(a.am)); break;
        case K_FM:     static_cast<FmStage *>(s)->fill(a.fm); break;
        case K_DEEMPH: static_cast<DeemphStage *>(s)->fill(a.de); break;
        case K_TF:     static_cast<TfStage *>(s)->fill(a.tf); break;
        default: return fail(LQB_EINVAL, "stage kind %d cannot run in the sequential kernel", (int)s->kind);
        }
    }
    // cascade -> single-precision gain loop -> discriminator (config 4): one warp per stage, TMA input ring (pipe.cu)
    if (extra_mask == 0 && !in_real && a.agc.fast && pipe_supported(g.mask, g.nsos) && !getenv("LQB_NO_PIPE") &&
        make_input_tmap(&a.tmap, x, n, (size_t)nch)) {
        a.use_tma = 1;
        LQB_CUDA(pipe_launch(g.nsos, a, stream));
        note_kernel(kn, "pipe_kernel<" + std::to_string(g.nsos) + ">");
        return LQB_OK;
    }
    // the front of the receiver (cascade + decimating resampler): one component per lane (lanes.cu)
    {
        const int lanes = (extra_mask == 0 && !in_real && !getenv("LQB_NO_LANES")) ? lanes_per_channel(g.mask, g.nsos, nch) : 0;
        if (lanes && make_input_tmap(&a.tmap, x, n, (size_t)nch, (unsigned)lanes_box_rows(lanes), 32, io.in_pitch)) {
            ResampStage *r = nullptr;
            for (lqb_stage_s *s : g.st) if (s->kind == K_RESAMP) r = static_cast<ResampStage *>(s);
            DevArr<char> &tb = r->tapbuf[stream];
            LQB_TRY(tb.reserve(lanes_tapstream_bytes((long long)n)));
            LQB_CUDA(lanes_tapstream_launch(a.rs, (long long)n, lanes_lag(g.nsos, lanes), tb.p, stream));
            a.tapstream = tb.p; a.use_tma = 1;
            LQB_CUDA(lanes_launch(g.nsos, lanes, a, stream));
            (*launches)++;
            note_kernel(kn, std::string("tapstream_kernel;") + lanes_kernel_name(g.nsos, lanes));
            return LQB_OK;
        }
    }
    if (io.in_pitch) return fail(LQB_EINVAL, "internal: a time slice needs the lane-split front kernel");
    // (A/B: LQB_NO_LANES=1) the same front with two channels per thread, packed arithmetic (front.cu)
    {
        bool two = nch >= 32768 && extra_mask == 0 && !in_real && front2_supported(g.mask, g.nsos);
        if (const char *e = getenv("LQB_FRONT2")) two = two && atoi(e) != 0;
        if (two && make_input_tmap(&a.tmap, x, n, (size_t)nch, kFront2BoxRows)) {
            a.cpw = 32; a.use_tma = 1;
            if (const char *e = getenv("LQB_FRONT_RING")) { const int v = atoi(e); if (v == 2 || v == 3) front_ring = v; }   // tuning override
            LQB_CUDA(front2_launch(g.nsos, a, stream, front_ring));
            note_kernel(kn, "front2_kernel<" + std::to_string(g.nsos) + ">");
            return LQB_OK;
        }
    }
    // Blackwell data path for the many-channel front kernels: TMA tiled loads instead of per-lane cp.async
    a.use_tma = (a.cpw == 32 && extra_mask == 0 && !in_real && seq_has_tma(g.mask, g.nsos) && make_input_tmap(&a.tmap, x, n, (size_t)nch)) ? 1 : 0;
    // full-rate complex results leave through a bulk tensor store when the rows allow the same box shape
    a.tma_out = 0;
    if (a.use_tma && !(g.mask & F_RS) && !out_real && (n_out & 1) == 0 && make_input_tmap(&a.tmap_out, y, n_out, (size_t)nch)) a.tma_out = 1;
    if (a.use_tma && !(g.mask & F_RS) && !out_real && !a.tma_out) a.use_tma = 0;     // the TMA instantiation stores by TMA only
    LQB_CUDA(seq_launch(g.mask | extra_mask, g.nsos, a, stream));
    note_kernel(kn, "seq_kernel<" + g.name + (a.use_tma ? ",tma>" : (a.cpw == 32 ? ">" : ",cpw" + std::to_string(a.cpw) + ">")));
    return LQB_OK;
}

static size_t elem_bytes(bool real) { return real ? 4 : 8; }

static int chain_validate(lqb_chain_s *c)
{
    if (!c || c->stages.empty()) return fail(LQB_EINVAL, "empty chain");
    {
        std::lock_guard<std::mutex> lk(g_live_mu);
        for (size_t i = 0; i < c->stages.size(); i++)
            if (!g_live.count(c->stages[i])) return fail(LQB_EINVAL, "stage %zu of the chain was destroyed (rebuild the chain after recreating a stage)", i);
    }
    if (c->poisoned) return fail(LQB_EINVAL, "an earlier call on this chain failed part-way (%s); reset the stages and rebuild the chain", c->poison_why.c_str());
    for (size_t i = 0; i < c->stages.size(); i++) {
        if (c->stages[i]->C != c->stages[0]->C) return fail(LQB_EINVAL, "stages of one chain must have the same channel count");
        if (c->stages[i]->device != c->stages[0]->device) return fail(LQB_EINVAL, "stages of one chain must live on one device");
        if (i && c->stages[i]->in_real() != c->stages[i - 1]->out_real())
            return fail(LQB_EINVAL, "stage %zu (%s) input type does not match the previous stage's output", i, kind_name(c->stages[i]->kind));
    }
    return LQB_OK;
}

// all segments over channel range [ch0, ch0+nch); tmp0/tmp1 hold intermediates
// can the first segment take interleaved int16 I/Q directly?
static bool first_takes_i16(const std::vector<Segment> &segs)
{
    return !segs.empty() && segs[0].type == Segment::SEQ && seq_supported(segs[0].mask | F_INI16, segs[0].nsos);
}

static int run_all(lqb_chain_s *c, const std::vector<Segment> &segs, const void *x, void *y, size_t n, int ch0, int nch,
                   char *tmp0, char *tmp1, cudaStream_t stream, int *launches, bool timed = false, bool in_i16 = false, char *cvt = nullptr)
{
    unsigned first_extra = 0;
    if (in_i16) {
        if (first_takes_i16(segs)) first_extra = F_INI16;          // bytes_to_iq fused into the front kernel
        else {                                                      // otherwise one conversion pass, then the usual plan
            LQB_CUDA(i16_to_c64_launch(x, (float2 *)cvt, (long long)nch * (long long)n, stream));
            note_kernel(&c->last_kernels, "i16_to_c64_kernel");
            (*launches)++; x = cvt;
        }
    }
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    if (timed && c->timed_calls.size() < 1024) {
        evs.resize(segs.size());
        for (auto &p : evs) { LQB_CUDA(cudaEventCreate(&p.first)); LQB_CUDA(cudaEventCreate(&p.second)); }
    }
    const void *cur = x; size_t cur_n = n; int flip = 0; bool in_tm = false;
    for (size_t k = 0; k < segs.size(); k++) {
        const size_t on = seg_out_len(segs[k], cur_n);
        void *dst = (k + 1 == segs.size()) ? y : (void *)(flip ? tmp1 : tmp0);
        // a segment of the same IIR stage split by section offset keeps the sample count
        // a decimating sequential kernel hands its output to the AM tail kernel time-major [sample][channel]:
        // both sides then touch HBM with warp-contiguous accesses and neither needs a staging tile
        const bool out_tm = k + 1 < segs.size() && segs[k].type == Segment::SEQ && (segs[k].mask & F_RS) &&
                            (segs[k + 1].type == Segment::AMTAIL || segs[k + 1].type == Segment::BAM);
        if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[k].first, stream));
        if (on > 0 || cur_n > 0) LQB_TRY(run_segment(segs[k], cur, dst, cur_n, on, ch0, nch, stream, in_tm, out_tm, launches, k == 0 ? first_extra : 0u, 3, &c->last_kernels));
        if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[k].second, stream));
        cur = dst; cur_n = on; flip ^= 1; in_tm = out_tm;
    }
    if (!evs.empty()) c->timed_calls.push_back(std::move(evs));
    return LQB_OK;
}

static size_t max_intermediate_bytes(const std::vector<Segment> &segs, size_t n, size_t rows)
{
    size_t worst = 0, cur = n;
    for (size_t k = 0; k + 1 < segs.size(); k++) {
        cur = seg_out_len(segs[k], cur);
        worst = std::max(worst, cur * rows * elem_bytes(segs[k].out_real >= 0 ? segs[k].out_real != 0 : segs[k].st.back()->out_real()));
    }
    return worst;
}

static void advance_all(lqb_chain_s *c, size_t n) { for (auto *s : c->stages) { const size_t on = s->out_len(n); s->advance(n); n = on; } }

static int chain_out_len(lqb_chain_s *c, size_t n, size_t *n_out) { for (auto *s : c->stages) n = s->out_len(n); *n_out = n; return LQB_OK; }

// ---- overlapped device calls -------------------------------------------------------------------------------------
// The AM receiver is a full-rate front kernel (97.6 % of the samples, FP32-pipe and HBM bound) and a decimated-rate tail
// (gain loop, carrier PLL, two 51-tap filters: latency bound, 18 % of a call's time on 2.4 % of the samples).  Calls
// carry state, so call k+1's front only needs call k's FRONT: with overlap enabled the tail of call k is issued on the
// chain's own stream behind an event and the caller's stream goes straight on to the next front.  The front then
// stages through a 2-deep ring (20 KB per CTA) so that one tail CTA fits beside its seven CTAs on every SM, and the
// tail's dependent chains fill issue slots the front leaves empty.  The caller's stream should be created with a
// higher priority than the default (lqb_stream_create) so that a front is dispatched ahead of the tail queued before it.
// Results of an overlapped call are complete once lqb_chain_wait(chain, stream) has been waited on.
static bool overlappable(const lqb_chain_s *c, const std::vector<Segment> &segs, bool in_i16)
{
    return c->overlap && !in_i16 && segs.size() == 2 && segs[0].type == Segment::SEQ && (segs[0].mask & F_RS) &&
           (segs[1].type == Segment::AMTAIL || segs[1].type == Segment::BAM);
}

static int chain_join_tails(lqb_chain_s *c, cudaStream_t stream)
{
    for (int b = 0; b < 2; b++)
        if (c->tail_pending[b]) { LQB_CUDA(cudaStreamWaitEvent(stream, c->ev_tail[b], 0)); }
    return LQB_OK;
}

static int run_overlapped(lqb_chain_s *c, const std::vector<Segment> &segs, const void *x, void *y, size_t n, int C, size_t tmpb, cudaStream_t stream)
{
    if (!c->tail_stream) {
        int lo = 0, hi = 0;
        LQB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        LQB_CUDA(cudaStreamCreateWithPriority(&c->tail_stream, cudaStreamNonBlocking, lo));
        for (int b = 0; b < 2; b++) {
            LQB_CUDA(cudaEventCreateWithFlags(&c->ev_front[b], cudaEventDisableTiming));
            LQB_CUDA(cudaEventCreateWithFlags(&c->ev_tail[b], cudaEventDisableTiming));
        }
    }
    LQB_TRY(c->d_tmp[0].reserve(tmpb)); LQB_TRY(c->d_tmp[1].reserve(tmpb));
    const int b = (int)(c->ov_calls++ & 1u);
    char *tmp = c->d_tmp[b].p;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    if (c->timing && c->timed_calls.size() < 1024) {
        evs.resize(2);
        for (auto &p : evs) { LQB_CUDA(cudaEventCreate(&p.first)); LQB_CUDA(cudaEventCreate(&p.second)); }
    }
    const size_t on0 = seg_out_len(segs[0], n), on = seg_out_len(segs[1], on0);
    // the front writes the hand-off buffer the tail of two calls ago read
    if (c->tail_pending[b]) LQB_CUDA(cudaStreamWaitEvent(stream, c->ev_tail[b], 0));
    if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[0].first, stream));
    LQB_TRY(run_segment(segs[0], x, tmp, n, on0, 0, C, stream, false, true, &c->last_launches, 0u, 2, &c->last_kernels));
    if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[0].second, stream));
    LQB_CUDA(cudaEventRecord(c->ev_front[b], stream));
    LQB_CUDA(cudaStreamWaitEvent(c->tail_stream, c->ev_front[b], 0));
    if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[1].first, c->tail_stream));
    if (on > 0 || on0 > 0) LQB_TRY(run_segment(segs[1], tmp, y, on0, on, 0, C, c->tail_stream, true, false, &c->last_launches, 0u, 3, &c->last_kernels));
    if (!evs.empty()) LQB_CUDA(cudaEventRecord(evs[1].second, c->tail_stream));
    LQB_CUDA(cudaEventRecord(c->ev_tail[b], c->tail_stream));
    c->tail_pending[b] = true;
    if (!evs.empty()) c->timed_calls.push_back(std::move(evs));
    return LQB_OK;
}

// ---- SM partitions (green contexts) --------------------------------------------------------------------------------------
// A front of few channels leaves SMs idle but needs every one of its warps alone on a scheduler; tail kernels placed beside
// them take their issue slots one for one (run_timepipe).  A green context confines a stream's kernels to a set of SMs: the
// front gets as many SMs as its warps fill four to an SM, the tail the rest.  Driver API, fetched at run time like the
// tensor-map encoder; every failure falls back to ordinary streams.
template <class F> static F drv(const char *name)
{
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
    return (F)p;
}
static bool sm_partition_create(SmPartition &sp, int want_front_sms)
{
    typedef CUresult (*GetDev)(CUdevice *, int);
    typedef CUresult (*GetRes)(CUdevice, CUdevResource *, CUdevResourceType);
    typedef CUresult (*Split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int);
    typedef CUresult (*GenDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int);
    typedef CUresult (*GCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int);
    typedef CUresult (*GStream)(CUstream *, CUgreenCtx, unsigned int, int);
    static GetDev get_dev = drv<GetDev>("cuDeviceGet");
    static GetRes get_res = drv<GetRes>("cuDeviceGetDevResource");
    static Split split = drv<Split>("cuDevSmResourceSplitByCount");
    static GenDesc gen = drv<GenDesc>("cuDevResourceGenerateDesc");
    static GCreate gcreate = drv<GCreate>("cuGreenCtxCreate");
    static GStream gstream = drv<GStream>("cuGreenCtxStreamCreate");
    if (!get_dev || !get_res || !split || !gen || !gcreate || !gstream) return false;
    int ord = 0; if (cudaGetDevice(&ord) != cudaSuccess) return false;
    CUdevice dev; if (get_dev(&dev, ord) != CUDA_SUCCESS) return false;
    CUdevResource all{}, grp{}, rest{};
    if (get_res(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned int nb = 1;
    if (split(&grp, &nb, &all, &rest, 0, (unsigned)want_front_sms) != CUDA_SUCCESS || nb != 1) return false;
    if (grp.sm.smCount < (unsigned)want_front_sms || rest.sm.smCount < 8) return false;
    CUdevResourceDesc da = nullptr, db = nullptr;
    if (gen(&da, &grp, 1) != CUDA_SUCCESS || gen(&db, &rest, 1) != CUDA_SUCCESS) return false;
    if (gcreate(&sp.ga, da, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (gcreate(&sp.gb, db, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    CUstream fa = nullptr, fb = nullptr, fc = nullptr;
    if (gstream(&fa, sp.ga, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS || gstream(&fb, sp.gb, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
        gstream(&fc, sp.gb, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) return false;
    sp.front = (cudaStream_t)fa; sp.tail = (cudaStream_t)fb; sp.tail2 = (cudaStream_t)fc; sp.sm_a = (int)grp.sm.smCount; sp.sm_b = (int)rest.sm.smCount; sp.ok = true;
    return true;
}
// ---- few channels: front and tail as a pipeline along TIME -------------------------------------------------------------
// With few channels both segments are latency-bound (one warp per scheduler walks a recurrence) and most of the machine is
// idle, so a call is cut into time slices: the front kernel of slice j + 1 runs on the caller's stream while the tail of
// slice j (gain loop, carrier loop, filters) runs on the chain's own stream.  Every stage carries its state from slice to
// slice exactly as it does from call to call (the host bookkeeping advances per slice), so the results are bit-identical
// to the unsliced call; what stays exposed behind the front is the last slice's tail instead of the whole tail.
// SMs the front of C channels fills with four lone warps each (rounded up to a multiple of eight), or 0 when too few are left over
static int partition_front_sms(const std::vector<Segment> &segs, int C)
{
    const int lanes = lanes_per_channel(segs[0].mask, segs[0].nsos, C);
    if (!lanes) return 0;
    const long long warps = ((long long)C * lanes + 31) / 32;
    const long long want = ((warps + 3) / 4 + 7) / 8 * 8;
    return (want >= 8 && want <= 100) ? (int)want : 0;                                 // (at least 48 SMs stay for the tail)
}
static SmPartition *timepipe_partition(lqb_chain_s *, const std::vector<Segment> &segs, int C)
{
    if (getenv("LQB_NO_PARTITION")) return nullptr;
    const int want = partition_front_sms(segs, C);
    if (!want) return nullptr;
    // one partition per device and front size for the whole process (green contexts are not free, and chains come and go):
    // chains that share one only share its streams' order
    static std::mutex mu; static std::map<std::pair<int, int>, SmPartition> parts;
    int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    SmPartition &sp = parts[std::make_pair(dev, want)];
    if (!sp.ok && !sp.ga) sm_partition_create(sp, want);           // (tried once per device and size)
    return sp.ok ? &sp : nullptr;
}
static int timepipe_slices(lqb_chain_s *c, const std::vector<Segment> &segs, const void *x, size_t n, int C, bool in_i16)
{
    if (getenv("LQB_NO_TIMEPIPE") || getenv("LQB_NO_LANES") || c->overlap || in_i16 || segs.size() != 2) return 0;
    if (segs[0].type != Segment::SEQ || segs[0].mask != (F_IIR | F_RS) || segs[1].type != Segment::AMTAIL) return 0;
    if (n < 16384 || (n & 1) || (((size_t)x) & 15) || !lanes_per_channel(segs[0].mask, segs[0].nsos, C)) return 0;
    // Up to 1024 channels the slices pay on ordinary streams.  Beyond that the concurrent tail takes the front's issue slots one
    // for one (8192 channels: 2.01 ms unsliced, 2.20 sliced) -- unless the two run on DISJOINT SMs: with an SM partition
    // (green contexts: the front on the SMs its lone warps fill, the tail on the rest) 4096 channels go from 1.95 to 1.58 ms.
    // At 8192 channels the front needs 128 SMs and 20 are too few for the tail to keep up (1.99 ms): no slices there.
    int cmax = 1024;
    if (const char *e = getenv("LQB_TIMEPIPE_MAX")) cmax = atoi(e);                    // tuning override
    if (C > cmax && !timepipe_partition(c, segs, C)) return 0;
    int k = 6;
    if (const char *e = getenv("LQB_TIMEPIPE_SLICES")) { const int v = atoi(e); if (v >= 2 && v <= 64) k = v; }
    return k;
}

static int run_timepipe(lqb_chain_s *c, const std::vector<Segment> &segs, const void *x, void *y, size_t n, int C, char *tmp,
                        cudaStream_t stream, int K, bool timed)
{
    // The front must keep every one of its CTAs resident (a CTA that finds its SM taken waits for a whole slice), so it runs
    // on a stream of its own with the HIGHEST priority and the small tail kernels fill what it leaves: with the tail ahead in
    // the block scheduler's queue, 8192 channels ran 2.5 ms instead of 1.6 (measured).
    if (!c->pipe_stream) {
        int lo = 0, hi = 0;
        LQB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        LQB_CUDA(cudaStreamCreateWithPriority(&c->front_stream, cudaStreamNonBlocking, hi));
        LQB_CUDA(cudaStreamCreateWithPriority(&c->pipe_stream, cudaStreamNonBlocking, lo));
    }
    while (c->ev_pipe.size() < (size_t)2 * K + 3) { cudaEvent_t e; LQB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); c->ev_pipe.push_back(e); }
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> evs;
    cudaStream_t fs = c->front_stream, ts = c->pipe_stream, ta = nullptr;      // (ta: the gain loop's stream inside an SM partition)
    if (getenv("LQB_TIMEPIPE_SERIAL")) ts = fs;                         // A/B: the slices alone, nothing concurrent
    // SM partition: the front on the SMs its warps fill (four lone warps per SM), the tail on the others
    bool parted = false;
    if (SmPartition *sp = timepipe_partition(c, segs, C)) { fs = sp->front; ts = sp->tail; ta = sp->tail2; parted = true; }
    struct Flag { bool on; Flag(bool v) : on(v) { if (on) t_no_tail_ahead = true; } ~Flag() { if (on) t_no_tail_ahead = false; } } flag(parted);
    // both follow everything already queued on the caller's stream
    LQB_CUDA(cudaEventRecord(c->ev_pipe[K], stream));
    LQB_CUDA(cudaStreamWaitEvent(fs, c->ev_pipe[K], 0));
    LQB_CUDA(cudaStreamWaitEvent(ts, c->ev_pipe[K], 0));
    if (ta) LQB_CUDA(cudaStreamWaitEvent(ta, c->ev_pipe[K], 0));
    if (timed && c->timed_calls.size() < 1024) {
        evs.resize(2);
        for (auto &p : evs) { LQB_CUDA(cudaEventCreate(&p.first)); LQB_CUDA(cudaEventCreate(&p.second)); }
        LQB_CUDA(cudaEventRecord(evs[0].first, fs));
    }
    size_t on_total; chain_out_len(c, n, &on_total);
    // Slice lengths in the proportion 4 : 4 : ... : 4 : 2 : 1 (whole tiles of the front kernel): what stays exposed behind the
    // front is the LAST slice's tail, so the last slices are the short ones (equal sixths left 0.13 ms of a one-channel block
    // exposed; LQB_TIMEPIPE_EQUAL=1 restores them for A/B runs)
    std::vector<size_t> cut;
    {
        const bool equal = getenv("LQB_TIMEPIPE_EQUAL") != nullptr || K < 3;
        size_t wsum = 0; std::vector<size_t> w((size_t)K);
        for (int i = 0; i < K; i++) { w[(size_t)i] = equal ? 1 : (i < K - 2 ? 4 : (i == K - 2 ? 2 : 1)); wsum += w[(size_t)i]; }
        size_t t = 0;
        for (int i = 0; i < K && t < n; i++) {
            size_t len = i == K - 1 ? n - t : std::max<size_t>(16, (n * w[(size_t)i] / wsum + 15) / 16 * 16);
            len = std::min(len, n - t);
            cut.push_back(len); t += len;
        }
        if (t < n) cut.back() += n - t;
    }
    size_t off = 0; int j = 0; size_t t0 = 0;
    for (; j < (int)cut.size(); t0 += cut[(size_t)j], j++) {
        const size_t ns = cut[(size_t)j];
        const size_t on0 = seg_out_len(segs[0], ns), on = seg_out_len(segs[1], on0);
        if (off + on > on_total) return fail(LQB_ESIZE, "internal: time slices produce more than the call");
        char *hand = tmp + off * (size_t)C * 8;                                        // time-major [sample][channel]: slices are contiguous
        SegIO fio; fio.in_pitch = n;
        LQB_TRY(run_segment(segs[0], (const char *)x + t0 * 8, hand, ns, on0, 0, C, fs, false, true, &c->last_launches, 0u, 3, j == 0 ? &c->last_kernels : nullptr, fio));
        LQB_CUDA(cudaEventRecord(c->ev_pipe[j], fs));
        LQB_CUDA(cudaStreamWaitEvent(ts, c->ev_pipe[j], 0));
        SegIO tio; tio.out_pitch = on_total;
        if (on > 0 || on0 > 0) {
            if (parted && ta) {
                // three stages: front | gain loop | demodulator, the last two on two streams of the tail's SMs
                LQB_CUDA(cudaStreamWaitEvent(ta, c->ev_pipe[j], 0));
                tio.tail_part = 1;
                LQB_TRY(run_segment(segs[1], hand, (char *)y + off * 4, on0, on, 0, C, ta, true, false, &c->last_launches, 0u, 3, j == 0 ? &c->last_kernels : nullptr, tio));
                LQB_CUDA(cudaEventRecord(c->ev_pipe[K + 3 + j], ta));
                LQB_CUDA(cudaStreamWaitEvent(ts, c->ev_pipe[K + 3 + j], 0));
                tio.tail_part = 2;
            }
            LQB_TRY(run_segment(segs[1], hand, (char *)y + off * 4, on0, on, 0, C, ts, true, false, &c->last_launches, 0u, 3, j == 0 ? &c->last_kernels : nullptr, tio));
        }
        advance_all(c, ns);
        off += on;
    }
    if (off != on_total) return fail(LQB_ESIZE, "internal: time slices produced %zu samples, the call %zu", off, on_total);
    LQB_CUDA(cudaEventRecord(c->ev_pipe[K + 1], fs));
    LQB_CUDA(cudaEventRecord(c->ev_pipe[K + 2], ts));
    if (!evs.empty()) {
        // front: first launch to last front kernel; tail: what stays exposed behind the front
        LQB_CUDA(cudaEventRecord(evs[0].second, fs));
        LQB_CUDA(cudaStreamWaitEvent(ts, c->ev_pipe[K + 1], 0));
        LQB_CUDA(cudaEventRecord(evs[1].first, fs)); LQB_CUDA(cudaEventRecord(evs[1].second, ts));
        LQB_CUDA(cudaEventRecord(c->ev_pipe[K + 2], ts));
        c->timed_calls.push_back(std::move(evs));
    }
    LQB_CUDA(cudaStreamWaitEvent(stream, c->ev_pipe[K + 1], 0));
    LQB_CUDA(cudaStreamWaitEvent(stream, c->ev_pipe[K + 2], 0));
    note_kernel(&c->last_kernels, "(x" + std::to_string(j) + " time slices, tail on its own stream" + (parted ? std::string(", front and tail on disjoint SMs") : std::string()) + ")");
    return LQB_OK;
}

// a failure after the first launch of a call leaves device state (filter memories, rings) ahead of the host-side phase /
// count bookkeeping; every later call would be silently misaligned, so the chain refuses them until it is reset
static int chain_poison(lqb_chain_s *c, int rc)
{
    if (c->last_launches > 0) { c->poisoned = true; c->poison_why = g_err; }
    return rc;
}

static int chain_execute_dev(lqb_chain_s *c, const void *x, size_t n, void *y, size_t cap, size_t *n_out, cudaStream_t stream, bool in_i16 = false)
{
    LQB_TRY(chain_validate(c));
    if (in_i16 && c->stages.front()->in_real()) return fail(LQB_EINVAL, "int16 I/Q input needs a chain that starts with a complex-input stage");
    for (auto *s : c->stages) LQB_TRY(s->ensure());
    std::vector<Segment> segs; LQB_TRY(build_plan(c, segs));
    size_t on; chain_out_len(c, n, &on);
    if (n_out) *n_out = on;
    if (on > cap) return fail(LQB_ESIZE, "output needs %zu samples per channel, capacity is %zu", on, cap);
    c->last_launches = 0; c->last_kernels.clear();
    if (n == 0) return LQB_OK;
    const int C = c->stages[0]->C;
    const size_t tmpb = max_intermediate_bytes(segs, n, (size_t)C);
    if (segs.size() > 1) LQB_TRY(c->d_tmp[0].reserve(tmpb));
    if (segs.size() > 2) LQB_TRY(c->d_tmp[1].reserve(tmpb));
    if (in_i16 && !first_takes_i16(segs)) LQB_TRY(c->d_cvt.reserve((size_t)C * n * 8));
    int rc;
    if (const int K = timepipe_slices(c, segs, x, n, C, in_i16)) {
        LQB_TRY(chain_join_tails(c, stream));
        rc = run_timepipe(c, segs, x, y, n, C, c->d_tmp[0].p, stream, K, c->timing);
        return rc == LQB_OK ? rc : chain_poison(c, rc);           // (the slices advanced the host bookkeeping themselves)
    }
    if (overlappable(c, segs, in_i16)) {
        rc = run_overlapped(c, segs, x, y, n, C, tmpb, stream);
    } else {
        LQB_TRY(chain_join_tails(c, stream));                    // tails of earlier overlapped calls come first
        rc = run_all(c, segs, x, y, n, 0, C, c->d_tmp[0].p, c->d_tmp[1].p, stream, &c->last_launches, c->timing, in_i16, c->d_cvt.p);
    }
    if (rc != LQB_OK) return chain_poison(c, rc);
    advance_all(c, n);
    return LQB_OK;
}

// ---- pageable host input ---------------------------------------------------------------------------------------------
// A plain numpy array is pageable memory: cudaMemcpy*Async from it is a synchronous, single-threaded staging copy inside the
// driver (10.8 GB/s measured here against 55 GB/s from pinned memory).  The reference's callers hand over exactly such arrays
// (README.md:60-63), so the host path stages them itself: a few host threads copy the slice's rows into a ring of pinned
// bounce buffers while the previous buffer crosses PCIe.
class StagePool {
    std::vector<std::thread> th;
    std::mutex m, callers; std::condition_variable cv, done;      // (callers: one parallel copy at a time -- chains may run on several host threads)
    std::function<void(int, int)> job; int gen = 0, pending = 0;
public:
    explicit StagePool(int n)
    {
        for (int i = 0; i < n; i++)
            th.emplace_back([this, i, n] {
                int seen = 0;
                for (;;) {
                    std::function<void(int, int)> f;
                    { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [&] { return gen != seen; }); seen = gen; f = job; }
                    f(i, n);
                    { std::lock_guard<std::mutex> lk(m); if (--pending == 0) done.notify_one(); }
                }
            });
        for (auto &t : th) t.detach();               // (workers live as long as the process: the library is never unloaded under Python)
    }
    int size() const { return (int)th.size(); }
    void run(const std::function<void(int, int)> &f)
    {
        std::lock_guard<std::mutex> one(callers);
        std::unique_lock<std::mutex> lk(m);
        job = f; pending = (int)th.size(); gen++;
        cv.notify_all();
        done.wait(lk, [&] { return pending == 0; });
    }
};
static StagePool *stage_pool()
{
    static StagePool *pool = nullptr; static std::once_flag once;
    std::call_once(once, [] {
        int n = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char *e = getenv("LQB_STAGE_THREADS")) { const int v = atoi(e); if (v >= 1 && v <= 64) n = v; }
        pool = new StagePool(n);
    });
    return pool;
}
static bool host_pinned(const void *p)
{
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
// rows x rowbytes from pageable host memory (row pitch src_pitch) to a dense device block, through the chain's bounce buffers
static int staged_h2d(lqb_chain_s *c, char *dst, const char *src, size_t src_pitch, size_t rowbytes, size_t rows, cudaStream_t copy)
{
    const size_t want = (size_t)64 << 20;
    if (c->pin_bytes < want) {
        for (int b = 0; b < lqb_chain_s::kPin; b++) {
            if (c->pin_buf[b]) { LQB_CUDA(cudaFreeHost(c->pin_buf[b])); c->pin_buf[b] = nullptr; }
            LQB_CUDA(cudaHostAlloc(&c->pin_buf[b], want, cudaHostAllocDefault));
            if (!c->ev_pin[b]) LQB_CUDA(cudaEventCreateWithFlags(&c->ev_pin[b], cudaEventDisableTiming));
            c->pin_used[b] = false;
        }
        c->pin_bytes = want;
    }
    if (rowbytes > c->pin_bytes) {                   // (rows longer than a bounce buffer: let the driver stage them)
        LQB_CUDA(cudaMemcpy2DAsync(dst, rowbytes, src, src_pitch, rowbytes, rows, cudaMemcpyHostToDevice, copy));
        return LQB_OK;
    }
    StagePool *pool = stage_pool();
    const size_t per = std::max<size_t>(1, c->pin_bytes / rowbytes);
    for (size_t r0 = 0; r0 < rows; r0 += per) {
        const size_t nr = std::min(per, rows - r0);
        const int b = (int)(c->pin_turn++ % lqb_chain_s::kPin);
        if (c->pin_used[b]) LQB_CUDA(cudaEventSynchronize(c->ev_pin[b]));       // its last H2D has left the buffer
        char *buf = (char *)c->pin_buf[b];
        const char *s0 = src + r0 * src_pitch;
        pool->run([=](int i, int nthr) {
            const size_t a = nr * (size_t)i / (size_t)nthr, z = nr * (size_t)(i + 1) / (size_t)nthr;
            for (size_t r = a; r < z; r++) memcpy(buf + r * rowbytes, s0 + r * src_pitch, rowbytes);
        });
        LQB_CUDA(cudaMemcpyAsync(dst + r0 * rowbytes, buf, nr * rowbytes, cudaMemcpyHostToDevice, copy));
        LQB_CUDA(cudaEventRecord(c->ev_pin[b], copy));
        c->pin_used[b] = true;
    }
    return LQB_OK;
}

// Host-pointer call, large input, decimating chain: the block is cut along TIME.  Every slice carries all channels
// (so the kernels run at full occupancy and a slice's kernels take a fraction of the call's), slice k+1 crosses PCIe
// (one strided 2-D copy from the caller's rows) while slice k is processed, the stages carry their state from slice
// to slice exactly as they do from call to call, and the small output collects on the device and goes back once.
// What stays exposed behind the H2D stream is one slice's kernels instead of a whole channel group's recurrence.
static int chain_execute_host_sliced(lqb_chain_s *c, const std::vector<Segment> &segs, const void *x, size_t n, void *y, size_t on_total,
                                     bool in_i16, size_t ib, size_t ob, bool need_cvt)
{
    const int C = c->stages[0]->C;
    size_t nsl = 16;
    if (const char *e = getenv("LQB_SLICES")) { const int v = atoi(e); if (v >= 1 && v <= 64) nsl = (size_t)v; }
    size_t slice = ((n + nsl - 1) / nsl + 255) / 256 * 256;
    slice = std::max<size_t>(slice, 4096);
    for (int s = 0; s < 2; s++) if (!c->streams[s]) LQB_CUDA(cudaStreamCreate(&c->streams[s]));
    for (int b = 0; b < 2; b++) {
        if (!c->ev_ready[b]) LQB_CUDA(cudaEventCreateWithFlags(&c->ev_ready[b], cudaEventDisableTiming));
        if (!c->ev_free[b]) LQB_CUDA(cudaEventCreateWithFlags(&c->ev_free[b], cudaEventDisableTiming));
        LQB_TRY(c->h_in[b].reserve((size_t)C * slice * ib));
    }
    cudaStream_t copy = c->streams[0], comp = c->streams[1];
    const bool pinned = host_pinned(x) || getenv("LQB_NO_STAGING");            // (A/B: the driver's own staging of pageable memory)
    const size_t tmpb = max_intermediate_bytes(segs, slice, (size_t)C);
    if (segs.size() > 1) LQB_TRY(c->h_tmp[0][0].reserve(tmpb));
    if (segs.size() > 2) LQB_TRY(c->h_tmp[0][1].reserve(tmpb));
    if (need_cvt) LQB_TRY(c->h_cvt[0].reserve((size_t)C * slice * 8));
    LQB_TRY(c->h_out[0].reserve(std::max<size_t>(16, (size_t)C * on_total * ob)));
    size_t on_max = 0; { size_t v; chain_out_len(c, slice, &v); on_max = v + 4 * c->stages.size() + 16; }
    LQB_TRY(c->h_out[1].reserve(std::max<size_t>(16, (size_t)C * on_max * ob)));
    size_t off = 0; int k = 0;
    for (size_t t0 = 0; t0 < n; t0 += slice, k++) {
        const size_t ns = std::min(slice, n - t0);
        const int b = k & 1;
        if (k >= 2) LQB_CUDA(cudaStreamWaitEvent(copy, c->ev_free[b], 0));
        if (pinned) LQB_CUDA(cudaMemcpy2DAsync(c->h_in[b].p, ns * ib, (const char *)x + t0 * ib, n * ib, ns * ib, (size_t)C, cudaMemcpyHostToDevice, copy));
        else LQB_TRY(staged_h2d(c, c->h_in[b].p, (const char *)x + t0 * ib, n * ib, ns * ib, (size_t)C, copy));
        LQB_CUDA(cudaEventRecord(c->ev_ready[b], copy));
        LQB_CUDA(cudaStreamWaitEvent(comp, c->ev_ready[b], 0));
        size_t on_k; chain_out_len(c, ns, &on_k);
        if (on_k > on_max || off + on_k > on_total) return fail(LQB_ESIZE, "internal: slice produced %zu samples, expected at most %zu", on_k, on_max);
        LQB_TRY(run_all(c, segs, c->h_in[b].p, c->h_out[1].p, ns, 0, C, c->h_tmp[0][0].p, c->h_tmp[0][1].p, comp, &c->last_launches, false, in_i16, c->h_cvt[0].p));
        LQB_CUDA(cudaEventRecord(c->ev_free[b], comp));
        if (on_k) LQB_CUDA(cudaMemcpy2DAsync(c->h_out[0].p + off * ob, on_total * ob, c->h_out[1].p, on_k * ob, on_k * ob, (size_t)C, cudaMemcpyDeviceToDevice, comp));
        advance_all(c, ns);
        off += on_k;
    }
    LQB_CUDA(cudaStreamSynchronize(comp));
    if (off != on_total) return fail(LQB_ESIZE, "internal: slices produced %zu samples, the call %zu", off, on_total);
    if (on_total) LQB_CUDA(cudaMemcpy(y, c->h_out[0].p, (size_t)C * on_total * ob, cudaMemcpyDeviceToHost));
    return LQB_OK;
}

static int chain_execute_host(lqb_chain_s *c, const void *x, size_t n, void *y, size_t cap, size_t *n_out, bool in_i16 = false)
{
    LQB_TRY(chain_validate(c));
    if (in_i16 && c->stages.front()->in_real()) return fail(LQB_EINVAL, "int16 I/Q input needs a chain that starts with a complex-input stage");
    for (auto *s : c->stages) LQB_TRY(s->ensure());
    std::vector<Segment> segs; LQB_TRY(build_plan(c, segs));
    size_t on; chain_out_len(c, n, &on);
    if (n_out) *n_out = on;
    if (on > cap) return fail(LQB_ESIZE, "output needs %zu samples per channel, capacity is %zu", on, cap);
    c->last_launches = 0; c->last_kernels.clear();
    if (n == 0) return LQB_OK;
    // tails of earlier overlapped device calls still update the carried state on the chain's tail stream
    if (c->tail_stream && (c->tail_pending[0] || c->tail_pending[1])) {
        LQB_CUDA(cudaStreamSynchronize(c->tail_stream));
        c->tail_pending[0] = c->tail_pending[1] = false;
    }
    const int C = c->stages[0]->C;
    const size_t ib = in_i16 ? 4 : elem_bytes(c->stages.front()->in_real()), ob = elem_bytes(c->stages.back()->out_real());
    const bool need_cvt = in_i16 && !first_takes_i16(segs);
    {
        // time slices when the call is big, decimates (output fits the device-side collection buffer) and no stage
        // gives meaning to call boundaries (blocked-scan IIR sizes its blocks per call; HilbertTransform's real branch
        // reproduces the reference's per-call overrun)
        bool sliceable = (size_t)C * n * ib >= ((size_t)32 << 20) && n >= 32768 && (size_t)C * on * ob <= ((size_t)256 << 20) && on * 4 <= n;
        for (auto *st : c->stages) {
            if (st->kind == K_IIR && static_cast<IirStage *>(st)->mode == 2) sliceable = false;
            if (st->kind == K_FIR && static_cast<FirStage *>(st)->mode == FIR_R2C) sliceable = false;
        }
        if (const char *e = getenv("LQB_HOST_MODE")) sliceable = sliceable && strcmp(e, "chunk") != 0;
        if (sliceable) { const int rc = chain_execute_host_sliced(c, segs, x, n, y, on, in_i16, ib, ob, need_cvt); return rc == LQB_OK ? rc : chain_poison(c, rc); }
    }
    // Channel chunks on three streams: the H2D of chunk i+1 overlaps the kernels of chunk i.  A sequential
    // kernel takes about as long for 64 channels as for 64K (it is bound by the per-channel recurrence), so
    // chunks are few and large: an eighth of the call, at least 64 MB of input.
    const size_t in_total = (size_t)C * n * ib;
    size_t chunk = std::max<size_t>(1, std::max<size_t>((size_t)64 << 20, in_total / 8) / std::max<size_t>(1, n * ib));
    if (chunk >= 64) chunk = chunk / 64 * 64;
    else if (chunk > 1) chunk &= ~(size_t)1;                       // paired real FIR rows: chunks start on even channels
    else if (C > 1) chunk = 2;
    chunk = std::min<size_t>(chunk, (size_t)C);
    for (auto *st : c->stages) if (st->kind == K_IIR && static_cast<IirStage *>(st)->mode == 2) chunk = (size_t)C;   // scan scratch is per stage
    for (int s = 0; s < lqb_chain_s::kStreams; s++) if (!c->streams[s]) LQB_CUDA(cudaStreamCreate(&c->streams[s]));
    const size_t nchunks = ((size_t)C + chunk - 1) / chunk;
    const size_t tmpb = max_intermediate_bytes(segs, n, chunk);
    // A D2H into pageable memory blocks the host, which would serialise the chunks; when the whole output is
    // small (decimating chains) it stays on the device until every chunk is enqueued and goes back in one copy.
    const size_t out_total = (size_t)C * on * ob;
    const bool deferred = out_total <= ((size_t)256 << 20);
    if (deferred) LQB_TRY(c->h_out[0].reserve(std::max<size_t>(16, out_total)));
    for (size_t s = 0; s < std::min<size_t>(nchunks, lqb_chain_s::kStreams); s++) {
        LQB_TRY(c->h_in[s].reserve(chunk * n * ib));
        if (need_cvt) LQB_TRY(c->h_cvt[s].reserve(chunk * n * 8));
        if (!deferred) LQB_TRY(c->h_out[s].reserve(std::max<size_t>(16, chunk * on * ob)));
        if (segs.size() > 1) LQB_TRY(c->h_tmp[s][0].reserve(tmpb));
        if (segs.size() > 2) LQB_TRY(c->h_tmp[s][1].reserve(tmpb));
    }
    if (nchunks == 1 && deferred) {
        if (const int K = timepipe_slices(c, segs, c->h_in[0].p, n, C, in_i16)) {
            LQB_CUDA(cudaMemcpyAsync(c->h_in[0].p, x, (size_t)C * n * ib, cudaMemcpyHostToDevice, c->streams[0]));
            const int rc = run_timepipe(c, segs, c->h_in[0].p, c->h_out[0].p, n, C, c->h_tmp[0][0].p, c->streams[0], K, false);
            if (rc != LQB_OK) return chain_poison(c, rc);
            LQB_CUDA(cudaStreamSynchronize(c->streams[0]));
            if (on) LQB_CUDA(cudaMemcpy(y, c->h_out[0].p, out_total, cudaMemcpyDeviceToHost));
            return LQB_OK;
        }
    }
    const bool chunk_pinned = host_pinned(x) || getenv("LQB_NO_STAGING");
    for (size_t k = 0; k < nchunks; k++) {
        const int s = (int)(k % lqb_chain_s::kStreams);
        const size_t c0 = k * chunk, nc = std::min(chunk, (size_t)C - c0);
        char *dst = deferred ? c->h_out[0].p + c0 * on * ob : c->h_out[s].p;
        if (!chunk_pinned && nc * n * ib >= ((size_t)16 << 20)) LQB_TRY(staged_h2d(c, c->h_in[s].p, (const char *)x + c0 * n * ib, n * ib, n * ib, nc, c->streams[s]));
        else LQB_CUDA(cudaMemcpyAsync(c->h_in[s].p, (const char *)x + c0 * n * ib, nc * n * ib, cudaMemcpyHostToDevice, c->streams[s]));
        { const int rc = run_all(c, segs, c->h_in[s].p, dst, n, (int)c0, (int)nc, c->h_tmp[s][0].p, c->h_tmp[s][1].p, c->streams[s], &c->last_launches, false, in_i16, c->h_cvt[s].p); if (rc != LQB_OK) return chain_poison(c, rc); }
        if (on && !deferred) LQB_CUDA(cudaMemcpyAsync((char *)y + c0 * on * ob, dst, nc * on * ob, cudaMemcpyDeviceToHost, c->streams[s]));
    }
    for (int s = 0; s < lqb_chain_s::kStreams; s++) LQB_CUDA(cudaStreamSynchronize(c->streams[s]));
    if (on && deferred) LQB_CUDA(cudaMemcpy(y, c->h_out[0].p, out_total, cudaMemcpyDeviceToHost));
    advance_all(c, n);
    return LQB_OK;
}

static lqb_chain_s *self_chain(lqb_stage_s *s)
{
    if (!s->self_chain) { s->self_chain = new lqb_chain_s(); s->self_chain->stages.push_back(s); }
    return s->self_chain;
}

template <class T> static T *as(lqb_stage s, Kind k) { return (s && s->kind == k) ? static_cast<T *>(s) : nullptr; }
#define LQB_GET(T, var, s, k) T *var = as<T>(s, k); if (!var) return fail(LQB_EINVAL, "%s: wrong or null stage handle", __func__)

}  // namespace lqb

using namespace lqb;

// ================================================================================ C ABI
extern "C" {

int lqb_version(void) { return 100; }
const char *lqb_last_error(void) { return g_err.c_str(); }
int lqb_device_count(int *count) { LQB_CUDA(cudaGetDeviceCount(count)); return LQB_OK; }
int lqb_set_device(int device) { LQB_CUDA(cudaSetDevice(device)); return LQB_OK; }
int lqb_device_synchronize(void) { LQB_CUDA(cudaDeviceSynchronize()); return LQB_OK; }
int lqb_host_alloc(void **p, size_t bytes) { LQB_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault)); return LQB_OK; }
int lqb_host_free(void *p) { LQB_CUDA(cudaFreeHost(p)); return LQB_OK; }
int lqb_dev_alloc(void **p, size_t bytes) { LQB_CUDA(cudaMalloc(p, bytes)); return LQB_OK; }
int lqb_dev_free(void *p) { LQB_CUDA(cudaFree(p)); return LQB_OK; }
int lqb_memcpy_h2d(void *d, const void *h, size_t b, void *st) { LQB_CUDA(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, (cudaStream_t)st)); return LQB_OK; }
int lqb_memcpy_d2h(void *h, const void *d, size_t b, void *st) { LQB_CUDA(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, (cudaStream_t)st)); return LQB_OK; }
int lqb_stream_synchronize(void *st) { LQB_CUDA(cudaStreamSynchronize((cudaStream_t)st)); return LQB_OK; }

// ---- generic
int lqb_stage_destroy(lqb_stage s) { if (s) { DevGuard g; if (s->ready) s->bind(); delete s; } return LQB_OK; }
int lqb_stage_reset(lqb_stage s) { if (!s) return fail(LQB_EINVAL, "null stage"); return s->reset(); }
int lqb_stage_channels(lqb_stage s, int *n) { if (!s) return fail(LQB_EINVAL, "null stage"); *n = s->C; return LQB_OK; }
int lqb_stage_out_len(lqb_stage s, size_t n, size_t *n_out) { if (!s) return fail(LQB_EINVAL, "null stage"); *n_out = s->out_len(n); return LQB_OK; }
int lqb_stage_execute(lqb_stage s, const void *x, size_t n, void *y, size_t cap, size_t *n_out)
{ if (!s) return fail(LQB_EINVAL, "null stage"); return chain_execute_host(self_chain(s), x, n, y, cap, n_out); }
int lqb_stage_execute_dev(lqb_stage s, const void *x, size_t n, void *y, size_t cap, size_t *n_out, void *stream)
{ if (!s) return fail(LQB_EINVAL, "null stage"); return chain_execute_dev(self_chain(s), x, n, y, cap, n_out, (cudaStream_t)stream); }

static int check_channels(int c) { return (c >= 1 && c <= (1 << 24)) ? LQB_OK : fail(LQB_EINVAL, "n_channels must be in [1, 2^24], got %d", c); }

// ---- iirfilt_crcf
int lqb_iirfilt_crcf_create_sos(const float *B, const float *A, int nsos, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!B || !A || nsos < 1 || nsos > 64 || !out) return fail(LQB_EINVAL, "iirfilt: need 1..64 sections");
    IirStage *q = new IirStage(C);
    int rc = q->init(std::vector<float>(B, B + 3 * nsos), std::vector<float>(A, A + 3 * nsos));
    if (rc != LQB_OK) { delete q; return rc; }
    *out = q; return LQB_OK;
}
int lqb_iirfilt_crcf_create_prototype(int ftype, int btype, int order, float fc, float f0, float ap, float as, int C, lqb_stage *out)
{
    std::vector<float> B, A;
    const int rc = design::iirdes_sos(ftype, btype, (unsigned)order, fc, f0, ap, as, B, A);
    if (rc != 0) return fail(LQB_EINVAL, "iirdes: invalid design parameters (order %d, fc %g, f0 %g, ap %g, as %g)", order, fc, f0, ap, as);
    return lqb_iirfilt_crcf_create_sos(B.data(), A.data(), (int)B.size() / 3, C, out);
}
int lqb_iirfilt_rrrf_create_sos(const float *B, const float *A, int nsos, int C, lqb_stage *out)
{
    LQB_TRY(lqb_iirfilt_crcf_create_sos(B, A, nsos, C, out));
    static_cast<IirStage *>(*out)->real_io = true;
    return LQB_OK;
}
int lqb_iirfilt_rrrf_create_prototype(int ftype, int btype, int order, float fc, float f0, float ap, float as, int C, lqb_stage *out)
{
    LQB_TRY(lqb_iirfilt_crcf_create_prototype(ftype, btype, order, fc, f0, ap, as, C, out));
    static_cast<IirStage *>(*out)->real_io = true;
    return LQB_OK;
}
int lqb_iirfilt_crcf_get_sos(lqb_stage s, float *B, float *A, int *nsos)
{
    LQB_GET(IirStage, q, s, K_IIR);
    if (B) std::copy(q->B.begin(), q->B.end(), B);
    if (A) std::copy(q->A.begin(), q->A.end(), A);
    if (nsos) *nsos = q->nsos;
    return LQB_OK;
}
int lqb_iirfilt_crcf_freqresponse(lqb_stage s, float fc, lqb_cf *H)
{
    LQB_GET(IirStage, q, s, K_IIR);
    const design::cplx h = design::sos_freqresponse(q->B, q->A, fc);
    H->re = h.real(); H->im = h.imag(); return LQB_OK;
}
int lqb_iirfilt_crcf_set_mode(lqb_stage s, int mode)
{
    LQB_GET(IirStage, q, s, K_IIR);
    if (mode < 0 || mode > 2) return fail(LQB_EINVAL, "iirfilt mode must be 0, 1 or 2");
    if (mode == 2 && q->nsos > kMaxSos) return fail(LQB_EINVAL, "blocked-scan IIR supports up to %d sections", kMaxSos);
    if (mode == 2 && q->real_io) return fail(LQB_ENOTIMPL, "blocked-scan IIR is built for complex samples only");
    q->mode = mode; return LQB_OK;
}

// ---- deemphasis
int lqb_deemph_create(float sr, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!(sr > 0.f) || !out) return fail(LQB_EINVAL, "deemph: sample_rate must be positive");
    DeemphStage *q = new DeemphStage(C);
    design::deemph_coeffs(sr, q->b0, q->a1);
    *out = q; return LQB_OK;
}
int lqb_deemph_get_coeffs(lqb_stage s, float *b0, float *a1) { LQB_GET(DeemphStage, q, s, K_DEEMPH); *b0 = q->b0; *a1 = q->a1; return LQB_OK; }
int lqb_deemph_freqresponse(lqb_stage s, float fc, lqb_cf *H)
{
    LQB_GET(DeemphStage, q, s, K_DEEMPH);
    const design::cplx e1 = std::polar(1.0f, (float)(-2 * design::kPi * fc));
    const design::cplx h = design::cplx(q->b0, 0.f) / (design::cplx(1.f, 0.f) + q->a1 * e1);
    H->re = h.real(); H->im = h.imag(); return LQB_OK;
}

// ---- firhilbf users
int lqb_firhilbf_create(int kind, int m, float as, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    std::vector<float> hq;
    if (!out || kind < LQB_FIRHILB_SSB_LSB || kind > LQB_FIRHILB_R2C) return fail(LQB_EINVAL, "firhilbf: unknown kind %d", kind);
    if (m < 2 || 4 * m > kFirMaxTaps || !design::firhilb_hq((unsigned)m, as, hq)) return fail(LQB_EINVAL, "firhilbf: m must be in 2..%d", kFirMaxTaps / 4);
    FirStage *q = new FirStage(C);
    q->hq = hq; q->delay = m;
    // tap index = delay in samples.  in-phase lane: a pure delay; quadrature lane: hq[i] meets the sample 2m-1-i pushes old
    if (kind == LQB_FIRHILB_SSB_LSB || kind == LQB_FIRHILB_SSB_USB) {
        // c2r_execute alternates between two window pairs: both lanes see every second sample
        q->h.assign((size_t)4 * m, 0.f); q->hlane_q.assign((size_t)4 * m, 0.f);
        q->h[(size_t)2 * m] = 1.f;
        for (int i = 0; i < 2 * m; i++) q->hlane_q[(size_t)(4 * m - 1 - 2 * i)] = hq[(size_t)i];
        q->mode = kind == LQB_FIRHILB_SSB_USB ? FIR_SSB_USB : FIR_SSB_LSB; q->out_r = true;
    } else if (kind == LQB_FIRHILB_R2C) {
        // decim_execute on overlapping pairs: quadrature filter on z[n], in-phase delay on z[n+1]
        q->h.assign((size_t)2 * m, 0.f); q->hlane_q.assign((size_t)2 * m, 0.f);
        q->h[(size_t)m - 1] = 1.f;
        for (int i = 0; i < 2 * m; i++) q->hlane_q[(size_t)(2 * m - 1 - i)] = hq[(size_t)i];
        q->mode = FIR_R2C; q->in_r = true; q->host_reset();
    } else {
        // interp_execute with overlapping writes: only the delayed imaginary part survives, sign alternating
        q->h.assign((size_t)m + 1, 0.f); q->hlane_q.assign((size_t)m + 1, 0.f);
        q->hlane_q[(size_t)m] = 1.f;
        q->mode = FIR_C2R; q->out_r = true;
    }
    *out = q; return LQB_OK;
}
int lqb_firhilbf_get_hq(lqb_stage s, float *hq, int *n)
{
    LQB_GET(FirStage, q, s, K_FIR);
    if (hq) std::copy(q->hq.begin(), q->hq.end(), hq);
    if (n) *n = (int)q->hq.size();
    return LQB_OK;
}

// ---- iirfilt, transfer-function form
static int tf_create(const float *b, int nb, const float *a, int na, int C, bool real_io, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!b || !a || !out || nb < 1 || na < 1 || nb > kMaxTf || na > kMaxTf)
        return fail(LQB_EINVAL, "iirfilt: need 1..%d numerator and denominator coefficients", kMaxTf);
    if (a[0] == 0.f) return fail(LQB_EINVAL, "iirfilt: a[0] must not be zero");
    TfStage *q = new TfStage(C);
    q->real_io = real_io;
    for (int i = 0; i < nb; i++) q->b.push_back(b[i] / a[0]);
    for (int i = 0; i < na; i++) q->a.push_back(a[i] / a[0]);
    *out = q; return LQB_OK;
}
int lqb_iirfilt_crcf_create(const float *b, int nb, const float *a, int na, int C, lqb_stage *out) { return tf_create(b, nb, a, na, C, false, out); }
int lqb_iirfilt_rrrf_create(const float *b, int nb, const float *a, int na, int C, lqb_stage *out) { return tf_create(b, nb, a, na, C, true, out); }
int lqb_iirfilt_tf_freqresponse(lqb_stage s, float fc, lqb_cf *H)
{
    LQB_GET(TfStage, q, s, K_TF);
    design::cplx hb(0.f, 0.f), ha(0.f, 0.f);
    for (size_t i = 0; i < q->b.size(); i++) hb += q->b[i] * design::cplx(cosf((float)(2 * design::kPi * fc * i)), -sinf((float)(2 * design::kPi * fc * i)));
    for (size_t i = 0; i < q->a.size(); i++) ha += q->a[i] * design::cplx(cosf((float)(2 * design::kPi * fc * i)), -sinf((float)(2 * design::kPi * fc * i)));
    const design::cplx h = hb / ha;
    H->re = h.real(); H->im = h.imag(); return LQB_OK;
}

// ---- firfilt_crcf
int lqb_firfilt_crcf_create(const float *h, int n, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!h || n < 1 || n > kFirMaxTaps || !out) return fail(LQB_EINVAL, "firfilt: need 1..%d taps", kFirMaxTaps);
    FirStage *q = new FirStage(C);
    q->h.assign(h, h + n);
    *out = q; return LQB_OK;
}
int lqb_firfilt_rrrf_create(const float *h, int n, int C, lqb_stage *out)
{
    LQB_TRY(lqb_firfilt_crcf_create(h, n, C, out));
    static_cast<FirStage *>(*out)->real_io = true;
    return LQB_OK;
}
int lqb_firfilt_rrrf_create_kaiser(int n, float fc, float as, float mu, int C, lqb_stage *out)
{
    std::vector<float> h;
    if (n < 1 || !design::firdes_kaiser((unsigned)n, fc, as, mu, h)) return fail(LQB_EINVAL, "firdes_kaiser: invalid parameters");
    return lqb_firfilt_rrrf_create(h.data(), n, C, out);
}
int lqb_firfilt_rrrf_create_dc_blocker(int m, float as, int C, lqb_stage *out)
{
    std::vector<float> h;
    if (m < 1 || !design::firdes_notch((unsigned)m, 0.0f, as, h)) return fail(LQB_EINVAL, "firdes_notch: invalid parameters");
    return lqb_firfilt_rrrf_create(h.data(), (int)h.size(), C, out);
}
int lqb_firfilt_crcf_create_kaiser(int n, float fc, float as, float mu, int C, lqb_stage *out)
{
    std::vector<float> h;
    if (n < 1 || !design::firdes_kaiser((unsigned)n, fc, as, mu, h)) return fail(LQB_EINVAL, "firdes_kaiser: invalid parameters");
    return lqb_firfilt_crcf_create(h.data(), n, C, out);
}
int lqb_firfilt_crcf_set_scale(lqb_stage s, float scale) { LQB_GET(FirStage, q, s, K_FIR); q->scale = scale; return LQB_OK; }
int lqb_firfilt_crcf_get_taps(lqb_stage s, float *h, int *n)
{
    LQB_GET(FirStage, q, s, K_FIR);
    if (h) std::copy(q->h.begin(), q->h.end(), h);
    if (n) *n = (int)q->h.size();
    return LQB_OK;
}
int lqb_firfilt_crcf_freqresponse(lqb_stage s, float fc, lqb_cf *H)
{
    LQB_GET(FirStage, q, s, K_FIR);
    const design::cplx h = design::fir_freqresponse(q->h, q->scale, fc);
    H->re = h.real(); H->im = h.imag(); return LQB_OK;
}

// ---- resamp
int lqb_resamp_create(float rate, int m, float fc, float as, int npfb, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!(rate > 0.f) || m < 1 || npfb < 1 || !out) return fail(LQB_EINVAL, "resamp: rate, m and npfb must be positive");
    if (!(rate >= 0.004f && rate <= 250.f)) return fail(LQB_EINVAL, "resamp: rate %g outside [0.004, 250]", rate);
    ResampStage *q = new ResampStage(C);
    if (!design::resamp_design((unsigned)m, fc, as, (unsigned)npfb, q->d)) { delete q; return fail(LQB_EINVAL, "resamp: invalid prototype (fc %g, As %g)", fc, as); }
    q->rate = rate; q->step = design::resamp_step(rate);
    *out = q; return LQB_OK;
}
static int resamp_variant(int variant, float rate, int m, float fc, float as, int npfb, int C, lqb_stage *out)
{
    LQB_TRY(lqb_resamp_create(rate, m, fc, as, npfb, C, out));
    static_cast<ResampStage *>(*out)->variant = variant;
    return LQB_OK;
}
int lqb_resamp_crcf_create(float rate, int m, float fc, float as, int npfb, int C, lqb_stage *out) { return resamp_variant(1, rate, m, fc, as, npfb, C, out); }
int lqb_resamp_rrrf_create(float rate, int m, float fc, float as, int npfb, int C, lqb_stage *out) { return resamp_variant(2, rate, m, fc, as, npfb, C, out); }
// resamp_*_create_default: m = 7, fc = min(0.49, rate / 2), As = 60 dB, 64 filters
int lqb_resamp_crcf_create_default(float rate, int C, lqb_stage *out)
{
    return resamp_variant(1, rate, 7, 0.5f * rate > 0.49f ? 0.49f : 0.5f * rate, 60.0f, 64, C, out);
}
int lqb_resamp_rrrf_create_default(float rate, int C, lqb_stage *out)
{
    return resamp_variant(2, rate, 7, 0.5f * rate > 0.49f ? 0.49f : 0.5f * rate, 60.0f, 64, C, out);
}
int lqb_wdelay_create(int delay, int real_samples, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out || delay < 0 || delay > (1 << 24)) return fail(LQB_EINVAL, "wdelay: delay must be in 0..2^24");
    DelayStage *q = new DelayStage(C);
    q->D = (long long)delay + 1; q->real_io = real_samples != 0;
    *out = q; return LQB_OK;
}
int lqb_resamp_set_rate(lqb_stage s, float rate)
{
    LQB_GET(ResampStage, q, s, K_RESAMP);
    if (!(rate >= 0.004f && rate <= 250.f)) return fail(LQB_EINVAL, "resamp: rate %g outside [0.004, 250]", rate);
    q->rate = rate; q->step = design::resamp_step(rate); return LQB_OK;
}
int lqb_resamp_get_state(lqb_stage s, uint32_t *step, uint32_t *phase) { LQB_GET(ResampStage, q, s, K_RESAMP); if (step) *step = q->step; if (phase) *phase = q->phase; return LQB_OK; }
int lqb_resamp_get_bank(lqb_stage s, float *bank, int *npfb, int *sublen)
{
    LQB_GET(ResampStage, q, s, K_RESAMP);
    if (bank) std::copy(q->d.bank.begin(), q->d.bank.end(), bank);
    if (npfb) *npfb = (int)q->d.npfb;
    if (sublen) *sublen = (int)q->d.sublen;
    return LQB_OK;
}

// ---- nco
int lqb_nco_create(int type, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out) return fail(LQB_EINVAL, "null out");
    NcoStage *q = new NcoStage(C);
    q->type = type == LQB_NCO ? 0 : 1;
    *out = q; return LQB_OK;
}
int lqb_nco_set_direction(lqb_stage s, int dir) { LQB_GET(NcoStage, q, s, K_NCO); if (dir != LQB_MIX_UP && dir != LQB_MIX_DOWN) return fail(LQB_EINVAL, "bad direction"); q->dir = dir; return LQB_OK; }
int lqb_nco_set_frequency(lqb_stage s, float f)     { LQB_GET(NcoStage, q, s, K_NCO); return q->rmw(q->dtheta, design::nco_constrain(f), true); }
int lqb_nco_adjust_frequency(lqb_stage s, float df) { LQB_GET(NcoStage, q, s, K_NCO); return q->rmw(q->dtheta, design::nco_constrain(df), false); }
int lqb_nco_set_phase(lqb_stage s, float phi)       { LQB_GET(NcoStage, q, s, K_NCO); return q->rmw(q->theta, design::nco_constrain(phi), true); }
int lqb_nco_adjust_phase(lqb_stage s, float dphi)   { LQB_GET(NcoStage, q, s, K_NCO); return q->rmw(q->theta, design::nco_constrain(dphi), false); }
int lqb_nco_get_frequency(lqb_stage s, float *f)
{ LQB_GET(NcoStage, q, s, K_NCO); LQB_TRY(q->ensure()); uint32_t d; LQB_CUDA(cudaDeviceSynchronize()); LQB_TRY(q->dtheta.download(&d, 1)); *f = design::nco_u32_to_frequency(d); return LQB_OK; }
int lqb_nco_get_phase(lqb_stage s, float *phi)
{ LQB_GET(NcoStage, q, s, K_NCO); LQB_TRY(q->ensure()); uint32_t t; LQB_CUDA(cudaDeviceSynchronize()); LQB_TRY(q->theta.download(&t, 1)); *phi = design::nco_u32_to_phase(t); return LQB_OK; }
int lqb_nco_pll_set_bandwidth(lqb_stage s, float bw) { LQB_GET(NcoStage, q, s, K_NCO); if (bw < 0.f) return fail(LQB_EINVAL, "pll bandwidth must be >= 0"); q->alpha = bw; q->beta = std::sqrt(bw); return LQB_OK; }
int lqb_nco_pll_step(lqb_stage s, float dphi)
{
    LQB_GET(NcoStage, q, s, K_NCO);
    LQB_TRY(q->rmw(q->dtheta, design::nco_constrain(dphi * q->alpha), false));
    return q->rmw(q->theta, design::nco_constrain(dphi * q->beta), false);
}
int lqb_nco_set_frequency_per_channel(lqb_stage s, const float *f, int n)
{
    LQB_GET(NcoStage, q, s, K_NCO);
    if (!f || n != q->C) return fail(LQB_EINVAL, "need one frequency per channel");
    LQB_TRY(q->ensure());
    std::vector<uint32_t> d(n); for (int i = 0; i < n; i++) d[i] = design::nco_constrain(f[i]);
    LQB_CUDA(cudaDeviceSynchronize());
    return q->dtheta.upload(d.data(), n);
}
int lqb_nco_get_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n)
{
    LQB_GET(NcoStage, q, s, K_NCO);
    if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count");
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    if (theta) LQB_TRY(q->theta.download(theta, n));
    if (d_theta) LQB_TRY(q->dtheta.download(d_theta, n));
    return LQB_OK;
}
int lqb_nco_set_u32(lqb_stage s, const uint32_t *theta, const uint32_t *d_theta, int n)
{
    LQB_GET(NcoStage, q, s, K_NCO);
    if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count");
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    if (theta) LQB_TRY(q->theta.upload(theta, n));
    if (d_theta) LQB_TRY(q->dtheta.upload(d_theta, n));
    return LQB_OK;
}

// ---- agc
int lqb_agc_create(int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out) return fail(LQB_EINVAL, "null out");
    *out = new AgcStage(C); return LQB_OK;
}
int lqb_agc_set_bandwidth(lqb_stage s, float bw) { LQB_GET(AgcStage, q, s, K_AGC); if (bw < 0.f || bw > 1.f) return fail(LQB_EINVAL, "agc bandwidth must be in [0, 1]"); q->alpha = bw; return LQB_OK; }
int lqb_agc_get_bandwidth(lqb_stage s, float *bw) { LQB_GET(AgcStage, q, s, K_AGC); *bw = q->alpha; return LQB_OK; }
static int agc_set_gain_all(AgcStage *q, float g, bool reset_y2)
{
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    LQB_TRY(q->g.fill(g));
    return reset_y2 ? q->y2p.fill(1.0f) : LQB_OK;
}
static int agc_gain0(AgcStage *q, float *g) { LQB_TRY(q->ensure()); LQB_CUDA(cudaDeviceSynchronize()); return q->g.download(g, 1); }
int lqb_agc_set_signal_level(lqb_stage s, float x2) { LQB_GET(AgcStage, q, s, K_AGC); if (!(x2 > 0.f)) return fail(LQB_EINVAL, "agc level must be > 0"); return agc_set_gain_all(q, 1.0f / x2, true); }
int lqb_agc_get_signal_level(lqb_stage s, float *x2) { LQB_GET(AgcStage, q, s, K_AGC); float g; LQB_TRY(agc_gain0(q, &g)); *x2 = 1.0f / g; return LQB_OK; }
int lqb_agc_set_rssi(lqb_stage s, float rssi)
{
    LQB_GET(AgcStage, q, s, K_AGC);
    float g = std::pow(10.0f, -rssi / 20.0f); if (g < 1e-16f) g = 1e-16f;
    return agc_set_gain_all(q, g, true);
}
int lqb_agc_get_rssi(lqb_stage s, float *rssi) { LQB_GET(AgcStage, q, s, K_AGC); float g; LQB_TRY(agc_gain0(q, &g)); *rssi = (float)(-20 * std::log10((double)g)); return LQB_OK; }
int lqb_agc_set_gain(lqb_stage s, float g) { LQB_GET(AgcStage, q, s, K_AGC); if (!(g > 0.f)) return fail(LQB_EINVAL, "agc gain must be > 0"); return agc_set_gain_all(q, g, false); }
int lqb_agc_get_gain(lqb_stage s, float *g) { LQB_GET(AgcStage, q, s, K_AGC); return agc_gain0(q, g); }
int lqb_agc_get_gain_per_channel(lqb_stage s, float *g, int n)
{ LQB_GET(AgcStage, q, s, K_AGC); if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count"); LQB_TRY(q->ensure()); LQB_CUDA(cudaDeviceSynchronize()); return q->g.download(g, n); }
int lqb_agc_set_scale(lqb_stage s, float sc) { LQB_GET(AgcStage, q, s, K_AGC); if (!(sc > 0.f)) return fail(LQB_EINVAL, "agc scale must be > 0"); q->scale = sc; return LQB_OK; }
int lqb_agc_get_scale(lqb_stage s, float *sc) { LQB_GET(AgcStage, q, s, K_AGC); *sc = q->scale; return LQB_OK; }
int lqb_agc_set_precision(lqb_stage s, int mode)
{
    LQB_GET(AgcStage, q, s, K_AGC);
    if (mode != LQB_AGC_AUTO && mode != LQB_AGC_EXACT && mode != LQB_AGC_FAST) return fail(LQB_EINVAL, "agc precision must be LQB_AGC_AUTO, _EXACT or _FAST");
    q->precision = mode; return LQB_OK;
}
int lqb_agc_get_precision(lqb_stage s, int *mode) { LQB_GET(AgcStage, q, s, K_AGC); *mode = q->precision; return LQB_OK; }
int lqb_agc_lock(lqb_stage s, int locked) { LQB_GET(AgcStage, q, s, K_AGC); q->locked = locked ? 1 : 0; return LQB_OK; }
int lqb_agc_squelch_enable(lqb_stage s, int en)
{
    LQB_GET(AgcStage, q, s, K_AGC);
    q->squelch = en != 0;
    if (!q->ready) return LQB_OK;          // materialize() starts from the squelch flag
    LQB_TRY(q->bind());
    LQB_CUDA(cudaDeviceSynchronize());
    return q->mode.fill(en ? 1 : 7);
}
int lqb_agc_squelch_set_threshold(lqb_stage s, float t) { LQB_GET(AgcStage, q, s, K_AGC); q->threshold = t; return LQB_OK; }
int lqb_agc_squelch_get_threshold(lqb_stage s, float *t) { LQB_GET(AgcStage, q, s, K_AGC); *t = q->threshold; return LQB_OK; }
int lqb_agc_squelch_set_timeout(lqb_stage s, unsigned t) { LQB_GET(AgcStage, q, s, K_AGC); q->timeout = t; return LQB_OK; }   /* takes effect at the next FALL, as in liquid */
int lqb_agc_squelch_get_status(lqb_stage s, int *st)
{
    LQB_GET(AgcStage, q, s, K_AGC);
    if (!q->ready) { *st = q->squelch ? 1 : 7; return LQB_OK; }
    LQB_TRY(q->bind()); LQB_CUDA(cudaDeviceSynchronize()); return q->mode.download(st, 1);
}
int lqb_agc_take_rise_count(lqb_stage s, unsigned *count)
{
    LQB_GET(AgcStage, q, s, K_AGC);
    if (!q->ready) { *count = 0; return LQB_OK; }
    LQB_TRY(q->bind());
    LQB_CUDA(cudaDeviceSynchronize());
    LQB_TRY(q->rise.download(count, 1));
    return q->rise.zero();
}

// ---- ampmodem
int lqb_ampmodem_create(float mod, int type, int suppressed, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out || !(mod > 0.f)) return fail(LQB_EINVAL, "ampmodem: modulation index must be > 0");
    if (type != LQB_AMPMODEM_DSB && type != LQB_AMPMODEM_USB && type != LQB_AMPMODEM_LSB) return fail(LQB_EINVAL, "ampmodem: unknown type %d", type);
    AmStage *q = new AmStage(C);
    q->mod = mod; q->type = type; q->suppressed = suppressed ? 1 : 0;
    if (type != LQB_AMPMODEM_DSB) {
        // firhilbf_create(m = 25, 60 dB) as in ampmodem_create; with carrier the DC blocker follows the Hilbert pair
        lqb_stage h = nullptr, d = nullptr;
        int rc = lqb_firhilbf_create(type == LQB_AMPMODEM_USB ? LQB_FIRHILB_SSB_USB : LQB_FIRHILB_SSB_LSB, kAmDelay, 60.0f, C, &h);
        if (rc == LQB_OK && !suppressed) rc = lqb_firfilt_rrrf_create_dc_blocker(kAmDelay, 20.0f, C, &d);
        if (rc != LQB_OK) { delete static_cast<FirStage *>(h); delete q; return rc; }
        q->hil = static_cast<FirStage *>(h); q->dcb = static_cast<FirStage *>(d);
        q->hil->post_div = mod;
    }
    // constants of liquid's ampmodem_create: m = 25, lowpass kaiser(2m+1, 0.01, 40 dB), dc blocker (25, 20 dB)
    design::firdes_kaiser(kAmTaps, 0.01f, 40.0f, 0.0f, q->lp);
    design::firdes_notch(kAmDelay, 0.0f, 20.0f, q->dc);
    *out = q; return LQB_OK;
}
// ---- FMStereo
int lqb_fmstereo_create(float iq_rate, float pcm_rate, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out || !(iq_rate > 0.f) || !(pcm_rate > 0.f)) return fail(LQB_EINVAL, "FMStereo: rates must be positive");
    const float rate = pcm_rate / iq_rate;
    if (!(rate >= 0.004f && rate <= 250.f)) return fail(LQB_EINVAL, "FMStereo: pcm_rate / iq_rate %g outside [0.004, 250] (resamp_rrrf_create_default)", rate);
    FmstStage *q = new FmstStage(C);
    // demod.hpp:20-32: 75 us de-emphasis at the I/Q rate, freqdem(4.0), resamp_rrrf_create_default(pcm_rate / iq_rate)
    const float a1 = (float)(-std::exp(-1.0 / (75.0E-6 * (double)iq_rate)));
    q->a1 = a1; q->b0 = (float)(1.0 + (double)a1);
    if (!design::resamp_design(7, 0.5f * rate > 0.49f ? 0.49f : 0.5f * rate, 60.0f, 64, q->d) || q->d.sublen > (unsigned)kFmstMaxSub) {
        delete q; return fail(LQB_EINVAL, "FMStereo: resampler design failed");
    }
    q->step = design::resamp_step(rate);
    *out = q; return LQB_OK;
}
int lqb_fmstereo_get_state(lqb_stage s, uint32_t *theta, uint32_t *d_theta, float *pe, int n)
{
    LQB_GET(FmstStage, q, s, K_FMST);
    if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count");
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    if (theta) LQB_TRY(q->theta.download(theta, n));
    if (d_theta) LQB_TRY(q->dtheta.download(d_theta, n));
    if (pe) LQB_TRY(q->pe.download(pe, n));
    return LQB_OK;
}
int lqb_fmstereo_get_deemph(lqb_stage s, float *b0, float *a1) { LQB_GET(FmstStage, q, s, K_FMST); *b0 = q->b0; *a1 = q->a1; return LQB_OK; }

// ---- BroadcastAM
int lqb_broadcast_am_create(int m, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out || m < 1 || m > kBamMaxM) return fail(LQB_EINVAL, "BroadcastAM: slen must be in 1..%d", kBamMaxM);
    BamStage *q = new BamStage(C);
    q->m = m; q->ntaps_pad = (2 * m + 1 + 7) / 8 * 8;
    // demod.hpp:101-106: PLL bandwidth 0.001, lowpass kaiser(2m+1, 0.01, 40 dB), cheby2 order-3 high-pass at 20 Hz / 48 kHz
    design::firdes_kaiser((unsigned)(2 * m + 1), 0.01f, 40.0f, 0.0f, q->lp);
    if (design::iirdes_sos(design::CHEBY2, design::HIGHPASS, 3, 20.0f / 48000.0f, 0.0f, 0.5f, 20.0f, q->B, q->A) != 0 || q->B.size() != 6) {
        delete q; return fail(LQB_EINVAL, "BroadcastAM: DC-block design failed");
    }
    *out = q; return LQB_OK;
}
int lqb_broadcast_am_get_design(lqb_stage s, float *lp, int *nlp, float *B, float *A)
{
    LQB_GET(BamStage, q, s, K_BAM);
    if (lp) std::copy(q->lp.begin(), q->lp.end(), lp);
    if (nlp) *nlp = (int)q->lp.size();
    if (B) std::copy(q->B.begin(), q->B.end(), B);
    if (A) std::copy(q->A.begin(), q->A.end(), A);
    return LQB_OK;
}
int lqb_broadcast_am_get_nco_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n)
{
    LQB_GET(BamStage, q, s, K_BAM);
    if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count");
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    if (theta) LQB_TRY(q->theta.download(theta, n));
    if (d_theta) LQB_TRY(q->dtheta.download(d_theta, n));
    return LQB_OK;
}
int lqb_ampmodem_get_taps(lqb_stage s, float *lp, int *nlp, float *dc, int *ndc)
{
    LQB_GET(AmStage, q, s, K_AM);
    if (lp) std::copy(q->lp.begin(), q->lp.end(), lp);
    if (dc) std::copy(q->dc.begin(), q->dc.end(), dc);
    if (nlp) *nlp = (int)q->lp.size();
    if (ndc) *ndc = (int)q->dc.size();
    return LQB_OK;
}
int lqb_ampmodem_get_nco_u32(lqb_stage s, uint32_t *theta, uint32_t *d_theta, int n)
{
    LQB_GET(AmStage, q, s, K_AM);
    if (n < 0 || n > q->C) return fail(LQB_EINVAL, "bad channel count");
    LQB_TRY(q->ensure());
    LQB_CUDA(cudaDeviceSynchronize());
    if (theta) LQB_TRY(q->theta.download(theta, n));
    if (d_theta) LQB_TRY(q->dtheta.download(d_theta, n));
    return LQB_OK;
}

// ---- freqdem
int lqb_freqdem_create(float kf, int C, lqb_stage *out)
{
    LQB_TRY(check_channels(C));
    if (!out || !(kf > 0.f)) return fail(LQB_EINVAL, "freqdem: kf must be > 0");
    FmStage *q = new FmStage(C);
    q->kf = kf; q->ref = (float)(1.0f / (2 * design::kPi * kf));
    *out = q; return LQB_OK;
}

// ---- chain
int lqb_chain_create(lqb_chain *out) { if (!out) return fail(LQB_EINVAL, "null out"); *out = new lqb_chain_s(); return LQB_OK; }
int lqb_chain_append(lqb_chain c, lqb_stage s) { if (!c || !s) return fail(LQB_EINVAL, "null handle"); c->stages.push_back(s); return LQB_OK; }
int lqb_chain_destroy(lqb_chain c) { if (c) { DevGuard g; if (!c->stages.empty()) { std::lock_guard<std::mutex> lk(g_live_mu); if (g_live.count(c->stages[0])) cudaSetDevice(c->stages[0]->device); } delete c; } return LQB_OK; }
int lqb_chain_out_len(lqb_chain c, size_t n, size_t *n_out) { if (!c || !n_out) return fail(LQB_EINVAL, "null handle"); return chain_out_len(c, n, n_out); }
int lqb_chain_execute(lqb_chain c, const void *x, size_t n, void *y, size_t cap, size_t *n_out) { if (!c) return fail(LQB_EINVAL, "null chain"); return chain_execute_host(c, x, n, y, cap, n_out); }
int lqb_chain_execute_dev(lqb_chain c, const void *x, size_t n, void *y, size_t cap, size_t *n_out, void *stream)
{ if (!c) return fail(LQB_EINVAL, "null chain"); return chain_execute_dev(c, x, n, y, cap, n_out, (cudaStream_t)stream); }
int lqb_chain_execute_i16(lqb_chain c, const int16_t *iq, size_t n, void *y, size_t cap, size_t *n_out)
{ if (!c) return fail(LQB_EINVAL, "null chain"); return chain_execute_host(c, iq, n, y, cap, n_out, true); }
int lqb_chain_execute_i16_dev(lqb_chain c, const int16_t *iq_dev, size_t n, void *y_dev, size_t cap, size_t *n_out, void *stream)
{ if (!c) return fail(LQB_EINVAL, "null chain"); return chain_execute_dev(c, iq_dev, n, y_dev, cap, n_out, (cudaStream_t)stream, true); }
int lqb_bytes_to_iq(const int16_t *iq, size_t n, lqb_cf *out)
{
    if (!iq || !out) return fail(LQB_EINVAL, "null argument");
    if (n == 0) return LQB_OK;
    DevArr<char> in, o;
    LQB_TRY(in.alloc(n * 4)); LQB_TRY(o.alloc(n * 8));
    LQB_CUDA(cudaMemcpy(in.p, iq, n * 4, cudaMemcpyHostToDevice));
    LQB_CUDA(i16_to_c64_launch(in.p, (float2 *)o.p, (long long)n, nullptr));
    LQB_CUDA(cudaMemcpy(out, o.p, n * 8, cudaMemcpyDeviceToHost));
    return LQB_OK;
}
int lqb_chain_plan(lqb_chain c, char *buf, size_t len)
{
    if (!c || !buf || !len) return fail(LQB_EINVAL, "null argument");
    LQB_TRY(chain_validate(c));
    std::vector<Segment> segs; LQB_TRY(build_plan(c, segs));
    snprintf(buf, len, "%s", c->plan.c_str());
    return LQB_OK;
}
int lqb_chain_last_kernels(lqb_chain c, char *buf, size_t len)
{
    if (!c || !buf || !len) return fail(LQB_EINVAL, "null argument");
    snprintf(buf, len, "%s", c->last_kernels.c_str());
    return LQB_OK;
}
int lqb_chain_clear_error(lqb_chain c) { if (!c) return fail(LQB_EINVAL, "null chain"); c->poisoned = false; c->poison_why.clear(); return LQB_OK; }
int lqb_chain_last_launches(lqb_chain c, int *n) { if (!c || !n) return fail(LQB_EINVAL, "null argument"); *n = c->last_launches; return LQB_OK; }
int lqb_chain_set_timing(lqb_chain c, int enabled)
{
    if (!c) return fail(LQB_EINVAL, "null chain");
    c->clear_timing(); c->timing = enabled != 0;
    return LQB_OK;
}
int lqb_chain_get_timing(lqb_chain c, float *ms, int cap, int *n_segments, int *n_calls)
{
    if (!c || !ms || !n_segments || !n_calls) return fail(LQB_EINVAL, "null argument");
    *n_calls = (int)c->timed_calls.size();
    *n_segments = c->timed_calls.empty() ? 0 : (int)c->timed_calls[0].size();
    for (int k = 0; k < cap; k++) ms[k] = 0.f;
    for (auto &call : c->timed_calls)
        for (size_t k = 0; k < call.size() && (int)k < cap; k++) {
            LQB_CUDA(cudaEventSynchronize(call[k].second));
            float t = 0.f; LQB_CUDA(cudaEventElapsedTime(&t, call[k].first, call[k].second));
            ms[k] += t;
        }
    return LQB_OK;
}
int lqb_chain_set_overlap(lqb_chain c, int enabled)
{
    if (!c) return fail(LQB_EINVAL, "null chain");
    if (c->tail_stream) LQB_CUDA(cudaStreamSynchronize(c->tail_stream));
    c->tail_pending[0] = c->tail_pending[1] = false;
    c->overlap = enabled != 0;
    return LQB_OK;
}
int lqb_chain_wait(lqb_chain c, void *stream)
{
    if (!c) return fail(LQB_EINVAL, "null chain");
    return chain_join_tails(c, (cudaStream_t)stream);
}
int lqb_stream_create(void **stream, int high_priority)
{
    if (!stream) return fail(LQB_EINVAL, "null argument");
    int lo = 0, hi = 0;
    LQB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    cudaStream_t s; LQB_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
    *stream = (void *)s; return LQB_OK;
}
int lqb_stream_destroy(void *stream) { if (stream) LQB_CUDA(cudaStreamDestroy((cudaStream_t)stream)); return LQB_OK; }
int lqb_chain_set_fusion(lqb_chain c, int level) { if (!c || level < 0 || level > 2) return fail(LQB_EINVAL, "fusion level must be 0, 1 or 2"); c->fuse = level; return LQB_OK; }

// ---- synthetic inputs
int lqb_synth_fill(int kind, void *x, int C, int ch0, size_t n, uint64_t n0, uint64_t seed, void *stream)
{
    if (kind < 0 || kind > 3 || !x) return fail(LQB_EINVAL, "synth: kind must be 0..3");
    LQB_CUDA(synth_launch(kind, (float2 *)x, C, ch0, (long long)n, n0, seed, (cudaStream_t)stream));
    return LQB_OK;
}

}  // extern "C"
