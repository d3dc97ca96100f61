#!/usr/bin/env python
"""bench.py -- the AM-receiver chain of BASELINE.json (config 5) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU path on the box's host cores

A step is one pass of the hot path over one block: 65536 channels x 65536 complex64 samples at 2 MS/s through
ComplexIIRFilter(cheby2-8) -> ComplexResampler(0.024) -> AGC -> AmpModem(dsb, carrier) -> DeemphasisFilter, state
carried from block to block.  Channels are independent: the 65536 channels are SHARDED over the N ranks (BASELINE
config 5, SURVEY 8e: 65536 / N per GPU -- strong scaling, the default), no collective on the data path
(torch.distributed only carries the barrier and the max-over-ranks of the device time).  For N > 1 the same line also
carries the weak-scaling figure (65536 channels on EVERY GPU) under "weak"; `--scaling weak` makes that the headline.

value   : whole-job input Msamples/s with the block already resident in HBM (CUDA events on the
          launching stream, max over ranks).  Each block is >= 4 GB, far larger than the 126 MB L2.
e2e     : the same metric through the C ABI's host-pointer entry (lqb_chain_execute) on the SAME channel share: pinned
          host input, H2D, kernels, D2H of the audio, all inside the timed region; plus the int16 wire format and a
          pageable numpy input (what an SDR callback hands over).
roofline: the full-rate kernel (first plan segment) timed by its own events, named by the library
          (lqb_chain_last_kernels); algorithmic bytes per input sample 8 (c64 in) + 0.024*8 (c64 out) = 8.192.
config1 : BASELINE config 1 on the GPU -- ONE channel, 64K blocks from a pageable numpy array, object by object (the
          README's five calls) and as one Chain call, beside one CPU core of the restatement.
side    : configs 2, 3, 4 of BASELINE.json at their stated sizes (device-resident, same timing rules).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "python-liquiddsp_b200"))

BLOCK = 65536
FS, PCM = 2.0e6, 48.0e3
METRIC = "AM-chain aggregate input Msamples/s"
UNIT = "Msamples/s"


# ------------------------------------------------------------------------------------- CPU arm
def cpu_worker(seconds, seed):
    """One channel of config 1 (README AMRadio, 64K blocks, state carried) on one core; prints samples/s."""
    import numpy as np
    from oracle import oracle as O
    try:
        os.sched_setaffinity(0, {seed % (os.cpu_count() or 1)})
    except (AttributeError, OSError):
        pass
    rng = np.random.default_rng(0xB200 + seed)
    t = np.arange(BLOCK * 8) / FS
    a = 0.6 * np.sin(2 * np.pi * 1000 * t) + 0.4 * np.sin(2 * np.pi * 2500 * t)
    x = 0.1 * (1 + 0.5 * a) * np.exp(1j * (2 * np.pi * (200 + 10 * (seed % 32)) * t)) + 0.05 * np.exp(2j * np.pi * 60e3 * t)
    x = (x + 0.02 / np.sqrt(2) * (rng.standard_normal(t.size) + 1j * rng.standard_normal(t.size))).astype(np.complex64)
    radio = O.AMRadio(15000, FS, PCM)
    for b in range(2):
        radio(x[b * BLOCK:(b + 1) * BLOCK])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for b in range(8):
            radio(x[b * BLOCK:(b + 1) * BLOCK]); n += BLOCK
    dt = time.perf_counter() - t0
    print(json.dumps({"samples": n, "seconds": dt}))


def run_cpu_pool(seconds, procs):
    """`procs` worker processes, one channel each; returns aggregate Msamples/s."""
    ps = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", str(seconds), str(i)],
                           stdout=subprocess.PIPE, text=True) for i in range(procs)]
    tot = 0.0
    for p in ps:
        out, _ = p.communicate()
        r = json.loads(out.strip().splitlines()[-1])
        tot += r["samples"] / r["seconds"]
    return tot / 1e6


def cpu_kind():
    """'reference' when a build of the real reference is present (driver-provisioned baseline/_ref or oracle/_ref with an
    importable liquiddsp extension), else 'port': the C restatement under oracle/ (liquid-dsp is absent from this image)."""
    for d in (os.path.join(ROOT, "baseline", "_ref"), os.path.join(ROOT, "oracle", "_ref")):
        if os.path.isdir(d) and any(f.startswith("liquiddsp") and f.endswith(".so") for f in os.listdir(d)):
            return "reference"
    return "port"


WORKLOAD = "config5: README AMRadio chain, 65536 channels x 65536-sample blocks @ 2 MS/s sharded by channel over the GPUs, state carried"


def common_config(channels_total, block):
    """The workload both arms are quoted on (the product arm adds its plan under "gpu", the reference arm its sample)."""
    return {"workload": WORKLOAD if (channels_total, block) == (65536, BLOCK) else
            "config5 variant: README AMRadio chain, %d channels x %d-sample blocks @ 2 MS/s sharded by channel, state carried" % (channels_total, block),
            "channels_total": channels_total, "block": block}


# ------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML poll every 5 ms on a thread (the timed region of the
    default run is ~100 ms, shorter than nvidia-smi's start-up); nvidia-smi -lms as the fallback when NVML is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.p, self.nv, self.run = [], None, None, True
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()
            return
        except Exception:
            self.nv = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except OSError:
            self.p = None

    def _poll(self):
        nv = self.nv
        bits = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        while self.run:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.perf_counter(), [str(sm), str(self.mx), ""] + ["Active" if mask & b else "Not Active" for _, b in bits]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.p is None and self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        if self.nv is not None:
            self.run = False; self.t.join(timeout=1.0)
        else:
            time.sleep(0.15); self.p.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1] or [r for (_, r) in self.rows[-3:]]
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvml" if self.nv is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------- GPU arm
def build_radio(L, channels):
    iir = L.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / FS, channels=channels)
    rs = L.ComplexResampler(rate=PCM / FS, Fc=PCM / FS, channels=channels)
    agc = L.AGC(channels=channels); agc.lock = False; agc.scale = 0.01
    am = L.AmpModem(modulation=0.5, type="dsb", carrier=True, channels=channels)
    de = L.DeemphasisFilter(PCM, channels=channels)
    return iir, rs, agc, am, de


def pin_to_gpu_numa(local):
    """Bind this rank's process (and so its page-locked buffers, first-touched after this) to the CPUs NVML reports as local
    to its GPU.  N ranks streaming 54 GB/s each from host memory otherwise all sit on whatever NUMA node the launcher
    left them on (round 1: CPU affinity 0-31 / NUMA 0 for all eight) and the far socket's GPUs pull every byte across
    the inter-socket link.  Returns a small record for the JSON line; never fails the run."""
    rec = {"cpus_before": None, "cpus": None, "source": None}
    try:
        rec["cpus_before"] = len(os.sched_getaffinity(0))
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        phys = int(vis.split(",")[local]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else local
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(range(os.cpu_count() or 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
            rec.update(cpus=len(cpus), first_cpu=min(cpus), last_cpu=max(cpus), source="nvmlDeviceGetCpuAffinity")
        try:
            node_mask = pynvml.nvmlDeviceGetMemoryAffinity(h, 4, pynvml.NVML_AFFINITY_SCOPE_NODE)
            rec["numa_nodes"] = [64 * w + b for w, m in enumerate(node_mask) for b in range(64) if (m >> b) & 1]
        except Exception:
            pass
    except Exception as e:          # no NVML, no permission: run unpinned and say so
        rec["error"] = repr(e)[:120]
    return rec


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "hbm_gbs" in peaks:
            return peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def side_setup(which, L, np, nxt="", agc_precision="auto"):
    """Chains of BASELINE configs 2-4 and the SURVEY 8(a)/(f) rows one at a time.
    Returns (chain, C, n, synth kind, algorithmic bytes per input sample, in_real, out_real, name)."""
    in_real = False
    if which == 2:
        C, n, kind, bps = 1024, 1 << 20, 1, 16.0
        fir = L.FIRFilter.kaiser(64, 0.1, 60.0, channels=C)               # the product's own firdes_kaiser
        chain = L.Chain(fir)
        name, out_real = "config2: FIRFilter 64-tap crcf, 1024 channels x 1M samples", False
    elif which == 3:
        C, n, kind, bps = 4096, 65536, 2, 8.192
        nco = L.NCO(channels=C); nco.set_frequencies((2 * np.pi * (0.05 + 0.4 * np.arange(C) / 4096)).astype(np.float32)); nco.set_direction(True)
        chain = L.Chain(nco, L.ComplexResampler(0.024, Fc=0.024, channels=C))
        name, out_real = "config3: NCO mix-down + ComplexResampler 2e6->48e3, 4096 channels x 64K blocks", False
    elif which == 4:
        C, n, kind, bps = int(os.environ.get("LQB_BENCH_C4_CHANNELS", "16384")), 65536, 3, 12.0      # (override: occupancy experiments)
        chain = L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C), L.AGC(channels=C), L.FreqDem(0.1, channels=C))
        name, out_real = "config4: ComplexIIRFilter cheby2-8 + AGC + FreqDem, %d channels x 64K blocks" % C, True
    else:
        out_real = True
        if nxt == "bam":
            C, n, kind, bps = 65536, 65536, 0, 8.096
            iir, rs, agc, _, de = build_radio(L, C)
            chain = L.Chain(iir, rs, agc, L.BroadcastAM(25, channels=C), de)
            name = "8f-3: bandpass + resampler + AGC + BroadcastAM + de-emphasis, 65536 channels x 64K blocks"
        elif nxt == "ssb":
            C, n, kind, bps = 1024, 1 << 20, 1, 12.0
            chain = L.Chain(L.SSBDemod("usb", channels=C)); name = "8f-3: SSBDemod (firhilbf 25, 60 dB), 1024 channels x 1M samples"
        elif nxt == "fmstereo":
            C, n, kind, bps = 16384, 65536, 3, 8.64
            chain = L.Chain(L.FMStereo(channels=C)); name = "8f-4: FMStereo 600 kHz -> 48 kHz stereo, 16384 channels x 64K blocks"
        elif nxt == "rrrf":
            C, n, kind, bps, in_real = 65536, 65536, 1, 8.0, True
            chain = L.Chain(L.RealIIRFilter("cheby2", "lowpass", 8, 0.05, channels=C)); name = "8f-1: RealIIRFilter cheby2-8, 65536 channels x 64K real samples"
        elif nxt == "cresamp":
            C, n, kind, bps, out_real = 4096, 65536, 1, 8.64, False
            chain = L.Chain(L.CResampler(0.08, channels=C)); name = "8f-4: CResampler(0.08), 4096 channels x 64K blocks"
        elif nxt == "rfir":
            C, n, kind, bps, in_real = 2048, 1 << 20, 1, 8.0, True
            chain = L.Chain(L.RealKaiserBessel(63, 0.1, 60.0, channels=C)); name = "8f-1: RealKaiserBessel 63 taps, 2048 channels x 1M real samples"
        elif nxt == "resamp":
            C, n, kind, bps, out_real = 65536, 65536, 1, 8.192, False
            chain = L.Chain(L.ComplexResampler(0.024, Fc=0.024, channels=C)); name = "8a: ComplexResampler(0.024) alone, 65536 channels x 64K blocks"
        elif nxt in ("iir", "nco", "agc", "fm", "deemph"):        # the path's stages one at a time
            C, n, kind = 65536, 16384, 1
            if nxt == "iir":
                chain = L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C)); bps, out_real = 16.0, False
            elif nxt == "nco":
                o = L.NCO(channels=C); o.set_frequencies((0.3 + 1e-5 * np.arange(C)).astype(np.float32)); o.set_direction(True)
                chain = L.Chain(o); bps, out_real = 16.0, False
            elif nxt == "agc":
                agc = L.AGC(channels=C); agc.precision = agc_precision
                chain = L.Chain(agc); bps, out_real = 16.0, False
            elif nxt == "fm":
                chain = L.Chain(L.FreqDem(0.1, channels=C)); bps, out_real, kind = 12.0, True, 3
            else:
                chain = L.Chain(L.DeemphasisFilter(48000, channels=C)); bps, out_real, in_real = 8.0, True, True
            name = "8a: %s alone, 65536 channels x 16384 samples" % nxt
            if nxt == "agc":
                name += " (precision %s)" % agc_precision
        else:
            raise SystemExit("unknown --next row")
    return chain, C, n, kind, bps, in_real, out_real, name


def side_measure(which, L, torch, np, dev, rank, steps, warmup, barrier, max_over_ranks, nxt="", agc_precision="auto", block=None):
    """Device-resident throughput of one side workload on this rank's GPU: warm-up, then `steps` calls timed with CUDA
    events on the launching stream (inputs far larger than L2); returns the record."""
    stream = torch.cuda.current_stream().cuda_stream
    chain, C, n, kind, bps, in_real, out_real, name = side_setup(which, L, np, nxt, agc_precision)
    if block:
        n = block                           # profiling runs use a shorter block
    x = torch.randn((C, n), dtype=torch.float32, device=dev) if in_real else torch.empty((C, n), dtype=torch.complex64, device=dev)
    cap = chain.out_len(n) + 2
    y = torch.empty((C, cap), dtype=torch.float32 if out_real else torch.complex64, device=dev)
    if not in_real:
        L.synth_fill(kind, x.data_ptr(), C, n, channel0=rank * C, stream=stream)
    for _ in range(max(warmup, 3)):
        chain.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chain.set_timing(True)
    launches = 0
    e0.record()
    for _ in range(steps):
        chain.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream); launches += chain.last_launches()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    seg, calls = chain.segment_ms()
    chain.set_timing(False)
    peak, _ = hbm_peak()
    rec = {"workload": name, "value": C * n / (ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms, "plan": chain.plan(), "kernels": chain.last_kernels(),
           "gpu_launches": launches, "algorithmic_bytes_per_sample": bps, "achieved_gbs": C * n * bps / (ms * 1e-3) / 1e9,
           "frac": C * n * bps / (ms * 1e-3) / 1e9 / peak, "segments_ms": [t / max(calls, 1) for t in seg],
           "l2": "%.1f GB streamed per step, larger than L2" % (C * n * bps / 1e9)}
    del x, y, chain
    torch.cuda.empty_cache()
    return rec


def measure_chain(L, torch, dev, rank, C, n, steps, warmup, fuse, overlap, barrier, max_over_ranks, sample_clocks=False, local=0):
    """The headline measurement on this rank's share: C channels x n samples per step, HBM-resident, state carried."""
    if overlap:
        # the launching stream outranks the chain's tail stream, so a block's front is dispatched ahead of the tail queued before it
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    stream = torch.cuda.current_stream().cuda_stream
    x = torch.empty((C, n), dtype=torch.complex64, device=dev)
    stages = build_radio(L, C)
    whole = L.Chain(*stages, fuse=fuse)
    whole.set_overlap(bool(overlap))
    n_mid = stages[1].out_len(n)
    cap = n_mid + 2
    y = torch.empty((C, cap), dtype=torch.float32, device=dev)
    L.synth_fill(0, x.data_ptr(), C, n, channel0=rank * C, n0=0, seed=0xB200, stream=stream)
    torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        whole.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream)
    barrier()
    sampler = ClockSampler(local) if sample_clocks else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    whole.set_timing(True)          # one CUDA-event pair per plan segment per call, on the stream the segment runs on (C ABI)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        whole.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream); launches += whole.last_launches()
    whole.wait(stream)               # overlapped calls: the timed region ends when the last block's tail has finished
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / steps
    seg_ms, seg_calls = whole.segment_ms()
    whole.set_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    return {"ms_step": ms_step, "launches": launches, "seg_ms": seg_ms, "seg_calls": seg_calls, "clocks": clocks, "plan": whole.plan(),
            "kernels": whole.last_kernels(), "n_mid": n_mid, "x": x, "chain": whole, "stages": stages}


def e2e_measure(L, torch, np, x_dev, Ce, n, fuse, reps, barrier, max_over_ranks, world):
    """The public API with HOST buffers: liquiddsp.Chain.__call__ -> lqb_chain_execute; H2D of the block and D2H of the
    audio inside the timed region.  Pinned c64, pinned int16 wire format, and a pageable numpy array."""
    out = {}
    ch_e = L.Chain(*build_radio(L, Ce), fuse=fuse)
    xh = torch.empty((Ce, n), dtype=torch.complex64).pin_memory()
    xh.copy_(x_dev[:Ce])
    xn = xh.numpy()

    def timed(call, arg, reps):
        for _ in range(2):
            y = call(arg)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            y = call(arg)
        L.synchronize()
        own = (time.perf_counter() - t0) / reps
        return max_over_ranks(own * reps) / reps, y, own

    dt, yh, own_pinned = timed(ch_e, xn, reps)
    out = {"value": world * Ce * n / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(xn.nbytes), "d2h_bytes_per_step": int(yh.nbytes),
           "channels_per_gpu": Ce, "ms_per_step": dt * 1e3, "host_memory": "pinned (cudaHostAlloc)",
           "h2d_gbs_per_gpu": xn.nbytes / dt / 1e9, "own_h2d_gbs": xn.nbytes / own_pinned / 1e9,
           "api": "liquiddsp.Chain.__call__ -> lqb_chain_execute (host pointers)"}
    # the same call fed the SDR wire format (interleaved int16 I/Q, bytes_to_iq fused into the first kernel)
    ih = torch.empty((Ce, 2 * n), dtype=torch.int16).pin_memory()
    ih.copy_((torch.view_as_real(xh).reshape(Ce, 2 * n) * 32767.0).clamp(-32767, 32767).to(torch.int16))
    inp = ih.numpy()
    dti, yi, _ = timed(ch_e, inp, reps)
    out["int16_iq"] = {"value": world * Ce * n / dti / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(inp.nbytes),
                       "d2h_bytes_per_step": int(yi.nbytes), "ms_per_step": dti * 1e3,
                       "api": "liquiddsp.Chain.__call__(int16 I/Q) -> lqb_chain_execute_i16"}
    del ih, inp
    # pageable numpy input (what an SDR callback hands over, README.md:60-63): a bounded share so the leg stays short
    Cp = min(Ce, 8192)
    xp = np.empty((Cp, n), np.complex64); xp[...] = xn[:Cp]
    ch_p = L.Chain(*build_radio(L, Cp), fuse=fuse)
    dtp, yp, own_page = timed(ch_p, xp, max(2, reps // 2))
    out["pageable"] = {"value": world * Cp * n / dtp / 1e6, "unit": UNIT, "channels_per_gpu": Cp, "h2d_bytes_per_step": int(xp.nbytes),
                       "d2h_bytes_per_step": int(yp.nbytes), "ms_per_step": dtp * 1e3, "host_memory": "pageable numpy array",
                       "h2d_gbs_per_gpu": xp.nbytes / dtp / 1e9, "own_h2d_gbs": xp.nbytes / own_page / 1e9}
    return out


def config1_record(cpu_one_core):
    """BASELINE config 1 through the drop-in classes (tools/config1_bench.py): one channel, pageable numpy blocks."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import config1_bench
    r = config1_bench.run(nblocks=48, warm=4)
    r["cpu_one_core_msamples_per_s"] = cpu_one_core
    r["note"] = "README AMRadio on ONE channel: `objects` = the README's five calls per block, `chain` = liquiddsp.Chain; CPU = one core of the restatement"
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=65536, help="channels of the whole job (config 5: 65536), sharded over the GPUs under strong scaling")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE config 5): --channels in total, channels/N per GPU; weak: --channels on every GPU")
    ap.add_argument("--e2e-channels", type=int, default=0, help="channels per GPU of the host-buffer (e2e) leg; 0 = the headline share")
    ap.add_argument("--fuse", type=int, default=1, help="chain fusion level (0, 1, 2)")
    ap.add_argument("--block", type=int, default=BLOCK, help="samples per channel per step (profiling runs use a shorter block)")
    ap.add_argument("--next", default="", help="side line: bam, ssb, fmstereo, rrrf, cresamp, rfir (SURVEY 8f); iir, nco, agc, fm, deemph (8a stages alone)")
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 5 (default, the headline AM receiver), 2 FIR, 3 NCO+resampler, 4 IIR+AGC+FM")
    ap.add_argument("--agc-precision", default="auto", choices=["auto", "exact", "fast"],
                    help="--next agc: gain-loop arithmetic of the stage alone (auto = exact for a stage on its own)")
    ap.add_argument("--overlap", default="0",
                    help="config 5: 1 = block k's decimated-rate tail overlaps block k+1's front (lqb_chain_set_overlap), 0 = serial calls "
                         "(measured slower both with a full machine and with 8192 channels per GPU: DESIGN 4.9)")
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the config 1 / 2 / 3 / 4 records")
    ap.add_argument("--cpu-worker", nargs=2, metavar=("SECONDS", "SEED"))
    args = ap.parse_args()

    if args.cpu_worker:
        cpu_worker(float(args.cpu_worker[0]), int(args.cpu_worker[1])); return

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = common_config(args.channels, args.block)

    if args.impl == "reference":
        # the reference's CPU implementation of the path, restated (liquid-dsp is absent): all host cores, one channel of
        # the workload per process (channels are independent and run the same chain); a step is a bounded sample
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        from oracle import oracle as O  # noqa: F401  (builds liboracle.so if missing)
        for _ in range(min(args.warmup, 1)):
            run_cpu_pool(1.0, cores)
        secs = max(1.0, min(6.0, 60.0 / max(1, args.steps)))
        vals = [run_cpu_pool(secs, cores) for _ in range(args.steps)]
        v = sum(vals) / len(vals)
        sample = ("%d of the workload's channels at a time, one per core (%d processes x 1 channel, 64K-sample blocks, state carried), "
                  "%.1f s per step; CPU restatement of liquid-dsp (oracle/), not liquid-dsp" % (cores, cores, secs))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import liquiddsp as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
    numa = pin_to_gpu_numa(local) if os.environ.get("LQB_BENCH_NO_PIN") is None else {"source": "disabled (LQB_BENCH_NO_PIN)"}
    torch.cuda.set_device(local); L.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # CPU baseline first (rank 0): separate processes, before the GPU is busy
    cpu = None
    if rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v = run_cpu_pool(args.cpu_seconds, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
               "sample": "%d processes x 1 channel of the workload (README AMRadio, 64K blocks) for %.0f s each; CPU restatement of liquid-dsp, not liquid-dsp" % (cores, args.cpu_seconds)}
    barrier()

    peak, peak_src = hbm_peak()
    if args.config != 5 or args.next:
        rec = side_measure(args.config, L, torch, np, dev, rank, args.steps, args.warmup, barrier, max_over_ranks, args.next, args.agc_precision,
                           args.block if args.block != BLOCK else None)
        if rank == 0:
            print(json.dumps({"metric": METRIC.replace("AM-chain", args.next or "config %d" % args.config), "value": world * rec["value"], "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
                              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": rec["workload"], "plan": rec["plan"], "kernels": rec["kernels"], "l2": rec["l2"]},
                              "gpu_launches": rec["gpu_launches"],
                              "roofline": {"bound": "hbm", "achieved": rec["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": rec["frac"],
                                           "algorithmic_bytes_per_sample": rec["algorithmic_bytes_per_sample"], "segments_ms": rec["segments_ms"], "traffic": None}}))
        if world > 1:
            dist.destroy_process_group()
        return

    n = args.block
    C = args.channels // world if args.scaling == "strong" else args.channels          # this rank's share
    if C < 1:
        raise SystemExit("bench.py: fewer channels than GPUs")
    ov = bool(int(args.overlap))
    m = measure_chain(L, torch, dev, rank, C, n, args.steps, args.warmup, args.fuse, ov, barrier, max_over_ranks, sample_clocks=(rank == 0), local=local)
    ms_step = m["ms_step"]
    value = world * C * n / (ms_step * 1e-3) / 1e6

    roof = None
    if m["seg_calls"] and "resamp]" in m["plan"].split(" -> ")[0]:
        # dominant kernel = first plan segment (the full-rate kernel), timed by its own events inside the timed region
        kms = m["seg_ms"][0] / m["seg_calls"]
        bytes_launch = C * n * 8 + C * m["n_mid"] * 8
        ach = bytes_launch / (kms * 1e-3) / 1e9
        kname = [k for k in m["kernels"] if k not in ("tapstream_kernel",)][0]
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            ent = tr.get("kernels", {}).get(kname)
            if ent and ent.get("channels") == C and ent.get("block") == n:      # an ncu capture of this kernel at this launch shape
                traffic = ent["bytes_per_launch"]
        except (OSError, ValueError):
            pass
        roof = {"bound": "hbm", "kernel": "%s (%s)" % (kname, m["plan"].split(" -> ")[0]), "kernels_per_step": m["kernels"], "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src, "traffic": traffic, "kernel_ms": kms,
                "algorithmic_bytes_per_launch": bytes_launch, "share_of_step": kms / ms_step,
                "segments_ms": [t / m["seg_calls"] for t in m["seg_ms"]],
                "chain_frac": (C * n * 8 + C * m["n_mid"] * 4) / (ms_step * 1e-3) / 1e9 / peak}

    # end to end through the host-pointer C ABI on the same share: host input -> H2D -> kernels -> D2H audio
    e2e = None
    if not args.no_e2e:
        Ce = min(args.e2e_channels or C, C)
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:
            avail = 1 << 62
        need = Ce * n * 8 * 1.6 * (world if world > 1 else 1)        # c64 pinned + int16 pinned + slack, all ranks on one host
        while Ce > 1024 and need > 0.6 * avail:
            Ce //= 2; need /= 2
        e2e = e2e_measure(L, torch, np, m["x"], Ce, n, args.fuse, max(2, min(args.steps, 4)), barrier, max_over_ranks, world)
        if Ce != C:
            e2e["note"] = "host memory bounds the e2e share to %d channels per GPU (headline share %d)" % (Ce, C)
        # every rank's own host -> device rate (the slowest one sets `value`): what tells a saturated host resource from a
        # badly placed rank
        mine = torch.tensor([e2e["own_h2d_gbs"], e2e["pageable"]["own_h2d_gbs"], float(numa.get("first_cpu", -1))], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, mine)
        else:
            allr = [mine]
        e2e["per_rank"] = [{"rank": i, "pinned_h2d_gbs": round(float(t[0]), 2), "pageable_h2d_gbs": round(float(t[1]), 2), "first_cpu": int(t[2])} for i, t in enumerate(allr)]
        e2e["host_binding"] = numa
    x_keep = None
    del m["x"], m["chain"], m["stages"]
    torch.cuda.empty_cache()

    # weak-scaling figure beside the strong one (N > 1): 65536 channels on every GPU
    weak = None
    if world > 1 and args.scaling == "strong":
        w = measure_chain(L, torch, dev, rank, args.channels, n, max(3, args.steps // 2), args.warmup, args.fuse, False, barrier, max_over_ranks)
        weak = {"value": world * args.channels * n / (w["ms_step"] * 1e-3) / 1e6, "unit": UNIT, "channels_per_gpu": args.channels,
                "ms_per_step": w["ms_step"], "segments_ms": [t / max(w["seg_calls"], 1) for t in w["seg_ms"]], "kernels": w["kernels"]}
        del w
        torch.cuda.empty_cache()

    # the other BASELINE configs, under the same clock (rank 0's GPU; the other ranks wait at the barrier)
    side, c1 = None, None
    if not args.no_side:
        solo = lambda: torch.cuda.synchronize()
        ident = lambda v: v
        if rank == 0:
            side = {}
            for k in (2, 3, 4):
                try:
                    side["config%d" % k] = side_measure(k, L, torch, np, dev, 0, max(3, args.steps // 2), args.warmup, solo, ident)
                except Exception as e:      # a side record must not take the headline down
                    side["config%d" % k] = {"error": repr(e)[:300]}
            try:
                c1 = config1_record(cpu["value"] / cpu["cores"] if cpu else None)
            except Exception as e:
                c1 = {"error": repr(e)[:300]}
        barrier()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": cfg,
               "gpu": {"channels_per_gpu": C, "fuse": args.fuse, "overlap": int(ov), "plan": m["plan"], "kernels": m["kernels"],
                       "l2": "each block is %.1f GB of input per GPU, larger than L2; no flush needed" % (C * n * 8 / 1e9)},
               "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
               "weak": weak, "config1": c1, "side": side}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
