#!/usr/bin/env python
"""bench.py -- the AM-receiver chain of BASELINE.json (config 5) on N B200s, one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference            # the CPU path on the box's host cores

A step is one pass of the hot path over one block: `channels` channels x 65536 complex64 samples at
2 MS/s through ComplexIIRFilter(cheby2-8) -> ComplexResampler(0.024) -> AGC -> AmpModem(dsb, carrier)
-> DeemphasisFilter, state carried from block to block.  Channels are independent, so each rank owns
its own `channels` (weak scaling, no collective on the data path; torch.distributed only carries the
barrier and the max-over-ranks of the device time).

value   : whole-job input Msamples/s with the block already resident in HBM (CUDA events on the
          launching stream, max over ranks).  Each block is >= 4 GB, far larger than the 126 MB L2.
e2e     : the same metric through the C ABI's host-pointer entry (lqb_chain_execute): pinned host
          input, H2D, kernels, D2H of the audio, all inside the timed region.
roofline: the full-rate kernel seq[iir4+resamp] timed by its own events; algorithmic bytes per input
          sample 8 (c64 in) + 0.024*8 (c64 out) = 8.192 (DESIGN.md "Roofline").
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
    return "port"      # oracle/_ref cannot exist: the reference's arithmetic is in liquid-dsp, absent here


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


def side_config(args, L, torch, dev, rank, world, barrier, max_over_ranks):
    """Configs 2-4 of BASELINE.json: device-resident throughput of the plan the C ABI builds, same timing rules."""
    import numpy as np
    stream = torch.cuda.current_stream().cuda_stream
    if args.config == 2:
        C, n, kind, bps = 1024, 1 << 20, 1, 16.0
        from oracle import oracle as O
        chain = L.Chain(L.FIRFilter(O.firdes_kaiser(64, 0.1, 60.0), channels=C))
        name, out_real = "config2: FIRFilter 64-tap crcf, 1024 channels x 1M samples", False
    elif args.config == 3:
        C, n, kind, bps = 4096, 65536, 2, 8.192
        nco = L.NCO(channels=C); nco.set_frequencies((2 * np.pi * (0.05 + 0.4 * np.arange(C) / 4096)).astype(np.float32)); nco.set_direction(True)
        chain = L.Chain(nco, L.ComplexResampler(0.024, Fc=0.024, channels=C))
        name, out_real = "config3: NCO mix-down + ComplexResampler 2e6->48e3, 4096 channels x 64K blocks", False
    elif args.config == 4:
        C, n, kind, bps = int(os.environ.get("LQB_BENCH_C4_CHANNELS", "16384")), 65536, 3, 12.0      # (override: occupancy experiments)
        chain = L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C), L.AGC(channels=C), L.FreqDem(0.1, channels=C))
        name, out_real = "config4: ComplexIIRFilter cheby2-8 + AGC + FreqDem, 16384 channels x 64K blocks", True
    in_real = False
    if args.next:                                       # the classes either side of the hot path
        out_real = True
        if args.next == "bam":
            C, n, kind, bps = 65536, 65536, 0, 8.096
            iir, rs, agc, _, de = build_radio(L, C)
            chain = L.Chain(iir, rs, agc, L.BroadcastAM(25, channels=C), de)
            name = "8f-3: bandpass + resampler + AGC + BroadcastAM + de-emphasis, 65536 channels x 64K blocks"
        elif args.next == "ssb":
            C, n, kind, bps = 1024, 1 << 20, 1, 12.0
            chain = L.Chain(L.SSBDemod("usb", channels=C)); name = "8f-3: SSBDemod (firhilbf 25, 60 dB), 1024 channels x 1M samples"
        elif args.next == "fmstereo":
            C, n, kind, bps = 16384, 65536, 3, 8.64
            chain = L.Chain(L.FMStereo(channels=C)); name = "8f-4: FMStereo 600 kHz -> 48 kHz stereo, 16384 channels x 64K blocks"
        elif args.next == "rrrf":
            C, n, kind, bps, in_real = 65536, 65536, 1, 8.0, True
            chain = L.Chain(L.RealIIRFilter("cheby2", "lowpass", 8, 0.05, channels=C)); name = "8f-1: RealIIRFilter cheby2-8, 65536 channels x 64K real samples"
        elif args.next == "cresamp":
            C, n, kind, bps, out_real = 4096, 65536, 1, 8.64, False
            chain = L.Chain(L.CResampler(0.08, channels=C)); name = "8f-4: CResampler(0.08), 4096 channels x 64K blocks"
        elif args.next == "rfir":
            C, n, kind, bps, in_real = 2048, 1 << 20, 1, 8.0, True
            chain = L.Chain(L.RealKaiserBessel(63, 0.1, 60.0, channels=C)); name = "8f-1: RealKaiserBessel 63 taps, 2048 channels x 1M real samples"
        elif args.next == "resamp":
            C, n, kind, bps, out_real = 65536, 65536, 1, 8.192, False
            chain = L.Chain(L.ComplexResampler(0.024, Fc=0.024, channels=C)); name = "8a: ComplexResampler(0.024) alone, 65536 channels x 64K blocks"
        elif args.next in ("iir", "nco", "agc", "fm", "deemph"):        # the path's stages one at a time
            C, n, kind = 65536, 16384, 1
            if args.next == "iir":
                chain = L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C)); bps, out_real = 16.0, False
            elif args.next == "nco":
                o = L.NCO(channels=C); o.set_frequencies((0.3 + 1e-5 * np.arange(C)).astype(np.float32)); o.set_direction(True)
                chain = L.Chain(o); bps, out_real = 16.0, False
            elif args.next == "agc":
                agc = L.AGC(channels=C); agc.precision = args.agc_precision
                chain = L.Chain(agc); bps, out_real = 16.0, False
            elif args.next == "fm":
                chain = L.Chain(L.FreqDem(0.1, channels=C)); bps, out_real, kind = 12.0, True, 3
            else:
                chain = L.Chain(L.DeemphasisFilter(48000, channels=C)); bps, out_real, in_real = 8.0, True, True
            name = "8a: %s alone, 65536 channels x 16384 samples" % args.next
            if args.next == "agc":
                name += " (precision %s)" % args.agc_precision
        else:
            raise SystemExit("unknown --next row")
    if args.block != BLOCK:
        n = args.block                      # profiling runs use a shorter block
    if in_real:
        x = torch.randn((C, n), dtype=torch.float32, device=dev)
    else:
        x = torch.empty((C, n), dtype=torch.complex64, device=dev)
    cap = chain.out_len(n) + 2
    y = torch.empty((C, cap), dtype=torch.float32 if out_real else torch.complex64, device=dev)
    if not in_real:
        L.synth_fill(kind, x.data_ptr(), C, n, channel0=rank * C, stream=stream)
    # inputs smaller than L2 (config 3: 2.1 GB, fine; all are > 126 MB) -- every config streams > L2 per step
    for _ in range(max(args.warmup, 3)):
        chain.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chain.set_timing(True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        chain.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream); launches += chain.last_launches()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    seg, calls = chain.segment_ms()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = peaks.get("hbm_gbs", 6650.0)
    val = world * C * n / (ms * 1e-3) / 1e6
    if rank == 0:
        print(json.dumps({"metric": METRIC.replace("AM-chain", args.next or "config %d" % args.config), "value": val, "unit": UNIT, "n_gpus": world,
                          "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": name, "plan": chain.plan(), "l2": "%.1f GB streamed per step, larger than L2" % (C * n * bps / 1e9)},
                          "gpu_launches": launches,
                          "roofline": {"bound": "hbm", "achieved": C * n * bps / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": C * n * bps / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_sample": bps,
                                       "segments_ms": [t / max(calls, 1) for t in seg], "traffic": None}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=65536, help="channels per GPU (config 5: 65536)")
    ap.add_argument("--e2e-channels", type=int, default=8192, help="channels per GPU of the host-buffer (e2e) leg")
    ap.add_argument("--fuse", type=int, default=1, help="chain fusion level (0, 1, 2)")
    ap.add_argument("--block", type=int, default=BLOCK, help="samples per channel per step (profiling runs use a shorter block)")
    ap.add_argument("--next", default="", help="side line: bam, ssb, fmstereo, rrrf, cresamp, rfir (SURVEY 8f); iir, nco, agc, fm, deemph (8a stages alone)")
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 5 (default, the headline AM receiver), 2 FIR, 3 NCO+resampler, 4 IIR+AGC+FM")
    ap.add_argument("--agc-precision", default="auto", choices=["auto", "exact", "fast"],
                    help="--next agc: gain-loop arithmetic of the stage alone (auto = exact for a stage on its own)")
    ap.add_argument("--overlap", type=int, default=0,
                    help="config 5: 1 = block k's decimated-rate tail overlaps block k+1's front (lqb_chain_set_overlap), 0 = serial calls")
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-worker", nargs=2, metavar=("SECONDS", "SEED"))
    args = ap.parse_args()

    if args.cpu_worker:
        cpu_worker(float(args.cpu_worker[0]), int(args.cpu_worker[1])); return

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "config5: README AMRadio chain, %d channels/GPU x %d-sample blocks @ 2 MS/s, state carried" % (args.channels, args.block)

    if args.impl == "reference":
        # the reference's CPU implementation of the path, restated (liquid-dsp is absent): all host cores,
        # one channel per process; a step is a bounded sample of the same workload
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        from oracle import oracle as O  # noqa: F401  (builds liboracle.so if missing)
        for _ in range(min(args.warmup, 1)):
            run_cpu_pool(1.0, cores)
        secs = max(1.0, min(6.0, 60.0 / max(1, args.steps)))
        vals = [run_cpu_pool(secs, cores) for _ in range(args.steps)]
        v = sum(vals) / len(vals)
        sample = "%d processes x 1 channel of config 1 (README AMRadio, 64K blocks), %.1f s per step" % (cores, secs)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import liquiddsp as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the product path)")
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

    # CPU baseline first (rank 0, N = 1): separate processes, before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        v = run_cpu_pool(args.cpu_seconds, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
               "sample": "%d processes x 1 channel of config 1 (README AMRadio, 64K blocks) for %.0f s each; CPU restatement of liquid-dsp, not liquid-dsp" % (cores, args.cpu_seconds)}

    if args.config != 5 or args.next:
        side_config(args, L, torch, dev, rank, world, barrier, max_over_ranks)
        if world > 1:
            dist.destroy_process_group()
        return

    C, n = args.channels, args.block
    if args.overlap:
        # the launching stream outranks the chain's tail stream, so a block's front is dispatched ahead of the tail queued before it
        torch.cuda.set_stream(torch.cuda.Stream(device=dev, priority=-1))
    stream = torch.cuda.current_stream().cuda_stream
    x = torch.empty((C, n), dtype=torch.complex64, device=dev)
    stages = build_radio(L, C)
    whole = L.Chain(*stages, fuse=args.fuse)
    whole.set_overlap(bool(args.overlap))
    n_mid = stages[1].out_len(n)
    cap = n_mid + 2
    y = torch.empty((C, cap), dtype=torch.float32, device=dev)
    L.synth_fill(0, x.data_ptr(), C, n, channel0=rank * C, n0=0, seed=0xB200, stream=stream)
    torch.cuda.synchronize()

    def step():
        whole.execute_dev(x.data_ptr(), n, y.data_ptr(), cap, stream)
        return whole.last_launches()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    whole.set_timing(True)          # one CUDA-event pair per plan segment per call, on this stream (C ABI)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for k in range(args.steps):
        launches += step()
    whole.wait(stream)               # overlapped calls: the timed region ends when the last block's tail has finished
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = max_over_ranks(e0.elapsed_time(e1))
    seg_ms, seg_calls = whole.segment_ms()
    whole.set_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    ms_step = ms / args.steps
    value = world * C * n / (ms_step * 1e-3) / 1e6

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    roof = None
    plan = whole.plan()
    if plan.startswith("seq[iir4+resamp]") and seg_calls:
        # dominant kernel = first plan segment (the full-rate kernel), timed by its own events inside the timed region
        kms = seg_ms[0] / seg_calls
        bytes_launch = C * n * 8 + C * n_mid * 8
        ach = bytes_launch / (kms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("bytes_per_launch")
        except (OSError, ValueError):
            pass
        kname = "front2_kernel<4>" if C >= 32768 and os.environ.get("LQB_FRONT2", "1") != "0" else "seq_kernel<F_IIR|F_RS, 4, TMA>"
        roof = {"bound": "hbm", "kernel": kname + " (seq[iir4+resamp])", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "peak_source": peak_src, "traffic": traffic, "kernel_ms": kms,
                "algorithmic_bytes_per_launch": bytes_launch, "share_of_step": kms / ms_step,
                "segments_ms": [t / seg_calls for t in seg_ms],
                "chain_frac": (C * n * 8 + C * n_mid * 4) / (ms_step * 1e-3) / 1e9 / peak}

    # end to end through the host-pointer C ABI: pinned host input -> H2D -> kernels -> D2H audio
    e2e = None
    if not args.no_e2e:
        Ce = min(args.e2e_channels, C)
        rs_e = build_radio(L, Ce)
        ch_e = L.Chain(*rs_e, fuse=args.fuse)
        xh = torch.empty((Ce, n), dtype=torch.complex64).pin_memory()
        xh.copy_(x[:Ce].cpu())
        xn = xh.numpy()
        for _ in range(2):
            yh = ch_e(xn)
        barrier()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            yh = ch_e(xn)
        L.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0) / reps
        e2e = {"value": world * Ce * n / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(xn.nbytes), "d2h_bytes_per_step": int(yh.nbytes),
               "channels_per_gpu": Ce, "ms_per_step": dt * 1e3, "api": "liquiddsp.Chain.__call__ -> lqb_chain_execute (host pointers)"}
        # the same call fed the SDR wire format (interleaved int16 I/Q, bytes_to_iq fused into the first kernel)
        ih = torch.empty((Ce, 2 * n), dtype=torch.int16).pin_memory()
        ih.copy_((torch.view_as_real(xh).reshape(Ce, 2 * n) * 32767.0).clamp(-32767, 32767).to(torch.int16))
        inp = ih.numpy()
        for _ in range(2):
            yi = ch_e(inp)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            yi = ch_e(inp)
        L.synchronize()
        dti = max_over_ranks(time.perf_counter() - t0) / reps
        e2e["int16_iq"] = {"value": world * Ce * n / dti / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(inp.nbytes),
                           "d2h_bytes_per_step": int(yi.nbytes), "ms_per_step": dti * 1e3,
                           "api": "liquiddsp.Chain.__call__(int16 I/Q) -> lqb_chain_execute_i16"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": {"workload": workload, "channels_per_gpu": C, "block": n, "fuse": args.fuse, "overlap": int(bool(args.overlap)),
                                               "plan": plan,
                                               "l2": "each block is %.1f GB of input, larger than L2; no flush needed" % (C * n * 8 / 1e9)},
               "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "e2e": e2e}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
