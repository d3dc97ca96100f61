"""Diagnostics run on the GPU box (not a test): per-stage error of the AM chain against the oracle, and
host<->device copy bandwidth through the C ABI.  Usage: python tools/diag_gpu.py [chain] [copy]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-liquiddsp_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import liquiddsp as L  # noqa: E402
from oracle import oracle as O  # noqa: E402
from util import am_iq, rel_l2  # noqa: E402


def chain():
    n, blk = 6 * 65536, 65536
    x = am_iq(n)
    g = [L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075), L.ComplexResampler(0.024, Fc=0.024), L.AGC(), L.AmpModem(0.5, "dsb", True), L.DeemphasisFilter(48000)]
    g2 = [L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075), L.ComplexResampler(0.024, Fc=0.024), L.AGC(), L.AmpModem(0.5, "dsb", True), L.DeemphasisFilter(48000)]
    o = [O.ComplexIIRFilter(_sos=g[0].sos()), O.ComplexResampler(0.024, Fc=0.024), O.AGC(), O.AmpModem(0.5, "dsb", True), O.DeemphasisFilter(48000)]
    for q in (g[2], g2[2], o[2]):
        q.scale = 0.01
    names = ["iir", "resamp", "agc", "ampmodem", "deemph"]
    cum = [[] for _ in names]; iso = [[] for _ in names]; ref = [[] for _ in names]
    for i in range(0, n, blk):
        a = b = x[i:i + blk]
        for k in range(5):
            r = o[k](b)                 # oracle fed by oracle
            cum[k].append(g[k](a))      # GPU fed by GPU
            iso[k].append(g2[k](b))     # GPU fed by oracle
            ref[k].append(r)
            a, b = cum[k][-1], r
    for k, nm in enumerate(names):
        r = np.concatenate(ref[k])
        print("%-9s cumulative %.3e   isolated %.3e   (n=%d)" % (nm, rel_l2(np.concatenate(cum[k]), r), rel_l2(np.concatenate(iso[k]), r), len(r)))
    print("agc gain gpu/oracle - 1 = %.3e" % (g[2].gain / o[2].gain - 1))
    print("am nco gpu", g[3].nco_u32(), "oracle", o[3].nco_u32())


def copy():
    import torch
    nbytes = 1 << 30
    xh = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    xd = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        xd.copy_(xh, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        xd.copy_(xh, non_blocking=True)
    torch.cuda.synchronize(); print("torch pinned H2D      %.1f GB/s" % (4 * nbytes / (time.perf_counter() - t0) / 1e9))
    import ctypes as C
    lib = L._lib
    for _ in range(2):
        lib.lqb_memcpy_h2d(C.c_void_p(xd.data_ptr()), C.c_void_p(xh.data_ptr()), nbytes, None)
    L.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        lib.lqb_memcpy_h2d(C.c_void_p(xd.data_ptr()), C.c_void_p(xh.data_ptr()), nbytes, None)
    L.synchronize(); print("lqb H2D, torch-pinned  %.1f GB/s" % (4 * nbytes / (time.perf_counter() - t0) / 1e9))
    p = C.c_void_p(); lib.lqb_host_alloc(C.byref(p), nbytes)
    for _ in range(2):
        lib.lqb_memcpy_h2d(C.c_void_p(xd.data_ptr()), p, nbytes, None)
    L.synchronize(); t0 = time.perf_counter()
    for _ in range(4):
        lib.lqb_memcpy_h2d(C.c_void_p(xd.data_ptr()), p, nbytes, None)
    L.synchronize(); print("lqb H2D, lqb-pinned    %.1f GB/s" % (4 * nbytes / (time.perf_counter() - t0) / 1e9))
    t0 = time.perf_counter()
    for _ in range(4):
        lib.lqb_memcpy_d2h(p, C.c_void_p(xd.data_ptr()), nbytes, None)
    L.synchronize(); print("lqb D2H, lqb-pinned    %.1f GB/s" % (4 * nbytes / (time.perf_counter() - t0) / 1e9))
    xp = np.empty(nbytes, np.uint8); xp[:] = 1
    t0 = time.perf_counter()
    lib.lqb_memcpy_h2d(C.c_void_p(xd.data_ptr()), xp.ctypes.data_as(C.c_void_p), nbytes, None)
    L.synchronize(); print("lqb H2D, pageable      %.1f GB/s" % (nbytes / (time.perf_counter() - t0) / 1e9))
    # chain host path: time per phase
    Cn, n = 8192, 65536
    st = (L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=Cn), L.ComplexResampler(0.024, Fc=0.024, channels=Cn), L.AGC(channels=Cn),
          L.AmpModem(0.5, "dsb", True, channels=Cn), L.DeemphasisFilter(48000, channels=Cn))
    ch = L.Chain(*st)
    xh = torch.zeros((Cn, n), dtype=torch.complex64).pin_memory(); xn = xh.numpy()
    for _ in range(2):
        y = ch(xn)
    t0 = time.perf_counter(); y = ch(xn); dt = time.perf_counter() - t0
    print("chain host call: %.1f ms for %.2f GB in (%.1f GB/s), launches %d" % (dt * 1e3, xn.nbytes / 1e9, xn.nbytes / dt / 1e9, ch.last_launches()))
    xd2 = torch.zeros((Cn, n), dtype=torch.complex64, device="cuda"); yd = torch.zeros((Cn, 1600), dtype=torch.float32, device="cuda")
    ch.execute_dev(xd2.data_ptr(), n, yd.data_ptr(), 1600, 0); L.synchronize()
    t0 = time.perf_counter(); ch.execute_dev(xd2.data_ptr(), n, yd.data_ptr(), 1600, 0); L.synchronize()
    print("chain device call (8192 ch): %.2f ms" % ((time.perf_counter() - t0) * 1e3))


def latency():
    """Single-channel call latency (the README use): sequential vs blocked-scan IIR, 64K-sample blocks."""
    x = am_iq(65536)
    for mode in ("sequential", "scan"):
        g = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075); g.set_mode(mode)
        g(x); g(x)
        t0 = time.perf_counter()
        for _ in range(10):
            g(x)
        print("ComplexIIRFilter 1 channel x 65536, %-10s %.3f ms per call" % (mode, (time.perf_counter() - t0) * 100))
    r = [L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075), L.ComplexResampler(0.024, Fc=0.024), L.AGC(), L.AmpModem(0.5, "dsb", True), L.DeemphasisFilter(48000)]
    def radio(v):
        for q in r:
            v = q(v)
        return v
    radio(x); radio(x)
    t0 = time.perf_counter()
    for _ in range(10):
        radio(x)
    print("README AMRadio 1 channel x 65536 (5 calls): %.3f ms per block (real time would be 32.8 ms)" % ((time.perf_counter() - t0) * 100))


if __name__ == "__main__":
    what = sys.argv[1:] or ["chain", "copy"]
    if "chain" in what:
        chain()
    if "copy" in what:
        copy()
    if "latency" in what:
        latency()
