#!/usr/bin/env python
"""Small cases that walk every hand-rolled synchronisation path of the kernels (TMA rings, mbarrier phases, warp-to-warp
shared-memory hand-offs, ragged last boxes, carried rings) -- the workload for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_cases.py

Sizes are tiny (the sanitizer slows kernels 10-100x); every case is checked against the oracle so a 'clean' run is also
a correct one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "python-liquiddsp_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import liquiddsp as L
from oracle import oracle as O
from util import am_iq


def crandn(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def radio(mod, **kw):
    iir = mod.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / 2e6, **kw)
    rs = mod.ComplexResampler(rate=0.024, Fc=0.024, **kw)
    agc = mod.AGC(**kw); agc.lock = False; agc.scale = 0.01
    am = mod.AmpModem(modulation=0.5, type="dsb", carrier=True, **kw)
    de = mod.DeemphasisFilter(48000, **kw)
    return iir, rs, agc, am, de


def check(name, got, ref, tol=0.0):
    err = float(np.linalg.norm(got.astype(np.complex128) - ref) / max(np.linalg.norm(ref), 1e-30))
    print("%-46s rel-L2 %.2e" % (name, err)); assert err <= tol, name


def main():
    L.set_device(0)
    # the receiver front with 2 / 4 / 8 lanes per channel, ragged calls (n not a multiple of 16, short last TMA box), state carried
    for lanes, C in ((2, 37), (4, 9), (8, 3)):
        os.environ["LQB_LANES"] = str(lanes)
        st = radio(L, channels=C); ch = L.Chain(*st)
        x = np.stack([am_iq(9000 + 2 * 17, seed=40 + c) for c in range(C)])
        cuts = [(0, 4098), (4098, 4100), (4100, 9034)]
        y = np.concatenate([ch(np.ascontiguousarray(x[:, s:e])) for s, e in cuts], axis=1)
        ro = radio(O); ro = (O.ComplexIIRFilter(_sos=st[0].sos()),) + ro[1:]
        ref = np.concatenate([ro[4](ro[3](ro[2](ro[1](ro[0](x[C - 1, s:e]))))) for s, e in cuts])
        check("front %d lanes/channel + tail, %s" % (lanes, ",".join(ch.last_kernels())), y[C - 1], ref, 1e-4)
    os.environ.pop("LQB_LANES")
    # thread-per-channel tail (many-channel path) on a small batch
    os.environ["LQB_NO_AMTAIL8"] = "1"
    st = radio(L, channels=70); ch = L.Chain(*st)
    x = np.stack([am_iq(8192, seed=80 + c) for c in range(70)])
    y = ch(x); ro = radio(O); ro = (O.ComplexIIRFilter(_sos=st[0].sos()),) + ro[1:]
    check("tail, one thread per channel", y[69], ro[4](ro[3](ro[2](ro[1](ro[0](x[69]))))), 1e-4)
    os.environ.pop("LQB_NO_AMTAIL8")
    # full-rate sequential kernel with TMA loads and bulk tensor stores (MODE 2): cascade alone, many channels x few samples
    i2 = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=12000)
    xb = L.DeviceBuffer(12000 * 256 * 8); yb = L.DeviceBuffer(12000 * 256 * 8)
    L.synth_fill(1, xb.ptr.value, 12000, 256)
    L.Chain(i2).execute_dev(xb.ptr.value, 256, yb.ptr.value, 256, 0); L.synchronize()
    xs = xb.download((12000, 256), np.complex64); ys = yb.download((12000, 256), np.complex64)
    check("seq[iir4] TMA in / TMA out", ys[11999], O.ComplexIIRFilter(_sos=i2.sos())(xs[11999]))
    # cascade -> gain loop -> discriminator: three-warp pipeline with mbarrier rings
    c4 = L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=40), L.AGC(channels=40), L.FreqDem(0.1, channels=40))
    x4 = np.stack([crandn(2050, seed=c) for c in range(40)])
    y4 = c4(x4)
    r4 = O.FreqDem(0.1)(O.AGC()(O.ComplexIIRFilter(_sos=c4.stages[0].sos())(x4[39])))
    print("%-46s rms %.2e (%s)" % ("pipe: cascade | gain loop | discriminator", float(np.sqrt(np.mean((y4[39] - r4) ** 2))), ",".join(c4.last_kernels())))
    # FIR (uniform taps, double-buffered tiles) and the time-parallel NCO + resampler (persistent CTAs, bulk copies)
    h = O.firdes_kaiser(64, 0.1, 60.0)
    xf = np.stack([crandn(5000, seed=9 + c) for c in range(3)])
    check("fir", L.FIRFilter(h, channels=3)(xf)[2], O.FIRFilter(h)(xf[2]), 1e-5)
    gn, on = L.NCO(channels=5), O.NCO(); gn.freq = 0.3; on.freq = 0.3; gn.set_direction(True)
    cp = L.Chain(gn, L.ComplexResampler(0.024, Fc=0.024, channels=5))
    xr = np.stack([crandn(7000, seed=3)] * 5)
    check("par[nco+resamp]", cp(xr)[4], O.ComplexResampler(0.024, Fc=0.024)(on.mix_down(xr[4])))
    print("sanitize cases ok")


if __name__ == "__main__":
    main()
