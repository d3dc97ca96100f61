"""Randomised chains: random stage combinations, channel counts, call splits and fusion levels against the oracle's
stage-by-stage composition.  Catches planner / hand-off / carried-state mistakes that the per-stage tests cannot.
Tolerances: 1e-4 end to end (BASELINE.json); chains that put a carrier PLL behind a stage whose rounding differs
from the oracle's (FIR summation order, atan2f) get 3e-3 -- the loop amplifies last-bit differences;
behind such a stage the arg()-detector loops (BroadcastAM, FMStereo) are checked for shape and finiteness only."""
import numpy as np
import pytest

import liquiddsp as L
from oracle import oracle as O
from util import rel_l2, am_iq, split_points

pytestmark = pytest.mark.gpu


def _complex_stage(rng, C):
    k = int(rng.integers(0, 8))
    if k == 7:
        nd = int(rng.integers(0, 40))
        g = L._DelayLine(nd, False, C); o = O.Delay(nd)
        return g, o, True, "delay%d" % nd
    if k == 0:
        f = float(rng.uniform(-0.5, 0.5)); down = bool(rng.integers(0, 2))
        g = L.NCO(channels=C); g.freq = f; g.set_direction(down)
        o = O.NCO(); o.freq = f
        return g, (lambda v, o=o, down=down: o.mix_down(v) if down else o.mix_up(v)), True, "nco"
    if k == 1:
        ft = ["butter", "cheby1", "cheby2", "ellip", "bessel"][int(rng.integers(0, 5))]
        bt = ["lowpass", "highpass", "bandpass", "bandstop"][int(rng.integers(0, 4))]
        order = int(rng.integers(1, 7))
        g = L.ComplexIIRFilter(ft, bt, order=order, Fc=0.12, F0=0.25, Ap=1.0, As=40.0, channels=C)
        return g, O.ComplexIIRFilter(_sos=g.sos()), True, "iir-%s-%s-%d" % (ft, bt, order)
    if k == 2:
        h = O.firdes_kaiser(int(rng.integers(2, 90)), 0.2, 50.0)
        return L.FIRFilter(h, channels=C), O.FIRFilter(h), False, "fir%d" % h.size
    if k == 3:
        rate = [0.024, 0.1, 0.5, 1.7][int(rng.integers(0, 4))]
        fc = min(0.45, rate / 2)
        return L.ComplexResampler(rate, Fc=fc, channels=C), O.ComplexResampler(rate, Fc=fc), True, "resamp%g" % rate
    if k == 4:
        g = L.AGC(channels=C); o = O.AGC()
        g.scale = o.scale = 0.5
        # precision 'auto': the bit-exact loop unless a FreqDem follows and no PLL demodulator does (then 1e-7 from it,
        # under a feed-forward discriminator checked at 1e-4)
        return g, o, True, "agc"
    if k == 5:
        b = (0.3 * rng.standard_normal(int(rng.integers(1, 6)))).astype(np.float32); a = np.array([1.0, -0.5, 0.1], np.float32)
        return L.CIIRFilter(b, a, channels=C), O.CIIRFilter(b, a), True, "tf"
    rate = [0.3, 0.9][int(rng.integers(0, 2))]
    return L.CResampler(rate, channels=C), O.CResampler(rate), True, "cresamp%g" % rate


def _demod(rng, C):
    k = int(rng.integers(0, 6))
    if k == 5:
        return L.FMStereo(48000.0, 12000.0, channels=C), O.FMStereo(48000.0, 12000.0), "arg-pll", "fmstereo"
    if k == 0:
        car = bool(rng.integers(0, 2))
        return L.AmpModem(0.5, "dsb", car, channels=C), O.AmpModem(0.5, "dsb", car), "pll", "am-dsb-%d" % car
    if k == 1:
        return L.FreqDem(0.2, channels=C), O.FreqDem(0.2), "atan", "fm"
    if k == 2:
        g = L.BroadcastAM(int(rng.integers(3, 40)), channels=C)
        return g, O.BroadcastAM(len(g.design()[0]) // 2, _dcblock=g.design()[1:]), "arg-pll", "bam"
    if k == 3:
        band = ["usb", "lsb"][int(rng.integers(0, 2))]
        return L.SSBDemod(band, channels=C), O.SSBDemod(band), "fir", "ssb-" + band
    typ = ["usb", "lsb"][int(rng.integers(0, 2))]
    return L.AmpModem(0.7, typ, False, channels=C), O.AmpModem(0.7, typ, False), "fir", "am-" + typ


def _real_stage(rng, C):
    k = int(rng.integers(0, 5))
    if k == 0:
        return L.DeemphasisFilter(48000, channels=C), O.DeemphasisFilter(48000), True, "deemph"
    if k == 1:
        g = L.RealIIRFilter("cheby1", "lowpass", int(rng.integers(1, 6)), 0.2, channels=C)
        return g, O.RealIIRFilter(_sos=g.sos()), True, "riir"
    if k == 2:
        h = O.firdes_kaiser(int(rng.integers(2, 60)), 0.2, 40.0)
        return L.RealFIRFilter(h, channels=C), O.RealFIRFilter(h), False, "rfir%d" % h.size
    if k == 3:
        return L.RealResampler(0.6, Fc=0.25, channels=C), O.RealResampler(0.6, Fc=0.25), True, "rresamp"
    b = np.array([0.2, 0.1], np.float32); a = np.array([1.0, -0.7], np.float32)
    return L.RIIRFilter(b, a, channels=C), O.RIIRFilter(b, a), True, "rtf"


@pytest.mark.parametrize("seed", range(64))
def test_random_chain(cuda, seed):
    rng = np.random.default_rng(1000 + seed)
    C = int([1, 2, 3, 33, 70, 130][int(rng.integers(0, 6))])
    n = int(rng.integers(3000, 9000)) if seed % 4 else int(rng.integers(40, 400))
    x = np.stack([am_iq(n, fs=48000.0, f_off=20.0 + 3 * c, phase=0.1 * c, seed=seed * 100 + c, noise=0.02, amp=0.6) for c in range(C)])
    gs, os_, names, exact = [], [], [], True
    for _ in range(int(rng.integers(1, 4))):
        g, o, ex, nm = _complex_stage(rng, C); gs.append(g); os_.append(o); names.append(nm); exact &= ex
    pll_after_inexact = arg_after_inexact = False
    if rng.integers(0, 3) > 0:
        g, o, kind, nm = _demod(rng, C); gs.append(g); os_.append(o); names.append(nm)
        pll_after_inexact = kind == "pll" and not exact
        # arg() is scale-invariant: where the filtered carrier is still ~0 (start of a stream) last-bit differences of
        # the input turn into O(1) phase-detector differences and a different lock-in transient -- values are only
        # comparable when everything upstream is bit-exact
        arg_after_inexact = kind == "arg-pll" and not exact
        exact &= kind in ("pll", "arg-pll")
        for _ in range(int(rng.integers(0, 3))):
            g, o, ex, nm = _real_stage(rng, C); gs.append(g); os_.append(o); names.append(nm); exact &= ex
    fuse = int(rng.integers(0, 3))
    chain = L.Chain(*gs, fuse=fuse)
    cuts = split_points(n, int(rng.integers(1, 4)), rng)
    y = np.concatenate([chain(x[:, s:e] if C > 1 else x[0, s:e]).reshape(C, -1) for s, e in cuts], axis=1)
    tol = 3e-3 if pll_after_inexact else 1e-4
    for c in sorted(set([0, C // 2, C - 1])):
        if c > 0:                                           # fresh oracle objects per probed channel
            rng2 = np.random.default_rng(1000 + seed)       # replay the construction draws
            int([1, 2, 3, 33, 70, 130][int(rng2.integers(0, 6))]); int(rng2.integers(3000, 9000)) if seed % 4 else int(rng2.integers(40, 400))
            os_c = []
            for _ in range(int(rng2.integers(1, 4))):
                os_c.append(_complex_stage(rng2, 1)[1])
            if rng2.integers(0, 3) > 0:
                gd, od, _, _ = _demod(rng2, 1)
                if isinstance(gs[len(os_c)], L.BroadcastAM):
                    od = O.BroadcastAM(len(gs[len(os_c)].design()[0]) // 2, _dcblock=gs[len(os_c)].design()[1:])
                os_c.append(od)
                for _ in range(int(rng2.integers(0, 3))):
                    os_c.append(_real_stage(rng2, 1)[1])
            for i, g in enumerate(gs):                       # filters designed on the product side: same coefficients
                if isinstance(g, (L.ComplexIIRFilter, L.RealIIRFilter)) and not isinstance(g, (L.CIIRFilter,)):
                    os_c[i] = (O.RealIIRFilter if isinstance(g, L.RealIIRFilter) else O.ComplexIIRFilter)(_sos=g.sos())
            cur = os_c
        else:
            cur = os_
        pieces = []
        for s, e in cuts:
            v = x[c, s:e]
            for o in cur:
                v = o(v)
            pieces.append(v)
        yo = np.concatenate(pieces)
        assert yo.shape == y[c].shape, (names, fuse, C, n, cuts)
        if arg_after_inexact:
            assert np.all(np.isfinite(y[c]))
            continue
        assert rel_l2(y[c], yo) <= tol, (rel_l2(y[c], yo), names, fuse, chain.plan(), C, n, cuts, c)


@pytest.mark.parametrize("seed", range(12))
def test_random_chain_wide(cuda, seed):
    """The same idea at channel counts that switch the kernels' regimes (8 / 16 / 32 channels per warp, TMA staging, the
    two-channel front kernel, paired real FIR rows): 64 distinct signals tiled over the channel axis."""
    rng = np.random.default_rng(5000 + seed)
    C = int([20000, 30001, 33000, 60007][int(rng.integers(0, 4))])
    n = int(rng.integers(400, 1400))
    base = np.stack([am_iq(n, fs=48000.0, f_off=5.0 + 0.5 * c, phase=0.1 * c, seed=seed * 100 + c, noise=0.02, amp=0.6) for c in range(64)])
    x = np.tile(base, (C // 64 + 1, 1))[:C].copy()
    gs, os_, names, exact = [], [], [], True
    for _ in range(int(rng.integers(1, 4))):
        g, o, ex, nm = _complex_stage(rng, C); gs.append(g); os_.append(o); names.append(nm); exact &= ex
    kind = None
    if rng.integers(0, 3) > 0:
        g, o, kind, nm = _demod(rng, C); gs.append(g); os_.append(o); names.append(nm)
        for _ in range(int(rng.integers(0, 3))):
            g, o, ex, nm = _real_stage(rng, C); gs.append(g); os_.append(o); names.append(nm)
    fuse = int(rng.integers(0, 3))
    chain = L.Chain(*gs, fuse=fuse)
    cut = int(rng.integers(1, n))
    y = np.concatenate([chain(x[:, :cut]), chain(x[:, cut:])], axis=1)
    assert np.array_equal(y[:64].view(np.uint32), y[64:128].view(np.uint32)), (names, chain.plan())     # position in the grid is irrelevant
    assert np.array_equal(y[C - 1].view(np.uint32), y[(C - 1) % 64].view(np.uint32)), (names, chain.plan())
    if kind in ("pll", "arg-pll") and not exact:
        return
    v1, v2 = x[0, :cut], x[0, cut:]
    for o in os_:
        v1 = o(v1)
    for o in os_:
        v2 = o(v2)
    yo = np.concatenate([v1, v2])
    assert yo.shape == y[0].shape and rel_l2(y[0], yo) <= 1e-4, (rel_l2(y[0], yo), names, fuse, chain.plan(), C, n, cut)
