"""CPU: the C-ABI library loads and exports every symbol include/liquiddsp_b200.h declares, the host-side
logic (design, planner, output-length bookkeeping, argument checking) behaves, and nothing computes
without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import liquiddsp as L
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "liquiddsp_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lqb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(L.lib_path)
    names = _declared()
    assert len(names) > 70
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.lqb_version() >= 100


def test_python_binding_covers_the_header():
    bound = set(L._SIG) | {"lqb_last_error"}
    assert set(_declared()) <= bound, sorted(set(_declared()) - bound)


def test_reference_api_surface():
    """Names, keyword arguments and defaults of wrapper.cpp."""
    import inspect
    sig = lambda c: {k: v.default for k, v in inspect.signature(c.__init__).parameters.items() if k not in ("self", "channels", "sos")}
    assert sig(L.ComplexIIRFilter) == dict(filter_type="butter", band_type="lowpass", order=2, Fc=0.2, F0=0.3, Ap=0.7, As=60.0)
    assert sig(L.ComplexResampler) == dict(rate=inspect._empty, len=20, Fc=None, As=60.0, nfilter=13)
    assert sig(L.AmpModem) == dict(modulation=0.75, type="dsb", carrier=False)
    assert sig(L.DeemphasisFilter) == dict(sample_rate=48000)
    assert sig(L.NCO) == dict(type="nco") and sig(L.FreqDem) == dict(kd=inspect._empty) and sig(L.AGC) == {}
    for name in ("squelch", "threshold", "bandwidth", "level", "level_dB", "lock", "gain", "scale", "status"):
        assert isinstance(getattr(L.AGC, name), property)
    for name in ("freq", "phase"):
        assert isinstance(getattr(L.NCO, name), property)
    for meth in ("adjust_frequency", "adjust_phase", "set_pll_bandwidth", "pll_step", "mix_up", "mix_down", "print"):
        assert callable(getattr(L.NCO, meth))
    with pytest.raises(TypeError):
        L.ComplexResampler(0.024)                          # Fc has no default in the reference either
    f = L.ComplexIIRFilter("nonsense", "whatever", order=2, Fc=0.2)    # silent fallback, empty property
    assert f.filter_type == "" and f.band_type == ""
    assert np.allclose(f.sos()[0], L.ComplexIIRFilter("butter", "lowpass", order=2, Fc=0.2).sos()[0])
    assert L.NCO("anything").type == "vco"                 # nco.hpp:16-24
    # the real-sample and fixed-band families (wrapper.cpp:36-132, 154-172, 244-257)
    assert sig(L.RealIIRFilter) == sig(L.ComplexIIRFilter)
    for cls in (L.CLowpassIIR, L.CHighpassIIR, L.RLowpassIIR, L.RHighpassIIR):
        assert sig(cls) == dict(filter_type="butter", order=2, Fc=0.2, Ap=0.5, As=20.0)
    for cls in (L.CBandpassIIR, L.CBandstopIIR, L.RBandpassIIR, L.RBandstopIIR):
        assert sig(cls) == dict(filter_type="butter", order=2, Fc=0.2, F0=0.3, Ap=0.5, As=20.0)
    assert sig(L.CIIRFilter) == sig(L.RIIRFilter) == dict(Bc=inspect._empty, Ac=inspect._empty)
    assert sig(L.RealResampler) == sig(L.ComplexResampler)
    assert sig(L.CResampler) == sig(L.RResampler) == dict(rate=inspect._empty) and sig(L.Delay) == dict(nd=1)
    d = L.Delay(3); d.delay = 7
    assert d.delay == 7 and d(np.zeros(4, np.int16)) is None
    for rate in (0.08, 0.5, 1.0, 2.2):                   # create_default: fc = min(0.49, rate/2), 64 filters of 14 taps
        g, o = L.CResampler(rate), O.CResampler(rate)
        assert g.state()[0] == o.step and g.bank().shape == (64, 14) and np.array_equal(g.bank(), o.bank())
        assert np.array_equal(L.RResampler(rate).bank(), o.bank())
    assert np.array_equal(L.RealResampler(0.3, Fc=0.1).bank(), O.RealResampler(0.3, Fc=0.1).bank())
    assert L.Chain(L.FreqDem(0.1), L.RResampler(0.08), L.DeemphasisFilter()).plan() == "seq[freqdem] -> par[resamp] -> seq[deemph]"
    assert sig(L.FMStereo) == dict(iq_rate=600000.0, pcm_rate=48000.0)
    assert L.FMStereo().deemph() == O.FMStereo().deemph() and L.FMStereo().out_len(600000) == 96000
    assert L.FMStereo(240000.0, 44100.0).deemph() == O.FMStereo(240000.0, 44100.0).deemph()
    # pcm_rate > iq_rate: pairs only where the resamplers yield exactly one sample each (demod.hpp:45-48): a rate in (1, 2)
    # keeps 2 - rate of the inputs, from 2 on nothing
    assert L.FMStereo(48000.0, 96000.0).out_len(1000) == 0 and L.FMStereo(48000.0, 60000.0).out_len(4000) == 2 * 3000
    x = np.exp(2j * np.pi * np.cumsum(0.05 * np.sin(0.01 * np.arange(4000)))).astype(np.complex64)
    assert O.FMStereo(48000.0, 60000.0)(x).size == 2 * 3000 and O.FMStereo(48000.0, 96000.0)(x).size == 0
    assert sig(L.BroadcastAM) == dict(slen=25) and sig(L.SSBDemod) == dict(band=inspect._empty)
    assert sig(L.HilbertTransform) == dict(m=5, As=60.0)
    assert np.array_equal(L.SSBDemod("usb").hq(), O.SSBDemod("usb").hq()) and L.SSBDemod("anything").usb is False
    assert L.HilbertTransform()(np.zeros(4, np.float64)) is None          # utility.hpp:104: other dtypes -> None
    with pytest.raises(ValueError):
        L.HilbertTransform(m=1)
    assert sig(L.RealDCBlocker) == dict(slen=25, As=20.0)
    assert sig(L.RealKaiserBessel) == dict(flen=25, Fc=None, As=20.0, offset=0.0)      # wrapper.cpp:255: Fc required
    with pytest.raises(TypeError):
        L.RealKaiserBessel(31)
    assert L.CBandpassIIR.__name__ == "CBandpassIIR" and L.RLowpassIIR("cheby1", 4, 0.1).band_type == "lowpass"


def test_fixed_band_and_real_designs_equal_oracle():
    for name, args in (("CLowpassIIR", ("cheby2", 5, 0.1)), ("CHighpassIIR", ("butter", 3, 0.2)), ("CBandpassIIR", ("cheby1", 4, 0.05, 0.2)),
                       ("CBandstopIIR", ("butter", 2, 0.05, 0.3)), ("RLowpassIIR", ("butter", 6, 0.05)), ("RHighpassIIR", ("cheby1", 3, 0.1)),
                       ("RBandpassIIR", ("cheby2", 3, 0.05, 0.25)), ("RBandstopIIR", ("butter", 4, 0.1, 0.2))):
        g, o = getattr(L, name)(*args), getattr(O, name)(*args)
        (B, A), (Bo, Ao) = g.sos(), o.sos()
        assert B.shape == Bo.shape and np.max(np.abs(B - Bo)) < 2e-6 and np.max(np.abs(A - Ao)) < 2e-6, name
        assert abs(g.freqresponse(0.07) - o.freqresponse(0.07)) < 1e-4 * max(1.0, abs(o.freqresponse(0.07)))
    g, o = L.RealDCBlocker(25, 20.0), O.RealDCBlocker(25, 20.0)
    assert g.taps().size == 51 and abs(g.freqresponse(0.0)) < 1e-3 and abs(g.freqresponse(0.0) - o.freqresponse(0.0)) < 1e-6
    g, o = L.RealKaiserBessel(31, 0.1, 40.0), O.RealKaiserBessel(31, 0.1, 40.0)
    assert np.array_equal(g.taps(), o.taps()) and abs(abs(g.freqresponse(0.0)) - 1.0) < 1e-6
    assert g.freqresponse(0.0) == o.freqresponse(0.0) and abs(g.freqresponse(0.3)) < 0.02
    with pytest.raises(ValueError):
        L.RealKaiserBessel(31, 0.7)                        # cut-off outside (0, 0.5)
    b, a = [0.2, 0.3, -0.1], [2.0, -0.8, 0.3, 0.05]
    for g, o in ((L.CIIRFilter(b, a), O.CIIRFilter(b, a)), (L.RIIRFilter(b, a), O.RIIRFilter(b, a))):
        assert abs(g.freqresponse(0.13) - o.freqresponse(0.13)) < 1e-6
    with pytest.raises(ValueError):
        L.CIIRFilter(b, [0.0, 1.0])
    with pytest.raises(ValueError):
        L.RIIRFilter(np.ones(17), [1.0])
    assert L.Chain(L.CIIRFilter(b, a), L.CIIRFilter(np.ones(9), [1.0])).plan() == "seq[tf] -> seq[tf]"


@pytest.mark.parametrize("ft", ["butter", "cheby1", "cheby2", "ellip", "bessel"])
@pytest.mark.parametrize("bt", ["lowpass", "highpass", "bandpass", "bandstop"])
@pytest.mark.parametrize("order", [1, 2, 3, 5, 8])
def test_product_design_equals_oracle_design(ft, bt, order):
    g = L.ComplexIIRFilter(ft, bt, order=order, Fc=0.1, F0=0.2, Ap=1.0, As=40.0)
    B, A = g.sos(); Bo, Ao = O.iirdes_sos(ft, bt, order, 0.1, 0.2, 1.0, 40.0)
    assert B.shape == Bo.shape and np.max(np.abs(B - Bo)) < 2e-6 and np.max(np.abs(A - Ao)) < 2e-6
    H, Ho = g.freqresponse(0.07), O.ComplexIIRFilter(ft, bt, order, 0.1, 0.2, 1.0, 40.0).freqresponse(0.07)
    assert abs(H - Ho) < 1e-4 * max(1.0, abs(Ho))


def test_product_tables_equal_oracle_tables():
    r, ro = L.ComplexResampler(0.024, Fc=0.024), O.ComplexResampler(0.024, Fc=0.024)
    assert r.state() == (0x29AAAAC0, 0) and np.array_equal(r.bank(), ro.bank())
    for rate in (0.5, 1.7, 0.03, 0.0101, 3.0):
        assert L.ComplexResampler(rate, Fc=0.1).state()[0] == O.ComplexResampler(rate, Fc=0.1).step
    lp, dc = L.AmpModem(0.5, "dsb", True).taps(); lpo, dco = O.AmpModem(0.5, "dsb", True).taps()
    assert np.array_equal(lp, lpo) and np.array_equal(dc, dco)
    assert L.DeemphasisFilter(48000).coeffs() == O.DeemphasisFilter.coeffs(48000)
    assert L.DeemphasisFilter(44100).coeffs() == O.DeemphasisFilter.coeffs(44100)
    h = O.firdes_kaiser(64, 0.1, 60.0)
    assert np.array_equal(L.FIRFilter(h).taps(), h)
    assert abs(L.FIRFilter(h).freqresponse(0.05) - O.FIRFilter(h).freqresponse(0.05)) < 1e-5


def test_broadcast_am_design_equals_oracle():
    for m in (25, 7, 40):
        (lp, B, A), (lpo, Bo, Ao) = L.BroadcastAM(m).design(), O.BroadcastAM(m).design()
        assert lp.size == 2 * m + 1 and np.array_equal(lp, lpo)
        assert np.max(np.abs(B - Bo)) < 2e-6 and np.max(np.abs(A - Ao)) < 2e-6
    assert A[1, 2] == 0 and B[1, 2] == 0                   # order 3: one biquad + one first-order section
    with pytest.raises(ValueError):
        L.BroadcastAM(0)
    with pytest.raises(ValueError):
        L.BroadcastAM(65)


def _radio(ch=1):
    return (L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=ch), L.ComplexResampler(0.024, Fc=0.024, channels=ch),
            L.AGC(channels=ch), L.AmpModem(0.5, "dsb", True, channels=ch), L.DeemphasisFilter(48000, channels=ch))


def test_planner():
    assert L.Chain(*_radio(), fuse=0).plan() == "seq[iir4] -> par[resamp] -> seq[agc] -> am[ampmodem] -> seq[deemph]"
    assert L.Chain(*_radio(), fuse=1).plan() == "seq[iir4+resamp] -> am[agc+ampmodem+deemph]"
    assert L.Chain(*_radio(), fuse=2).plan() == "seq[iir4+resamp+agc+ampmodem+deemph]"
    assert L.Chain(L.NCO(), L.ComplexResampler(0.024, Fc=0.024)).plan() == "par[nco+resamp]"          # few channels: time-parallel
    big = 32768
    assert L.Chain(L.NCO(channels=big), L.ComplexResampler(0.024, Fc=0.024, channels=big)).plan() == "seq[nco+resamp]"
    assert L.Chain(L.NCO(channels=big), L.ComplexResampler(0.5, Fc=0.2, channels=big)).plan() == "par[nco+resamp]"
    assert L.Chain(L.NCO()).plan() == "par[nco]" and L.Chain(L.NCO(channels=big)).plan() == "seq[nco]"
    assert L.Chain(L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075), L.AGC(), L.FreqDem(0.1)).plan() == "seq[iir4+agc+freqdem]"
    assert L.Chain(L.ComplexResampler(0.5, Fc=0.2)).plan() == "par[resamp]"          # not decimating enough to fuse
    assert L.Chain(L.ComplexIIRFilter("butter", "bandpass", order=10, Fc=0.1, F0=0.2)).plan() == "seq[iir8] -> seq[iir2]"
    h = np.ones(8, np.float32)
    assert L.Chain(L.FIRFilter(h), L.ComplexIIRFilter(order=2), L.FIRFilter(h)).plan() == "fir -> seq[iir1] -> fir"
    assert L.Chain(*_radio()).out_len(65536) == 1573
    assert L.Chain(L.RealIIRFilter("butter", "bandpass", order=10, Fc=0.1, F0=0.2), L.DeemphasisFilter()).plan() == "seq[iir8] -> seq[iir2] -> seq[deemph]"
    assert L.Chain(L.AmpModem(0.5, "dsb", True), L.RealDCBlocker(), L.RLowpassIIR("butter", 4, 0.1)).plan() == "am[ampmodem] -> fir -> seq[iir2]"
    rs = lambda ch=1: (L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=ch), L.ComplexResampler(0.024, Fc=0.024, channels=ch))
    assert L.Chain(*rs(), L.BroadcastAM()).plan() == "seq[iir4+resamp] -> bam[broadcast_am]"
    assert L.Chain(*rs(), L.AGC(), L.BroadcastAM(), L.DeemphasisFilter()).plan() == "seq[iir4+resamp] -> bam[agc+broadcast_am+deemph]"
    assert L.Chain(L.AGC(), L.BroadcastAM(30)).plan() == "seq[agc] -> bam[broadcast_am]"      # no hand-off buffer to run the AGC in
    with pytest.raises(ValueError):                        # real-input stage after a complex-output stage
        L.Chain(L.ComplexIIRFilter(order=2), L.DeemphasisFilter()).plan()
    with pytest.raises(ValueError):
        L.Chain(L.ComplexIIRFilter(order=2, channels=2), L.AGC(channels=3)).plan()


def test_argument_checking():
    with pytest.raises(ValueError):
        L.ComplexIIRFilter("butter", order=0)
    with pytest.raises(ValueError):
        L.ComplexIIRFilter("butter", order=2, Fc=0.7)
    with pytest.raises(ValueError):
        L.ComplexIIRFilter("ellip", order=4, Fc=0.1, As=-3.0)
    assert L.Chain(L.AmpModem(0.5, "usb", True)).plan() == "am[carrier-loop] -> fir[hilbert] -> fir[dcblock]"
    assert L.Chain(L.AGC(), L.AmpModem(0.5, "lsb", False), L.DeemphasisFilter()).plan() == "seq[agc] -> fir[hilbert] -> seq[deemph]"
    # fusion level 2 away from the one fully fused pattern still finds a kernel for the ampmodem
    assert L.Chain(L.FIRFilter(np.ones(4, np.float32)), L.AmpModem(0.5, "dsb", True), L.DeemphasisFilter(), fuse=2).plan() == "fir -> am[ampmodem+deemph]"
    with pytest.raises(ValueError):
        L.ComplexResampler(1e-4, Fc=0.1)
    with pytest.raises(ValueError):
        L.FreqDem(-1.0)
    with pytest.raises(ValueError):
        L.FIRFilter(np.zeros(0, np.float32))
    with pytest.raises(ValueError):
        L.AGC(channels=0)
    a = L.AGC()
    with pytest.raises(ValueError):
        a.bandwidth = 2.0
    with pytest.raises(ValueError):
        a.scale = 0.0
    a.bandwidth = 0.05; a.scale = 0.5; a.threshold = -12.0
    assert (a.bandwidth, a.scale, a.threshold) == (np.float32(0.05), 0.5, -12.0)
    a.squelch = True
    assert a.squelch and a.status == 1
    a.squelch = False
    assert a.status == 7


def test_no_compute_without_a_gpu():
    """There is no CPU fallback: with no device every execute fails loudly (skipped where a GPU exists)."""
    try:
        if L.device_count() > 0:
            pytest.skip("a CUDA device is present")
    except RuntimeError:
        pass
    x = np.zeros(64, np.complex64)
    for obj in (L.ComplexIIRFilter(order=2), L.FIRFilter(np.ones(4, np.float32)), L.ComplexResampler(0.024, Fc=0.024),
                L.NCO(), L.AGC(), L.AmpModem(0.5, "dsb", True), L.FreqDem(0.1), L.Chain(*_radio())):
        with pytest.raises(RuntimeError):
            obj(x)
    with pytest.raises(RuntimeError):
        L.DeemphasisFilter()(np.zeros(8, np.float32))
    with pytest.raises(RuntimeError):
        L.NCO().freq = 0.1


def test_pcm_framer_follows_the_readme_loop():
    """README.md:53-58: self.pcm += pcm.tobytes(); while len(self.pcm) > 4096: emit self.pcm[:4096]."""
    rng = np.random.default_rng(5)
    fr = L.PcmFramer()
    pcm, ref_chunks, got = b"", [], []
    for n in (1000, 24, 1024, 3000, 0, 1, 1023, 2048):
        audio = rng.standard_normal(n).astype(np.float32)
        pcm += audio.tobytes()
        while len(pcm) > 4096:
            ref_chunks.append(pcm[:4096]); pcm = pcm[4096:]
        got += fr.push(audio)
        assert fr.pending == len(pcm)
    assert got == ref_chunks and len(got) >= 7
    assert fr.flush() == pcm and fr.pending == 0
    # exactly one chunk pending stays behind under the README's strict '>' and leaves at once with strict=False
    a = np.zeros(1024, np.float32)
    assert L.PcmFramer().push(a) == [] and len(L.PcmFramer(strict=False).push(a)) == 1
    # batched: channels advance in lockstep, row c of a chunk is channel c's bytes
    C = 3
    frb, frs = L.PcmFramer(channels=C), [L.PcmFramer() for _ in range(C)]
    for n in (1500, 700, 2100):
        blk = rng.standard_normal((C, n)).astype(np.float32)
        chunks = frb.push(blk)
        singles = [f.push(blk[c]) for c, f in enumerate(frs)]
        assert all(len(s_) == len(chunks) for s_ in singles)
        for k, ch in enumerate(chunks):
            assert ch.shape == (C, 4096) and all(ch[c].tobytes() == singles[c][k] for c in range(C))
    with pytest.raises(ValueError):
        frb.push(np.zeros((2, 8), np.float32))
