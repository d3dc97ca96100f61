"""GPU parity: every stage of the hot path and the fused chains, called through
the C ABI (ctypes module `liquiddsp`) and compared with the CPU oracle on identical inputs and
identical coefficients.  Tolerances are BASELINE.json's: rel-L2 <= 1e-5 per stage, <= 1e-4 end to end;
integer phase / output counts bit-exact.  Where the kernel evaluates the oracle's operations in the
oracle's order the test demands bit equality.
"""
import numpy as np
import pytest

import liquiddsp as L
from oracle import oracle as O
from util import rel_l2, crandn, am_iq, fm_iq, fm_stereo_iq, split_points

pytestmark = pytest.mark.gpu
TOL_STAGE = 1e-5
TOL_E2E = 1e-4


def _bits_or_close(y, yo, tol=TOL_STAGE):
    """The PLL branch takes arg() through a double-precision atan2 on both sides; the two math libraries agree after
    rounding to float except with probability ~1e-8 per sample, so equality is expected and 1e-5 is demanded."""
    return np.array_equal(y.view(np.uint32), yo.view(np.uint32)) or rel_l2(y, yo) <= tol


# ------------------------------------------------------------------------------------------- a1
@pytest.mark.parametrize("ft,bt,order", [("cheby2", "lowpass", 8), ("butter", "lowpass", 2), ("butter", "lowpass", 1),
                                          ("cheby1", "highpass", 5), ("cheby2", "bandpass", 4), ("butter", "bandstop", 10),
                                          ("ellip", "lowpass", 7), ("ellip", "bandpass", 3), ("bessel", "highpass", 6)])
def test_iir_single_channel_bit_exact(cuda, ft, bt, order):
    rng = np.random.default_rng(1)
    fc = 0.0075 if (ft, order) == ("cheby2", 8) else 0.1
    g = L.ComplexIIRFilter(ft, bt, order=order, Fc=fc, F0=0.2, Ap=1.0, As=60.0)
    o = O.ComplexIIRFilter(_sos=g.sos())
    x = crandn(rng, 20000)
    y, yo = g(x), o(x)
    assert y.dtype == np.complex64 and y.shape == x.shape
    assert np.array_equal(y.view(np.uint32), yo.view(np.uint32)), rel_l2(y, yo)


@pytest.mark.parametrize("n", [1000, 1001, 7, 16, 17])
def test_iir_batched_ragged(cuda, n):
    rng = np.random.default_rng(2)
    C = 130                                             # not a multiple of the 64-channel CTA
    g = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C)
    x = crandn(rng, C, n)
    y = g(x)
    for c in (0, 1, 63, 64, 129):
        o = O.ComplexIIRFilter(_sos=g.sos())
        assert np.array_equal(y[c].view(np.uint32), o(x[c]).view(np.uint32)), (c, n)


@pytest.mark.parametrize("C", [28416 + 7, 56832 + 33])
def test_iir_resampler_wide_batches(cuda, C):
    """The channel-count regimes of the sequential kernel (8 / 16 / 32 channels per warp) give the same bits."""
    rng = np.random.default_rng(21)
    n = 1500
    iir = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C); rs = L.ComplexResampler(0.024, Fc=0.024, channels=C)
    chain = L.Chain(iir, rs)
    assert chain.plan() == "seq[iir4+resamp]"
    base = crandn(rng, 64, n)
    x = np.tile(base, (C // 64 + 1, 1))[:C].copy()
    ys = [chain(x), chain(x)]                               # two calls: carried state
    for c in (0, 63, 64 * 100 + 5, C - 1):
        oi, ors = O.ComplexIIRFilter(_sos=iir.sos()), O.ComplexResampler(0.024, Fc=0.024)
        for k in range(2):
            assert np.array_equal(ys[k][c].view(np.uint32), ors(oi(x[c])).view(np.uint32)), (C, c, k)
    assert np.array_equal(ys[1][:64], ys[1][64:128])        # identical inputs -> identical outputs across CTAs


def test_iir_streaming_invariance(cuda):
    rng = np.random.default_rng(3)
    x = crandn(rng, 3, 30011)
    a = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=3)
    b = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=3)
    whole = a(x)
    parts = np.concatenate([b(x[:, s:e]) for s, e in split_points(x.shape[1], 7, rng)], axis=1)
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))
    a.reset()
    assert np.array_equal(a(x).view(np.uint32), whole.view(np.uint32))


def test_iir_blocked_scan_mode(cuda):
    """Time-parallel blocked scan (opt-in): fp32 block-local pass + double carried state.  A reordered IIR cannot be
    within 1e-5 of the sequential fp32 result (the filter's own rounding noise is ~5e-5); the bar is
    <= 1e-4 against the oracle and no further from the fp64 evaluation than the oracle itself is."""
    n, blk = 6 * 65536, 65536
    x = am_iq(n, seed=11)
    g = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075); g.set_mode("scan")
    assert L.Chain(g).plan() == "scan[iir4]"
    B, A = g.sos()
    o = O.ComplexIIRFilter(_sos=(B, A))
    y = np.concatenate([g(x[i:i + blk]) for i in range(0, n, blk)])       # state carried across scan calls
    yo = np.concatenate([o(x[i:i + blk]) for i in range(0, n, blk)])
    truth = O.iir_f64_truth(B, A, x)
    e_scan, e_orc = rel_l2(y, truth), rel_l2(yo, truth)
    assert rel_l2(y, yo) <= 1e-4
    assert e_scan <= 1.25 * e_orc + 1e-6
    # a call length that cannot be cut into equal blocks takes the sequential kernel, from the scan's carried state
    tail_x = am_iq(1001, seed=12)
    yt, yto = g(tail_x), o(tail_x)
    assert rel_l2(yt, yto) <= 1e-3
    # batched
    gb = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=3); gb.set_mode("scan")
    xb = np.stack([x[:blk], x[blk:2 * blk], x[2 * blk:3 * blk]])
    yb = gb(xb)
    for c in range(3):
        assert rel_l2(yb[c], O.ComplexIIRFilter(_sos=(B, A))(xb[c])) <= 1e-4


def test_iir_empty_and_dtype_cast(cuda):
    g = L.ComplexIIRFilter("butter", order=2, Fc=0.2)
    assert g(np.zeros(0, np.complex64)).shape == (0,)
    x = np.linspace(0, 1, 64)                            # float64 -> forcecast to complex64, as pybind11 does
    o = O.ComplexIIRFilter(_sos=g.sos())
    assert np.array_equal(g(x), o(x))
    with pytest.raises(ValueError):
        g(np.zeros((2, 8), np.complex64))


# ------------------------------------------------------------------------------------------- a2
@pytest.mark.parametrize("ntaps,C,n", [(64, 3, 5000), (64, 1, 2047), (51, 2, 2049), (1, 2, 100), (17, 1, 5), (200, 1, 9000)])
def test_fir_crcf(cuda, ntaps, C, n):
    rng = np.random.default_rng(4)
    h = O.firdes_kaiser(ntaps, 0.1, 60.0) if ntaps > 1 else np.array([0.5], np.float32)
    g = L.FIRFilter(h, channels=C)
    x = crandn(rng, C, n)
    y = g(x) if C > 1 else g(x[0]).reshape(1, -1)
    for c in range(C):
        assert rel_l2(y[c], O.FIRFilter(h)(x[c])) <= TOL_STAGE


def test_fir_streaming_invariance_and_scale(cuda):
    rng = np.random.default_rng(5)
    h = O.firdes_kaiser(64, 0.1, 60.0)
    x = crandn(rng, 2, 12000)
    a, b = L.FIRFilter(h, channels=2), L.FIRFilter(h, channels=2)
    whole = a(x)
    parts = np.concatenate([b(x[:, s:e]) for s, e in split_points(x.shape[1], 6, rng)], axis=1)
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))
    a.reset(); a.set_scale(0.25)
    assert rel_l2(a(x), 0.25 * whole) <= 1e-6


# ------------------------------------------------------------------------- 8f row 1: rrrf families
@pytest.mark.parametrize("name,args", [("RLowpassIIR", ("cheby2", 8, 0.05)), ("RHighpassIIR", ("butter", 3, 0.2)),
                                       ("RBandpassIIR", ("cheby1", 4, 0.05, 0.2)), ("RBandstopIIR", ("butter", 10, 0.05, 0.3))])
def test_real_iir_families_bit_exact(cuda, name, args):
    rng = np.random.default_rng(31)
    g = getattr(L, name)(*args)
    o = O.RealIIRFilter(_sos=g.sos())
    x = rng.standard_normal(9001).astype(np.float32)
    y, yo = np.concatenate([g(x[:4000]), g(x[4000:])]), o(x)          # state carried across calls
    assert y.dtype == np.float32 and np.array_equal(y.view(np.uint32), yo.view(np.uint32)), rel_l2(y, yo)
    g.reset()
    assert np.array_equal(g(x[:100]), yo[:100])


def test_real_iir_batched_matches_complex_lanes(cuda):
    """C real channels through RealIIRFilter = the real lane of ComplexIIRFilter on the same samples."""
    rng = np.random.default_rng(32)
    C, n = 200, 777
    x = rng.standard_normal((C, n)).astype(np.float32)
    r = L.RealIIRFilter("cheby2", "lowpass", 8, 0.05, channels=C)
    c = L.ComplexIIRFilter("cheby2", "lowpass", 8, 0.05, channels=C)
    yr, yc = r(x), c(x.astype(np.complex64))
    assert np.array_equal(yr, yc.real) and not np.any(yc.imag)
    o = O.RealIIRFilter(_sos=r.sos())
    assert np.array_equal(yr[137], o(x[137]))


@pytest.mark.parametrize("name,args", [("CLowpassIIR", ("cheby2", 5, 0.1)), ("CHighpassIIR", ("butter", 3, 0.2)),
                                       ("CBandpassIIR", ("cheby1", 4, 0.05, 0.2)), ("CBandstopIIR", ("butter", 2, 0.05, 0.3))])
def test_fixed_band_complex_iir_bit_exact(cuda, name, args):
    rng = np.random.default_rng(33)
    g = getattr(L, name)(*args)
    o = O.ComplexIIRFilter(_sos=g.sos())
    x = crandn(rng, 5000)
    y = g(x)
    assert np.array_equal(y.view(np.uint32), o(x).view(np.uint32))
    # and against the oracle's own design of the same class (coefficients agree to ~1e-6): tolerance, not bits
    assert rel_l2(y, getattr(O, name)(*args)(x)) <= 5e-4


@pytest.mark.parametrize("nb,na", [(1, 2), (3, 3), (2, 5), (9, 1), (7, 12), (16, 16), (1, 1)])
def test_transfer_function_iir_bit_exact(cuda, nb, na):
    """CIIRFilter / RIIRFilter: iirfilt_execute_norm, any mix of numerator / denominator lengths up to 16."""
    import scipy.signal as ss
    rng = np.random.default_rng(36)
    # stable denominator: a Butterworth polynomial of the right length (scaled so a[0] != 1 exercises normalisation)
    a = (1.7 * ss.butter(na - 1, 0.3)[1]).astype(np.float32) if na > 1 else np.array([1.7], np.float32)
    b = (0.2 * rng.standard_normal(nb)).astype(np.float32)
    x = crandn(rng, 4099)
    gc, oc = L.CIIRFilter(b, a), O.CIIRFilter(b, a)
    yc = np.concatenate([gc(x[:1000]), gc(x[1000:])])
    assert np.array_equal(yc.view(np.uint32), oc(x).view(np.uint32)), rel_l2(yc, oc(x))
    gr, orr = L.RIIRFilter(b, a, channels=3), O.RIIRFilter(b, a)
    xr = np.stack([x.real, x.imag, x.real[::-1]]).copy()
    yr = gr(xr)
    assert yr.dtype == np.float32 and np.array_equal(yr[0], orr(xr[0])) and np.array_equal(yr[1], yc.imag)
    gr.reset()
    assert np.array_equal(gr(xr[:, :50])[2], O.RIIRFilter(b, a)(xr[2, :50]))


@pytest.mark.parametrize("ntaps,C,n", [(64, 3, 5000), (51, 1, 2049), (1, 2, 100), (200, 1, 9000)])
def test_fir_rrrf(cuda, ntaps, C, n):
    rng = np.random.default_rng(34)
    h = O.firdes_kaiser(ntaps, 0.1, 60.0) if ntaps > 1 else np.array([0.5], np.float32)
    g = L.RealFIRFilter(h, channels=C)
    x = rng.standard_normal((C, n)).astype(np.float32)
    cut = n // 3
    y = np.concatenate([g(x[:, :cut]), g(x[:, cut:])], axis=1) if C > 1 else np.concatenate([g(x[0, :cut]), g(x[0, cut:])]).reshape(1, -1)
    assert y.dtype == np.float32
    for c in range(C):
        assert rel_l2(y[c], O.RealFIRFilter(h)(x[c])) <= TOL_STAGE


def test_real_dc_blocker_and_kaiser_bessel(cuda):
    rng = np.random.default_rng(35)
    x = (rng.standard_normal(6000) + 3.0).astype(np.float32)
    g, o = L.RealDCBlocker(25, 20.0), O.RealDCBlocker(25, 20.0)
    y, yo = g(x), o(x)
    assert rel_l2(y, yo) <= TOL_STAGE and abs(float(np.mean(y[200:]))) < 0.05          # the +3 offset is gone
    g, o = L.RealKaiserBessel(31, 0.1, 40.0), O.RealKaiserBessel(31, 0.1, 40.0)
    y, yo = g(x), o(x)
    assert rel_l2(y, yo) <= TOL_STAGE and abs(float(np.mean(y[200:])) - 3.0) < 0.05     # unit gain at DC


def test_real_chain_am_dcblock_lowpass(cuda):
    """demod.hpp BroadcastAM-style tail built from the widened classes: ampmodem -> RealDCBlocker -> RLowpassIIR."""
    x = am_iq(20000, fs=48000.0, f_off=20.0, noise=0.01, amp=1.0)
    am, dc, lp = L.AmpModem(0.5, "dsb", True), L.RealDCBlocker(), L.RLowpassIIR("butter", 4, 0.1)
    chain = L.Chain(am, dc, lp)
    oam, odc, olp = O.AmpModem(0.5, "dsb", True), O.RealDCBlocker(), O.RealIIRFilter(_sos=lp.sos())
    assert rel_l2(chain(x), olp(odc(oam(x)))) <= TOL_E2E


# ------------------------------------------------------------------------- 8f row 4: resampler variants, Delay
@pytest.mark.parametrize("rate", [0.08, 0.5, 1.0, 2.2, 48000.0 / 600000.0])
def test_default_resamplers_bit_exact(cuda, rate):
    """CResampler / RResampler (resamp_*_create_default): output count, phase and values, streaming."""
    rng = np.random.default_rng(61)
    n = 6001
    x = crandn(rng, n)
    g, o = L.CResampler(rate), O.CResampler(rate)
    cuts = split_points(n, 5, rng)
    y = np.concatenate([g(x[s:e]) for s, e in cuts]); yo = np.concatenate([o(x[s:e]) for s, e in cuts])
    assert y.shape == yo.shape and np.array_equal(y.view(np.uint32), yo.view(np.uint32)) and g.state()[1] == o.phase
    gr, orr = L.RResampler(rate, channels=3), O.RResampler(rate)
    xr = np.stack([x.real, x.imag, x.real[::-1]]).copy()
    yr = np.concatenate([gr(xr[:, s:e]) for s, e in cuts], axis=1)
    assert yr.dtype == np.float32 and np.array_equal(yr[0], y.real) and np.array_equal(yr[1], y.imag)
    assert np.array_equal(yr[2], np.concatenate([orr(xr[2, s:e]) for s, e in cuts]))


def test_real_resampler_rate_property_and_reset(cuda):
    rng = np.random.default_rng(62)
    x = rng.standard_normal(5000).astype(np.float32)
    g, o = L.RealResampler(0.3, Fc=0.12), O.RealResampler(0.3, Fc=0.12)
    assert np.array_equal(g(x[:2000]), o(x[:2000]))
    g.rate = 0.41; o.rate = 0.41
    assert g.rate == 0.41 and np.array_equal(g(x[2000:]), o(x[2000:]))
    g.reset(); o.reset()
    assert np.array_equal(g(x[:100]), o(x[:100]))


@pytest.mark.parametrize("nd", [0, 1, 7, 300])
def test_delay(cuda, nd):
    rng = np.random.default_rng(63)
    g, o = L.Delay(nd), O.Delay(nd)
    z = crandn(rng, 1000); r = rng.standard_normal(1000).astype(np.float32)
    for s, e in [(0, 1), (1, 5), (5, 400), (400, 1000)]:            # the two lines keep separate state
        assert np.array_equal(g(z[s:e]), o(z[s:e])) and np.array_equal(g(r[s:e]), o(r[s:e]))
    g.delay = nd; o.delay = nd                                       # the setter rebuilds and clears both lines
    assert np.array_equal(g(z[:50]), o(z[:50])) and np.array_equal(g(r[:50]), o(r[:50]))
    gb = L.Delay(nd, channels=5)
    zb = crandn(rng, 5, 700)
    yb = np.concatenate([gb(zb[:, :123]), gb(zb[:, 123:])], axis=1)
    ref = np.concatenate([np.zeros((5, nd + 1), np.complex64), zb], axis=1)[:, :700]
    assert np.array_equal(yb, ref)


# pcm_rate > iq_rate (demod.hpp:17-32 takes any ratio): pairs survive only where each resampler yields exactly one sample, and
# the right resampler reads the left one's second output where it yields two (demod.hpp:79-83, the shared y[] slots)
@pytest.mark.parametrize("iq_rate,pcm_rate,n", [(600000.0, 48000.0, 60000), (240000.0, 44100.0, 20001), (96000.0, 96000.0, 3000),
                                                (48000.0, 60000.0, 6001), (44100.0, 48000.0, 5000), (48000.0, 96000.0, 2000), (48000.0, 130000.0, 1500)])
def test_fmstereo_single_channel(cuda, iq_rate, pcm_rate, n):
    x = fm_stereo_iq(n, fs=iq_rate)
    g, o = L.FMStereo(iq_rate, pcm_rate), O.FMStereo(iq_rate, pcm_rate)
    rng = np.random.default_rng(64)
    cuts = split_points(n, 4, rng)
    y = np.concatenate([g(x[s:e]) for s, e in cuts]); yo = np.concatenate([o(x[s:e]) for s, e in cuts])
    assert y.dtype == np.float32 and y.shape == yo.shape and y.size % 2 == 0
    assert _bits_or_close(y, yo, TOL_E2E), rel_l2(y, yo)
    (t, d, pe), (to, do, peo) = g.state(), o.state()
    assert (int(t[0]), int(d[0])) == (to, do) or rel_l2(y, yo) > 0           # PLL words identical when the run was bit-exact
    g.reset(); o.reset()                                                      # resamplers only: the PLL keeps its lock
    assert _bits_or_close(g(x[:2000]), o(x[:2000]), TOL_E2E)


def test_fmstereo_batched(cuda):
    C, n = 70, 24000
    x = np.stack([fm_stereo_iq(n, fl=800.0 + 20 * c, fr=1500.0 + 10 * c, seed=c, phase=0.1 * c) for c in range(C)])
    g = L.FMStereo(channels=C)
    y = np.concatenate([g(x[:, :10001]), g(x[:, 10001:])], axis=1)
    assert y.shape == (C, 2 * 1920)
    for c in (0, 31, 64, 69):
        o = O.FMStereo()
        yo = np.concatenate([o(x[c, :10001]), o(x[c, 10001:])])
        assert _bits_or_close(y[c], yo, TOL_E2E), (c, rel_l2(y[c], yo))


# ------------------------------------------------------------------------- 8f row 3: SSBDemod, HilbertTransform
@pytest.mark.parametrize("band", ["usb", "lsb"])
def test_ssb_demod(cuda, band):
    rng = np.random.default_rng(51)
    C, n = 3, 10000
    x = crandn(rng, C, n)
    g = L.SSBDemod(band, channels=C)
    y = np.concatenate([g(x[:, s:e]) for s, e in split_points(n, 5, rng)], axis=1)      # odd / even cut points: toggle carried
    assert y.dtype == np.float32
    for c in range(C):
        assert rel_l2(y[c], O.SSBDemod(band)(x[c])) <= TOL_STAGE
    g.reset()
    assert rel_l2(g(x[:, :333])[1], O.SSBDemod(band)(x[1, :333])) <= TOL_STAGE
    one = L.SSBDemod(band)
    assert rel_l2(one(x[0]), O.SSBDemod(band)(x[0])) <= TOL_STAGE


@pytest.mark.parametrize("m", [5, 2, 25])
def test_hilbert_transform_both_branches(cuda, m):
    rng = np.random.default_rng(52)
    n = 4001
    g, o = L.HilbertTransform(m, 60.0), O.HilbertTransform(m, 60.0)
    z = crandn(rng, n)
    cuts = [(0, 1), (1, 3), (3, 4), (4, 1200), (1200, 1201), (1201, n)]                 # calls shorter than the filter delay too
    yc = np.concatenate([g(z[s:e]) for s, e in cuts]); yco = np.concatenate([o(z[s:e]) for s, e in cuts])
    assert yc.dtype == np.float32 and np.array_equal(yc, yco)                           # delay and sign only: exact
    r = rng.standard_normal(n).astype(np.float32)
    yr = np.concatenate([g(r[s:e]) for s, e in cuts]); yro = np.concatenate([o(r[s:e]) for s, e in cuts])
    assert yr.dtype == np.complex64 and rel_l2(yr, yro) <= TOL_STAGE
    assert np.array_equal(yr.real, yro.real)                                            # in-phase branch: pure delay, sign, end-of-call zero
    assert g(np.zeros(8)) is None


@pytest.mark.parametrize("typ", ["usb", "lsb"])
@pytest.mark.parametrize("carrier", [False, True])
def test_ampmodem_single_sideband(cuda, typ, carrier):
    """AmpModem(type='usb'|'lsb'): suppressed carrier = Hilbert pair scaled by 0.5 / mod_index; with carrier the PLL loop
    runs first (bit-exact words) and the Hilbert pair and DC blocker follow on the FIR kernel (summation order: 1e-5)."""
    rng = np.random.default_rng(71)
    C, n = 3, 12000
    t = np.arange(n) / 48000.0
    sgn = 1.0 if typ == "usb" else -1.0
    x = np.stack([(0.4 * np.exp(sgn * 2j * np.pi * (900 + 200 * c) * t) + (1.0 if carrier else 0.0) * np.exp(1j * (0.3 + 2 * np.pi * 15 * t))
                   + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64) for c in range(C)])
    g = L.AmpModem(0.5, typ, carrier, channels=C)
    y = np.concatenate([g(x[:, s:e]) for s, e in split_points(n, 4, rng)], axis=1)
    for c in range(C):
        o = O.AmpModem(0.5, typ, carrier)
        yo = o(x[c])
        assert rel_l2(y[c], yo) <= (TOL_E2E if carrier else TOL_STAGE), (c, rel_l2(y[c], yo))
        if carrier:
            t_g, d_g = g.nco_u32()
            assert (int(t_g[c]), int(d_g[c])) == o.nco_u32()
    assert float(np.std(y[0, 4000:])) > 0.3                       # the tone comes through its own side-band
    g.type = "lsb" if typ == "usb" else "usb"                      # property setter rebuilds the object (demod.hpp:251-257)
    assert float(np.std(g(x)[0, 4000:])) < 0.2


# ------------------------------------------------------------------------- 8f row 3: BroadcastAM
@pytest.mark.parametrize("m,n", [(25, 30000), (7, 5001), (40, 3000), (64, 777), (25, 5)])
def test_broadcast_am_single_channel(cuda, m, n):
    x = am_iq(n, fs=48000.0, f_off=35.0, noise=0.01, amp=1.0)
    # the oracle runs the product's DC-block sections: that filter's poles sit 2.6e-3 from z = 1, and the one-ulp
    # differences between two float evaluations of the same design move its output by percent (measured 3.8 %)
    g = L.BroadcastAM(m); o = O.BroadcastAM(m, _dcblock=g.design()[1:])
    y, yo = g(x), o(x)
    assert y.dtype == np.float32 and y.shape == (n,)
    assert _bits_or_close(y, yo), rel_l2(y, yo)
    t, d = g.nco_u32(); to, do = o.nco_u32()
    assert abs(int(d[0]) - do) <= 256 and abs(int(np.int32(np.uint32(int(t[0]) - to)))) <= 1 << 16


def test_broadcast_am_streaming_batched_reset(cuda):
    rng = np.random.default_rng(41)
    C, n = 70, 9000
    x = np.stack([am_iq(n, fs=48000.0, f_off=10.0 + c, phase=0.1 * c, seed=c, noise=0.02, amp=0.5 + 0.01 * c) for c in range(C)])
    a, b = L.BroadcastAM(25, channels=C), L.BroadcastAM(25, channels=C)
    whole = a(x)
    parts = np.concatenate([b(x[:, s:e]) for s, e in split_points(n, 7, rng)], axis=1)
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))           # carried state: bit-identical
    for c in (0, 33, 69):
        assert _bits_or_close(whole[c], O.BroadcastAM(25, _dcblock=a.design()[1:])(x[c])), c
    a.reset()
    assert np.array_equal(a(x[:, :100]), whole[:, :100])


def test_broadcast_am_receiver_chain(cuda):
    """The author's preferred receiver: bandpass -> resample -> [AGC ->] BroadcastAM -> de-emphasis, fused plan
    (time-major hand-off, gain control in place) against the stage-by-stage oracle and the unfused plan."""
    C, n = 67, 65536
    x = np.stack([am_iq(n, f_off=100.0 + 3 * c, phase=0.05 * c, seed=100 + c) for c in range(C)])
    mk = lambda: (L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C), L.ComplexResampler(0.024, Fc=0.024, channels=C),
                  L.AGC(channels=C), L.BroadcastAM(25, channels=C), L.DeemphasisFilter(48000, channels=C))
    st = mk(); st[2].scale = 0.01
    fused = L.Chain(*st)
    assert fused.plan() == "seq[iir4+resamp] -> bam[agc+broadcast_am+deemph]"
    st0 = mk(); st0[2].scale = 0.01
    plain = L.Chain(*st0, fuse=0)
    yf = np.concatenate([fused(x[:, :30000]), fused(x[:, 30000:])], axis=1)
    yp = np.concatenate([plain(x[:, :30000]), plain(x[:, 30000:])], axis=1)
    assert np.array_equal(yf.view(np.uint32), yp.view(np.uint32))
    for c in (0, 66):
        oi, ors, oa, ob, od = (O.ComplexIIRFilter(_sos=st[0].sos()), O.ComplexResampler(0.024, Fc=0.024), O.AGC(),
                               O.BroadcastAM(25, _dcblock=st[3].design()[1:]), O.DeemphasisFilter(48000))
        oa.scale = 0.01
        yo = od(ob(oa(ors(oi(x[c])))))
        assert yo.shape == yf[c].shape and (np.array_equal(yf[c], yo) or rel_l2(yf[c], yo) <= TOL_E2E), (c, rel_l2(yf[c], yo))


# ------------------------------------------------------------------------------------------- a3
def test_resampler_readme_rate_counts_and_values(cuda):
    rng = np.random.default_rng(6)
    g = L.ComplexResampler(rate=48e3 / 2e6, Fc=48e3 / 2e6)
    o = O.ComplexResampler(rate=48e3 / 2e6, Fc=48e3 / 2e6)
    assert g.state()[0] == o.step == 0x29AAAAC0
    counts = []
    for blk in range(9):                                  # 64K blocks: 1573 x7 then 1572
        x = crandn(rng, 65536)
        y, yo = g(x), o(x)
        counts.append(len(y))
        assert len(y) == len(yo)
        assert g.state()[1] == o.phase                    # fixed-point phase bit-exact after every block
        assert np.array_equal(y.view(np.uint32), yo.view(np.uint32)), rel_l2(y, yo)   # cccf arithmetic, same order
    assert counts[:8] == [1573] * 7 + [1572]


@pytest.mark.parametrize("rate", [0.5, 1.7, 0.03, 0.024, 3.0, 0.0101])
def test_resampler_general_rates(cuda, rate):
    rng = np.random.default_rng(7)
    g = L.ComplexResampler(rate=rate, Fc=min(0.45, 0.45 * rate) if rate < 1 else 0.45, channels=2)
    o = [O.ComplexResampler(rate=rate, Fc=min(0.45, 0.45 * rate) if rate < 1 else 0.45) for _ in range(2)]
    for n in (4000, 1, 39, 2500):
        x = crandn(rng, 2, n)
        y = g(x)
        for c in range(2):
            yo = o[c](x[c])
            assert y.shape[1] == len(yo)
            assert np.array_equal(y[c].view(np.uint32), yo.view(np.uint32)), (rate, n, c)
        assert g.state()[1] == o[0].phase


def test_resampler_set_rate_and_reset(cuda):
    rng = np.random.default_rng(8)
    g, o = L.ComplexResampler(0.024, Fc=0.024), O.ComplexResampler(0.024, Fc=0.024)
    x = crandn(rng, 5000)
    assert rel_l2(g(x), o(x)) <= TOL_STAGE
    g.rate = 0.02; o.rate = 0.02                          # set_rate keeps the window (resampler.hpp:151-154)
    assert g.state()[0] == o.step
    assert rel_l2(g(x), o(x)) <= TOL_STAGE
    g.reset(); o.reset()
    assert rel_l2(g(x), o(x)) <= TOL_STAGE


# ------------------------------------------------------------------------------------------- a4
def test_nco_phase_bit_exact_and_mix(cuda):
    rng = np.random.default_rng(9)
    g, o = L.NCO(), O.NCO()
    g.freq = 0.3; o.freq = 0.3
    g.phase = 1.0; o.phase = 1.0
    x = crandn(rng, 10001)
    yu, you = g.mix_up(x), o.mix_up(x)
    assert np.array_equal(yu.view(np.uint32), you.view(np.uint32))
    t, d = g.u32(); assert (int(t[0]), int(d[0])) == (o.theta_u32, o.dtheta_u32)
    yd, yod = g.mix_down(x), o.mix_down(x)
    assert np.array_equal(yd.view(np.uint32), yod.view(np.uint32))
    for df in (-1e-3, 1e-9, -1e-9, 0.5):                 # constrain() edge cases incl. the tiny-negative wrap
        g.adjust_frequency(df); o.adjust_frequency(df)
        g.adjust_phase(-df); o.adjust_phase(-df)
    g.set_pll_bandwidth(0.05); o.set_pll_bandwidth(0.05); g.pll_step(0.01); o.pll_step(0.01)
    t, d = g.u32(); assert (int(t[0]), int(d[0])) == (o.theta_u32, o.dtheta_u32)
    assert abs(g.freq - o.freq) < 1e-6 and abs(g.phase - o.phase) < 1e-6


def test_nco_batched_per_channel_closed_form(cuda):
    C, n = 70, 3000
    g = L.NCO(channels=C)
    f = (2 * np.pi * (0.05 + 0.4 * np.arange(C) / C)).astype(np.float32)
    g.set_frequencies(f)
    x = np.ones((C, n), np.complex64)
    g.set_direction(True); y = L._Stage.__call__(g, x)
    t, d = g.u32()
    assert np.array_equal(t, (d.astype(np.uint64) * n % (1 << 32)).astype(np.uint32))   # theta_n = n * d_theta mod 2^32
    for c in (0, 33, 69):
        o = O.NCO(); o.freq = float(f[c])
        assert o.dtheta_u32 == int(d[c])
        assert np.array_equal(y[c].view(np.uint32), o.mix_down(x[c]).view(np.uint32))


def test_vco_type(cuda):
    rng = np.random.default_rng(10)
    g, o = L.NCO("vco"), O.NCO("vco")
    g.freq = 0.123; o.freq = 0.123
    x = crandn(rng, 4000)
    assert rel_l2(g(x), o(x)) <= TOL_STAGE


# ------------------------------------------------------------------------------------------- a5
def test_agc_default_and_readme_settings(cuda):
    rng = np.random.default_rng(11)
    x = crandn(rng, 30000, scale=0.05)
    x[15000:] *= 20
    for scale, bw in ((1.0, 1e-2), (0.01, 1e-2), (1.0, 1e-3)):
        g, o = L.AGC(), O.AGC()
        g.lock = False; o.lock = False
        g.scale = scale; o.scale = scale; g.bandwidth = bw; o.bandwidth = bw
        y, yo = g(x), o(x)
        assert rel_l2(y, yo) <= TOL_STAGE
        assert abs(g.gain / o.gain - 1) < 1e-5 and abs(g.level_dB - o.level_dB) < 1e-3


def test_agc_precision_modes(cuda):
    """'exact' reproduces the oracle's gain loop bit for bit; 'fast' (single precision, lg2.approx) stays within 1e-6 of it --
    a tenth of the per-stage tolerance -- over acquisition, a 26 dB level step and the settled loop, at three bandwidths."""
    rng = np.random.default_rng(14)
    x = crandn(rng, 40000, scale=0.05)
    x[20000:] *= 20
    t = np.arange(40000) / 48000.0
    am = (0.1 * (1 + 0.5 * np.sin(2 * np.pi * 1000 * t)) * np.exp(2j * np.pi * 200 * t)).astype(np.complex64) + crandn(rng, 40000, scale=0.01)
    cw = (0.3 * np.exp(1j * (0.3 * np.arange(40000)))).astype(np.complex64)          # noise-free: liquid's dead zone around y2' = 1
    for sig in (x, am, cw):
        for bw in (1e-2, 1e-3, 0.2, 1.0):
            ge, gf, o = L.AGC(), L.AGC(), O.AGC()
            assert ge.precision == "auto"
            ge.precision = "exact"; gf.precision = "fast"
            assert (ge.precision, gf.precision) == ("exact", "fast")
            for q in (ge, gf, o):
                q.lock = False; q.scale = 0.5; q.bandwidth = bw
            yo = np.concatenate([o(sig[:17001]), o(sig[17001:])])
            ye = np.concatenate([ge(sig[:17001]), ge(sig[17001:])])
            yf = np.concatenate([gf(sig[:17001]), gf(sig[17001:])])
            assert np.array_equal(ye.view(np.uint32), yo.view(np.uint32)), bw
            assert rel_l2(yf, yo) <= 1e-6, (bw, rel_l2(yf, yo))
            assert abs(gf.gain / o.gain - 1) < 2e-6
    with pytest.raises(ValueError):
        L.AGC().precision = "sloppy"


def test_agc_auto_precision_follows_the_chain(cuda):
    """AUTO: single precision only where a FreqDem follows and no carrier PLL does; in front of AmpModem (the README chain)
    and on its own the AGC stays bit-identical to the oracle."""
    rng = np.random.default_rng(15)
    C, n = 40, 6000
    x = np.stack([am_iq(n, fs=48000.0, f_off=30.0 + c, phase=0.1 * c, seed=900 + c, noise=0.02, amp=0.3) for c in range(C)])
    # AGC -> AmpModem: exact
    chain = L.Chain(L.AGC(channels=C), L.AmpModem(0.5, "dsb", True, channels=C))
    y = chain(x)
    oa, om = O.AGC(), O.AmpModem(0.5, "dsb", True)
    assert np.array_equal(y[7].view(np.uint32), om(oa(x[7])).view(np.uint32))
    # AGC alone in a chain: exact
    y = L.Chain(L.AGC(channels=C))(x)
    assert np.array_equal(y[3].view(np.uint32), O.AGC()(x[3]).view(np.uint32))
    # AGC -> FreqDem: the fast loop -- not bit-identical to the exact chain, and well inside the tolerance
    fast = L.Chain(L.AGC(channels=C), L.FreqDem(0.2, channels=C))
    ex_agc = L.AGC(channels=C); ex_agc.precision = "exact"
    exact = L.Chain(ex_agc, L.FreqDem(0.2, channels=C))
    yf, ye = fast(x), exact(x)
    yo = O.FreqDem(0.2)(O.AGC()(x[5]))
    assert np.linalg.norm(ye[5] - yo) / np.sqrt(n) <= 2.5 * TOL_STAGE
    assert np.linalg.norm(yf[5] - yo) / np.sqrt(n) <= 2.5 * TOL_STAGE     # phase-difference scale: |y| <= 1 / (2 kf) = 2.5
    assert abs(fast.stages[0].gains()[5] / exact.stages[0].gains()[5] - 1) < 2e-6


def test_agc_lock_and_properties(cuda):
    rng = np.random.default_rng(12)
    x = crandn(rng, 5000, scale=0.2)
    g, o = L.AGC(), O.AGC()
    g.gain = 3.0; o.gain = 3.0
    g.lock = True; o.lock = True
    assert rel_l2(g(x), o(x)) <= 1e-6                     # locked: y = x*g, no output scale applied
    g.level = 0.5; o.level = 0.5
    assert abs(g.gain - o.gain) < 1e-6
    g.level_dB = -30.0; o.level_dB = -30.0
    assert abs(g.gain / o.gain - 1) < 1e-6
    g.reset(); o.reset()
    assert g.gain == o.gain == 1.0
    assert rel_l2(g(x), o(x)) <= TOL_STAGE


def test_agc_squelch_states_and_on_rise(cuda):
    rng = np.random.default_rng(13)
    n = 6000
    env = np.concatenate([np.full(1500, 1e-3), np.full(2000, 1.0), np.full(2500, 1e-3)])
    x = (env * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)).astype(np.complex64)
    g, o = L.AGC(), O.AGC()
    fired = []
    g.onRise = lambda: fired.append(1)
    for q in (g, o):
        q.bandwidth = 0.05; q.squelch = True; q.threshold = -30.0
    g.set_timeout(100); o.set_timeout(100)
    y, yo = g(x), o(x)
    assert np.array_equal(y == 0, yo == 0)                # zeroed in ENABLED / SIGNALLO (agc.hpp:124-125)
    assert rel_l2(y, yo) <= TOL_STAGE
    assert g.status == o.status
    assert len(fired) == len(o.rise_indices) >= 1


# ------------------------------------------------------------------------------------------- a6
@pytest.mark.parametrize("carrier", [True, False])
def test_ampmodem_dsb(cuda, carrier):
    n = 20000
    t = np.arange(n) / 48e3
    m = 0.6 * np.sin(2 * np.pi * 1000 * t) + 0.4 * np.sin(2 * np.pi * 2500 * t)
    rng = np.random.default_rng(14)
    if carrier:
        x = 0.01 * (1 + 0.5 * m) * np.exp(1j * (2 * np.pi * 5.0 * t + 0.4))
    else:
        x = 0.5 * m * np.exp(1j * (2 * np.pi * 2.0 * t + 0.2))
    x = (x + 1e-4 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    g, o = L.AmpModem(0.5, "dsb", carrier), O.AmpModem(0.5, "dsb", carrier)
    lp, dc = g.taps(); lpo, dco = o.taps()
    assert np.array_equal(lp, lpo) and np.array_equal(dc, dco)
    ys, yos = [], []
    for s, e in split_points(n, 5, rng):                  # state carried across ragged calls
        ys.append(g(x[s:e])); yos.append(o(x[s:e]))
        tg, dg = g.nco_u32()
        assert (int(tg[0]), int(dg[0])) == o.nco_u32()    # PLL phase and frequency words bit-exact
    y, yo = np.concatenate(ys), np.concatenate(yos)
    assert y.dtype == np.float32
    assert rel_l2(y, yo) <= TOL_STAGE
    g.reset(); o.reset()
    assert rel_l2(g(x[:3000]), o(x[:3000])) <= TOL_STAGE
    g.modulation = 0.8; o.modulation = 0.8                # setter rebuilds the modem (demod.hpp:258-262)
    assert rel_l2(g(x[:3000]), o(x[:3000])) <= TOL_STAGE


# ------------------------------------------------------------------------------------------- a7 / a8
def test_freqdem(cuda):
    x = fm_iq(50000)
    g, o = L.FreqDem(0.1), O.FreqDem(0.1)
    y = np.concatenate([g(x[:20001]), g(x[20001:])]); yo = o(x)
    assert rel_l2(y, yo) <= TOL_STAGE
    assert abs(np.mean(y[1000:] ** 2) - 0.5) < 0.05      # recovers the unit-amplitude 1 kHz tone


def test_deemphasis_bit_exact(cuda):
    rng = np.random.default_rng(15)
    x = rng.standard_normal(10000).astype(np.float32)
    g, o = L.DeemphasisFilter(48000), O.DeemphasisFilter(48000)
    assert g.coeffs() == O.DeemphasisFilter.coeffs(48000)
    y = np.concatenate([g(x[:4097]), g(x[4097:])])
    assert np.array_equal(y.view(np.uint32), o(x).view(np.uint32))
    gb = L.DeemphasisFilter(48000, channels=5)
    yb = gb(np.tile(x, (5, 1)))
    assert all(np.array_equal(yb[c], y) for c in range(5))


# ------------------------------------------------------------------------------------------- chains
def _amradio_stages(mod, channels=1):
    return (mod.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / 2e6, **({"channels": channels} if mod is L else {})),
            mod.ComplexResampler(rate=48e3 / 2e6, Fc=48e3 / 2e6, **({"channels": channels} if mod is L else {})))


class _Radio:
    """README.md:41-58 AMRadio.__init__/__call__, parameterised by the module that provides the classes."""

    def __init__(self, mod, channels=None):
        kw = {} if channels is None else {"channels": channels}
        self.bandpass = mod.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / 2000000, **kw)
        self.resample = mod.ComplexResampler(rate=48000 / 2000000, Fc=48000 / 2000000, **kw)
        self.am = mod.AmpModem(modulation=0.5, type="dsb", carrier=True, **kw)
        self.audio_filter = mod.DeemphasisFilter(48000, **kw)
        self.agc = mod.AGC(**kw)
        self.agc.lock = False
        self.agc.scale = 0.01

    def stages(self):
        return (self.bandpass, self.resample, self.agc, self.am, self.audio_filter)

    def __call__(self, iq):
        return self.audio_filter(self.am(self.agc(self.resample(self.bandpass(iq)))))


def _oracle_radio_with_product_coeffs(gr):
    r = _Radio(O)
    r.bandpass = O.ComplexIIRFilter(_sos=gr.bandpass.sos())
    return r


def test_readme_amradio_drop_in_config1(cuda):
    """BASELINE config 1 (shortened): README AMRadio, one channel, 64K blocks, state carried."""
    n, blk = 10 * 65536, 65536
    x = am_iq(n)
    gr = _Radio(L); orr = _oracle_radio_with_product_coeffs(gr)
    pcm = np.concatenate([gr(x[i:i + blk]) for i in range(0, n, blk)])
    ref = np.concatenate([orr(x[i:i + blk]) for i in range(0, n, blk)])
    assert pcm.shape == ref.shape and pcm.dtype == np.float32
    assert rel_l2(pcm, ref) <= TOL_E2E
    seg = pcm[6000:] * np.hanning(len(pcm) - 6000)       # the two audio tones come out on top
    spec = np.abs(np.fft.rfft(seg)); f = np.fft.rfftfreq(len(seg), 1 / 48e3)
    top = sorted(f[np.argsort(spec)[-4:]])
    assert abs(top[0] - 1000) < 5 and abs(top[-1] - 2500) < 5


@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_fused_chain_matches_stagewise(cuda, fuse):
    n, blk = 3 * 65536 + 1234, 65536
    x = am_iq(n, seed=5)
    a, b = _Radio(L), _Radio(L)
    chain = L.Chain(*b.stages(), fuse=fuse)
    expect = {0: 6, 1: 4, 2: 1}[fuse]        # level 1: tap stream + full-rate kernel, AGC in place on the hand-off, AM tail
    ys, yc = [], []
    for i in range(0, n, blk):
        ys.append(a(x[i:i + blk])); yc.append(chain(x[i:i + blk]))
        if fuse == 1:                        # the library names what it dispatched (lqb_chain_last_kernels)
            k = chain.last_kernels()         # one channel: 4 lane pairs; long blocks are cut into 6 time slices, the tail of one under the front of the next
            assert k[:2] == ["tapstream_kernel", "lanes_kernel<1,4>"] and k[2].startswith("agc_tmajor_kernel") and k[3].startswith("amtail8_kernel")
            sliced = any("time slices" in s for s in k)
            assert sliced == (len(x[i:i + blk]) >= 16384)
            assert chain.last_launches() == (6 * expect if sliced else expect)
        else:
            assert chain.last_launches() == expect
    ys, yc = np.concatenate(ys), np.concatenate(yc)
    assert np.array_equal(ys.view(np.uint32), yc.view(np.uint32)), rel_l2(yc, ys)
    assert b.resample.state() == a.resample.state()


def test_batched_amradio_vs_oracle_subset(cuda):
    """BASELINE config 5 in miniature: many channels, device-generated input, oracle on a channel subset."""
    C, n = 200, 32768
    r = _Radio(L, channels=C)
    chain = L.Chain(*r.stages())
    xb = L.DeviceBuffer(C * n * 8)
    orr = {}
    outs = []
    for blk in range(3):
        L.synth_fill(0, xb.ptr.value, C, n, n0=blk * n)
        x = xb.download((C, n), np.complex64)
        y = chain(x)
        outs.append(y)
        for c in (0, 1, 63, 64, 127, 199):
            if c not in orr:
                orr[c] = [_oracle_radio_with_product_coeffs(r), []]
            orr[c][1].append(orr[c][0](x[c]))
    y = np.concatenate(outs, axis=1)
    for c, (_, parts) in orr.items():
        assert rel_l2(y[c], np.concatenate(parts)) <= TOL_E2E, c
    assert np.std(y[0][1500:]) > 1e-4                    # there is audio


def test_config3_nco_resampler(cuda):
    C, n = 96, 65536
    nco = L.NCO(channels=C); rs = L.ComplexResampler(0.024, Fc=0.024, channels=C)
    f = (2 * np.pi * (0.05 + 0.4 * np.arange(C) / 4096)).astype(np.float32)
    nco.set_frequencies(f); nco.set_direction(True)
    chain = L.Chain(nco, rs)
    assert chain.plan() == "par[nco+resamp]"             # 96 channels: time-parallel, mixer applied while staging
    xb = L.DeviceBuffer(C * n * 8)
    refs = {c: (O.NCO(), O.ComplexResampler(0.024, Fc=0.024)) for c in (0, 50, 95)}
    for c, (on, _) in refs.items():
        on.freq = float(f[c])
    for blk in range(3):
        L.synth_fill(2, xb.ptr.value, C, n, n0=blk * n)
        x = xb.download((C, n), np.complex64)
        y = chain(x)
        for c, (on, ors) in refs.items():
            yo = ors(on.mix_down(x[c]))
            assert y.shape[1] == len(yo)
            assert rel_l2(y[c], yo) <= TOL_STAGE
        t, _ = nco.u32()
        assert all(int(t[c]) == refs[c][0].theta_u32 for c in refs)   # oscillator phase bit-exact
        assert rs.state()[1] == refs[0][1].phase


def test_nco_resampler_sequential_equals_time_parallel(cuda):
    """The channel-parallel sequential kernel (many channels) and the time-parallel kernels (few) agree bit for bit."""
    C, n = 16384 + 64, 4096
    nco = L.NCO(channels=C); rs = L.ComplexResampler(0.024, Fc=0.024, channels=C)
    f = (2 * np.pi * (0.05 + 0.4 * (np.arange(C) % 4096) / 4096)).astype(np.float32)
    nco.set_frequencies(f); nco.set_direction(True)
    big = L.Chain(nco, rs)
    assert big.plan() == "seq[nco+resamp]"
    xb = L.DeviceBuffer(C * n * 8)
    small = {c: L.Chain(L.NCO(), L.ComplexResampler(0.024, Fc=0.024)) for c in (0, 4095, C - 1)}
    for c, ch in small.items():
        ch.stages[0].freq = float(f[c]); ch.stages[0].set_direction(True)
        assert ch.plan() == "par[nco+resamp]"
    for blk in range(3):
        L.synth_fill(2, xb.ptr.value, C, n, n0=blk * n)
        x = xb.download((C, n), np.complex64)
        y = big(x)
        for c, ch in small.items():
            yc = ch(x[c])
            assert np.array_equal(y[c].view(np.uint32), yc.view(np.uint32)), (blk, c)
    t, _ = nco.u32()
    for c, ch in small.items():
        assert int(t[c]) == int(ch.stages[0].u32()[0][0])


def test_config4_iir_agc_fm(cuda):
    C, n = 80, 20000
    iir = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C)
    agc = L.AGC(channels=C); fm = L.FreqDem(0.1, channels=C)
    chain = L.Chain(iir, agc, fm)
    assert chain.plan() == "seq[iir4+agc+freqdem]"
    xb = L.DeviceBuffer(C * n * 8)
    refs = {c: (O.ComplexIIRFilter(_sos=iir.sos()), O.AGC(), O.FreqDem(0.1)) for c in (0, 40, 79)}
    for blk in range(2):
        L.synth_fill(3, xb.ptr.value, C, n, n0=blk * n)
        x = xb.download((C, n), np.complex64)
        y = chain(x)
        assert y.dtype == np.float32 and y.shape == (C, n)
        for c, (oi, oa, of) in refs.items():
            yo = of(oa(oi(x[c])))
            # FM demod of a noisy narrow-band signal: compare on phase-difference scale (|y| <= 1/(2 kf) = 5)
            assert np.linalg.norm(y[c] - yo) / np.sqrt(n) <= 5 * TOL_E2E, c


def _to_i16(rng, C, n, amp=6000):
    a = rng.integers(-amp, amp, size=(C, 2 * n)).astype(np.int16)
    t = np.arange(n)
    car = (8000 * (1 + 0.5 * np.sin(2 * np.pi * 1000 * t / 2e6)) * np.exp(2j * np.pi * 300 * t / 2e6))
    a[:, 0::2] += car.real.astype(np.int16); a[:, 1::2] += car.imag.astype(np.int16)
    return a


def test_bytes_to_iq_all_values(cuda):
    s16 = np.arange(-32768, 32768, dtype=np.int32)
    raw = np.stack([s16, s16[::-1]], axis=1).astype("<i2").tobytes()
    assert np.array_equal(L.bytes_to_iq(raw).view(np.uint32), O.bytes_to_iq(raw).view(np.uint32))


@pytest.mark.parametrize("C,n", [(1, 65536), (70, 8192), (70, 8190), (60000, 2048)])
def test_int16_ingest_fused_equals_convert_then_filter(cuda, C, n):
    """Interleaved int16 I/Q fed straight to the chain == bytes_to_iq on the host followed by the same chain."""
    rng = np.random.default_rng(31)
    raw = _to_i16(rng, min(C, 64), n)
    raw = np.tile(raw, (C // raw.shape[0] + 1, 1))[:C].copy()
    a, b = _Radio(L, channels=C), _Radio(L, channels=C)
    ca, cb = L.Chain(*a.stages()), L.Chain(*b.stages())
    for blk in range(2):
        x = np.stack([O.bytes_to_iq(raw[c].tobytes()) for c in range(min(C, 64))])
        x = np.tile(x, (C // x.shape[0] + 1, 1))[:C].copy()
        ya = ca(raw if C > 1 else raw[0])
        yb = cb(x if C > 1 else x[0])
        assert np.array_equal(np.asarray(ya).view(np.uint32), np.asarray(yb).view(np.uint32)), (C, n, blk)
    if C == 70:
        o = _oracle_radio_with_product_coeffs(a); ref = None
        for blk in range(2):
            ref = o(O.bytes_to_iq(raw[5].tobytes()))
        assert rel_l2(np.asarray(ya)[5], ref) <= TOL_E2E


def test_execute_dev_and_chunked_host_agree(cuda):
    """Device-pointer entry point == host entry point (which chunks channels over three streams)."""
    C, n = 300, 65536                                     # 157 MB of input -> 3 chunks
    r1, r2 = _Radio(L, channels=C), _Radio(L, channels=C)
    c1, c2 = L.Chain(*r1.stages()), L.Chain(*r2.stages())
    xb = L.DeviceBuffer(C * n * 8)
    L.synth_fill(0, xb.ptr.value, C, n)
    x = xb.download((C, n), np.complex64)
    y1 = c1(x)
    n_out = c2.out_len(n)
    yb = L.DeviceBuffer(C * n_out * 4)
    got = c2.execute_dev(xb.ptr.value, n, yb.ptr.value, n_out, 0)
    L.synchronize()
    y2 = yb.download((C, got), np.float32)
    assert got == n_out == y1.shape[1]
    assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32))


def test_host_call_time_sliced_equals_channel_chunked(cuda, monkeypatch):
    """Large host-pointer calls are pipelined in time slices (all channels per slice, state carried between slices);
    the older channel-chunk pipeline must give the same bits -- c64 and int16 input."""
    C, n = 192, 65536 + 4096
    x = np.stack([am_iq(n, f_off=50.0 + c, phase=0.01 * c, seed=c % 7) for c in range(C)])
    iq16 = np.empty((C, n, 2), np.int16)
    iq16[..., 0] = np.clip(np.round(x.real * 20000), -32767, 32767); iq16[..., 1] = np.clip(np.round(x.imag * 20000), -32767, 32767)
    outs = {}
    for mode in ("slice", "chunk"):
        monkeypatch.setenv("LQB_HOST_MODE", mode)
        st = (L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C), L.ComplexResampler(0.024, Fc=0.024, channels=C), L.AGC(channels=C),
              L.AmpModem(0.5, "dsb", True, channels=C), L.DeemphasisFilter(48000, channels=C))
        st[2].scale = 0.01
        ch = L.Chain(*st)
        a = ch(x); b = ch(x[:, :40000])                         # second call: carried state after a sliced call
        ch.reset()
        outs[mode] = (a, b, ch(iq16.reshape(C, 2 * n)))
    for u, v in zip(outs["slice"], outs["chunk"]):
        assert u.shape == v.shape and np.array_equal(u.view(np.uint32), v.view(np.uint32))


def test_full_rate_stages_wide_batches_tma(cuda):
    """ComplexIIRFilter and NCO alone at a channel count that selects 32 channels per warp: TMA tile loads and bulk
    tensor stores, ragged last box (C not a multiple of 32) and ragged last tile (n not a multiple of 16)."""
    rng = np.random.default_rng(81)
    C, n = 56832 + 13, 1000 + 6
    base = crandn(rng, 64, n)
    x = np.tile(base, (C // 64 + 1, 1))[:C].copy()
    iir = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C)
    ys = [iir(x), iir(x)]
    for c in (0, 31, 32, C - 14, C - 1):
        o = O.ComplexIIRFilter(_sos=iir.sos())
        for k in range(2):
            assert np.array_equal(ys[k][c].view(np.uint32), o(x[c]).view(np.uint32)), (c, k)
    nco = L.NCO(channels=C); nco.set_frequencies((0.2 + 1e-5 * np.arange(C)).astype(np.float32))
    y = nco.mix_down(x)
    for c in (0, 33, C - 1):
        o = O.NCO(); o.freq = float(np.float32(0.2 + 1e-5 * c))
        assert np.array_equal(y[c].view(np.uint32), o.mix_down(x[c]).view(np.uint32)), c


@pytest.mark.parametrize("C,tail", [(33000, "am"), (200, "am"), (33000, "bam")])
def test_overlapped_device_calls_equal_serial(cuda, C, tail):
    """lqb_chain_set_overlap: block k's decimated-rate tail on the chain's own stream under block k+1's front -- same
    words out as the serial calls, block after block with the state carried (two hand-off buffers in rotation)."""
    n, nblk = 4096, 5
    def radio():
        r = _Radio(L, channels=C)
        st = list(r.stages())
        if tail == "bam":
            st[3] = L.BroadcastAM(25, channels=C)
        return L.Chain(*st)
    serial, over = radio(), radio()
    over.set_overlap(True)
    xb = L.DeviceBuffer(C * n * 8)
    cap = serial.out_len(n) + 2
    ys, yo = L.DeviceBuffer(C * cap * 4), [L.DeviceBuffer(C * cap * 4) for _ in range(nblk)]
    st = L.stream_create(True)
    outs_s, lens = [], []
    for blk in range(nblk):
        L.synth_fill(0, xb.ptr.value, C, n, n0=blk * n)
        L.synchronize()
        got = serial.execute_dev(xb.ptr.value, n, ys.ptr.value, cap, 0)
        L.synchronize()
        outs_s.append(ys.download((C * cap,), np.float32)[:C * got].reshape(C, got)); lens.append(got)     # rows are dense: [C][n_out]
        assert over.execute_dev(xb.ptr.value, n, yo[blk].ptr.value, cap, st) == got
        L.stream_synchronize(st)            # the input buffer is refilled next; the TAIL of this block may still be running
    over.wait(st)
    L.stream_synchronize(st)
    for blk in range(nblk):
        y = yo[blk].download((C * cap,), np.float32)[:C * lens[blk]].reshape(C, lens[blk])
        d = np.argwhere(y.view(np.uint32) != outs_s[blk].view(np.uint32))
        assert len(d) == 0, (blk, len(d), d[:4].tolist(), int(d[:, 0].min()), int(d[:, 0].max()))
    # a serial call on the overlapped chain joins the outstanding tails by itself
    over.set_overlap(False)
    L.synth_fill(0, xb.ptr.value, C, n, n0=nblk * n); L.synchronize()
    g1 = serial.execute_dev(xb.ptr.value, n, ys.ptr.value, cap, 0); L.synchronize()
    a = ys.download((C * cap,), np.float32)[:C * g1]
    g2 = over.execute_dev(xb.ptr.value, n, yo[0].ptr.value, cap, 0); L.synchronize()
    assert g1 == g2 and np.array_equal(a.view(np.uint32), yo[0].download((C * cap,), np.float32)[:C * g2].view(np.uint32))
    L.stream_destroy(st)


@pytest.mark.parametrize("C,n,order", [(1, 37, 2), (33, 4101, 4), (70, 8000, 6), (200, 16, 8), (129, 2048, 8)])
def test_pipeline_kernel_equals_one_warp_kernel(cuda, C, n, order, monkeypatch):
    """Config 4's chain takes the three-warp pipeline (pipe.cu); the one-warp seq_kernel (LQB_NO_PIPE) evaluates the same
    operations -- bit-identical outputs and carried state over split calls, ragged lengths (no tensor map: odd n),
    partial tiles and channel counts that do not fill a warp; both within tolerance of the oracle."""
    rng = np.random.default_rng(77)
    x = np.stack([fm_iq(n, seed=300 + c) * np.float32(0.05 + 0.9 * c / max(C - 1, 1)) for c in range(C)]).astype(np.complex64)
    def make():
        iir = L.ComplexIIRFilter("cheby2", order=order, Fc=0.05, channels=C)
        return L.Chain(iir, L.AGC(channels=C), L.FreqDem(0.1, channels=C)), iir
    cuts = split_points(n, 3, rng)
    monkeypatch.delenv("LQB_NO_PIPE", raising=False)
    pipe, iir = make()
    yp = np.concatenate([pipe(x[:, s:e] if C > 1 else x[0, s:e]).reshape(C, -1) for s, e in cuts], axis=1)
    monkeypatch.setenv("LQB_NO_PIPE", "1")
    seq, _ = make()
    ys = np.concatenate([seq(x[:, s:e] if C > 1 else x[0, s:e]).reshape(C, -1) for s, e in cuts], axis=1)
    monkeypatch.delenv("LQB_NO_PIPE", raising=False)
    assert yp.shape == (C, n) and np.array_equal(yp.view(np.uint32), ys.view(np.uint32))
    assert np.array_equal(pipe.stages[1].gains().view(np.uint32), seq.stages[1].gains().view(np.uint32))
    for c in sorted({0, C // 2, C - 1}):
        yo = O.FreqDem(0.1)(O.AGC()(O.ComplexIIRFilter(_sos=iir.sos())(x[c])))
        assert np.linalg.norm(yp[c] - yo) / np.sqrt(n) <= 5 * TOL_E2E, c


def test_amtail_eight_lane_kernel_equals_thread_per_channel(cuda, monkeypatch):
    """Few channels run the AM tail with eight lanes per channel (am.cu amtail8_kernel); LQB_NO_AMTAIL8 selects the
    one-thread-per-channel kernel.  Same arithmetic: bit-identical audio and PLL words, ragged calls, carried state."""
    C, n = 37, 3 * 4096 + 123
    rng = np.random.default_rng(88)
    x = np.stack([am_iq(n, seed=300 + c, f_off=150.0 + 11 * c) for c in range(C)])
    cuts = split_points(n, 5, rng)
    outs, words = [], []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("LQB_NO_AMTAIL8", env)
        else:
            monkeypatch.delenv("LQB_NO_AMTAIL8", raising=False)
        r = _Radio(L, channels=C)
        ch = L.Chain(*r.stages())
        ys = [ch(np.ascontiguousarray(x[:, s:e])) for s, e in cuts]
        assert ("amtail8_kernel" in ch.last_kernels()) == (env is None)
        outs.append(np.concatenate(ys, axis=1)); words.append(r.am.nco_u32())
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(words[0][0], words[1][0]) and np.array_equal(words[0][1], words[1][1])


@pytest.mark.parametrize("C", [5, 300, 2500])
def test_time_slices_on_disjoint_sms_equal_plain_call(cuda, monkeypatch, C):
    """Up to 6400 channels a long call is cut into six time slices: the front of slice j + 1 runs on the SMs its lone warps fill,
    gain loop and demodulator of slice j on the others (capi.cu run_timepipe, green-context SM partition; ordinary streams for
    <= 1024 channels when the driver lacks the API).  State carries between slices as between calls: bit-identical to the
    unsliced call (LQB_NO_TIMEPIPE=1) and, with LQB_NO_PARTITION=1, to the slices on shared SMs; carried over a second call.
    Device-pointer calls (large host-pointer calls take the host path's own time slices instead)."""
    n1, n2 = 32768, 4098
    base = np.stack([am_iq(n1 + n2, seed=40 + c, f_off=100.0 + 3 * c) for c in range(min(C, 125))])
    x = np.tile(base, ((C + base.shape[0] - 1) // base.shape[0], 1))[:C]          # (2500 channels: 125 signals, twenty times)
    parts = [np.ascontiguousarray(x[:, :n1]), np.ascontiguousarray(x[:, n1:])]
    xb = [L.DeviceBuffer(p.nbytes) for p in parts]
    for b, p in zip(xb, parts):
        b.upload(p.view(np.uint8).ravel())
    outs, notes = [], []
    for env in ({"LQB_NO_TIMEPIPE": "1"}, {}, {"LQB_NO_PARTITION": "1", "LQB_TIMEPIPE_MAX": "100000"}):
        for k in ("LQB_NO_TIMEPIPE", "LQB_NO_PARTITION", "LQB_TIMEPIPE_MAX"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        r = _Radio(L, channels=C); ch = L.Chain(*r.stages())
        ys = []
        for b, nn in zip(xb, (n1, n2)):
            cap = ch.out_len(nn)
            yb = L.DeviceBuffer(max(16, C * cap * 4))
            got = ch.execute_dev(b.ptr.value, nn, yb.ptr.value, cap, 0); L.synchronize()
            assert got == cap
            if nn == n1:
                notes.append(ch.last_kernels()[-1])
            ys.append(yb.download((C, cap), np.float32))
        outs.append(np.concatenate(ys, axis=1))
    assert "time slices" not in notes[0] and "time slices" in notes[2] and "disjoint" not in notes[2], notes
    assert "time slices" in notes[1] or C > 1024, notes      # (above 1024 channels the default slices only with an SM partition)
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert np.array_equal(outs[0].view(np.uint32), outs[2].view(np.uint32))


def test_pageable_input_staged_by_host_threads(cuda, monkeypatch):
    """A plain numpy array is pageable memory; the time-sliced host path copies it through pinned bounce buffers with a few host
    threads (capi.cu staged_h2d) instead of leaving it to the driver's single-threaded staging.  Same bytes reach the device:
    the audio is bit-identical to the driver-staged call (LQB_NO_STAGING=1) for complex64 and for int16 I/Q input, with a row
    count that does not divide among the threads and rows that do not fill a bounce buffer evenly."""
    C, n = 141, 3 * 32768
    x = np.stack([am_iq(n, seed=700 + c, f_off=90.0 + 7 * c) for c in range(C)])
    xi = np.empty((C, 2 * n), np.int16)
    xi[:, 0::2] = np.clip(np.round(x.real * 32767), -32767, 32767); xi[:, 1::2] = np.clip(np.round(x.imag * 32767), -32767, 32767)
    outs = {}
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("LQB_NO_STAGING", env)
        else:
            monkeypatch.delenv("LQB_NO_STAGING", raising=False)
        for name, arr in (("c64", x), ("i16", xi)):
            r = _Radio(L, channels=C); ch = L.Chain(*r.stages())
            y = np.concatenate([ch(np.ascontiguousarray(arr[:, :arr.shape[1] // 3 * 2])), ch(np.ascontiguousarray(arr[:, arr.shape[1] // 3 * 2:]))], axis=1)
            outs[(env, name)] = y
    for name in ("c64", "i16"):
        assert np.array_equal(outs[(None, name)].view(np.uint32), outs[("1", name)].view(np.uint32)), name
    ref = _oracle_radio_with_product_coeffs(_Radio(L))(x[57])
    assert np.array_equal(outs[(None, "c64")][57].view(np.uint32), ref.view(np.uint32))
    # the channel-chunk path (full-rate output: a FIR filter) stages its chunks the same way
    h = O.firdes_kaiser(33, 0.1, 60.0)
    ys = []
    for env in (None, "1"):
        if env:
            monkeypatch.setenv("LQB_NO_STAGING", env)
        else:
            monkeypatch.delenv("LQB_NO_STAGING", raising=False)
        f = L.FIRFilter(h, channels=C)
        ys.append(np.concatenate([f(np.ascontiguousarray(x[:, :70000])), f(np.ascontiguousarray(x[:, 70000:]))], axis=1))
    assert np.array_equal(ys[0].view(np.uint32), ys[1].view(np.uint32))


@pytest.mark.parametrize("lanes,C,n", [("2", 37, 4098), ("4", 9, 1234), ("8", 3, 530), ("2", 16, 16), ("8", 1, 7)])
def test_guard_bands_around_device_buffers(cuda, monkeypatch, lanes, C, n):
    """compute-sanitizer is closed on this pool, so out-of-bounds writes are hunted the old way: input, output and
    nothing else sit between canary bands in one device allocation; ragged shapes (samples not a multiple of the 16-sample
    tile, channels not a multiple of a warp's share, a last TMA box hanging over both edges) run through the receiver
    chain by device pointers; the bands must come back untouched and the audio must equal the host-pointer call's."""
    monkeypatch.setenv("LQB_LANES", lanes)
    guard = 4096                                              # bytes of canary on each side (16-byte aligned offsets)
    r = _Radio(L, channels=C); ch = L.Chain(*r.stages())
    r2 = _Radio(L, channels=C); ch2 = L.Chain(*r2.stages())
    x = np.stack([am_iq(n, seed=900 + c) for c in range(C)])
    n_out = ch.out_len(n)
    xin, yout = C * n * 8, max(16, C * n_out * 4)
    pad = lambda b: (b + 15) // 16 * 16
    total = guard + pad(xin) + guard + pad(yout) + guard
    buf = L.DeviceBuffer(total)
    host = np.full(total, 0xA5, np.uint8)
    host[guard:guard + xin] = x.view(np.uint8).ravel()
    buf.upload(host)
    px, py = buf.ptr.value + guard, buf.ptr.value + guard + pad(xin) + guard
    got = ch.execute_dev(px, n, py, n_out, 0); L.synchronize()
    back = buf.download((total,), np.uint8)
    assert got == n_out
    for lo, hi in ((0, guard), (guard + xin, guard + pad(xin) + guard), (guard + pad(xin) + guard + C * n_out * 4, total)):
        assert np.all(back[lo:hi] == 0xA5), "canary band [%d, %d) was written" % (lo, hi)
    assert np.array_equal(back[guard:guard + xin], host[guard:guard + xin])          # the input is read-only
    y = back[guard + pad(xin) + guard:guard + pad(xin) + guard + C * n_out * 4].view(np.float32).reshape(C, n_out)
    assert np.array_equal(y.view(np.uint32), ch2(x).view(np.uint32))
