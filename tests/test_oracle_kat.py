"""CPU: pins the oracle.  The reference has no tests or golden vectors and liquid-dsp is absent, so the
oracle (oracle/liquid_oracle.c) is checked against what CAN be known independently:
scipy.signal designs, closed-form integer identities, analytic steady states, and its own frozen outputs.
"""
import os

import numpy as np
import pytest
import scipy.signal as ss

from oracle import oracle as O
from util import rel_l2, crandn, am_iq

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KAT = np.load(os.path.join(G, "kat_scipy.npz"))
VEC = np.load(os.path.join(G, "oracle_vectors.npz"))


def _match_roots(a, b, tol):
    a = sorted(np.asarray(a, np.complex128), key=lambda c: (round(c.real, 6), c.imag))
    b = sorted(np.asarray(b, np.complex128), key=lambda c: (round(c.real, 6), c.imag))
    assert len(a) == len(b)
    assert max(abs(u - v) for u, v in zip(a, b)) < tol


# liquid designs in float; the Chebyshev ellipse axes suffer cancellation (t0 - 1/eps at eps = 1e-3), which
# moves the poles of the README filter by ~3e-5 from the double-precision design.  That error belongs to the
# algorithm being restated, so the KAT tolerance is 1e-4 on root positions, and the same formulas evaluated
# in float64 (below) must agree with scipy to 1e-9.
@pytest.mark.parametrize("key,ft,bt,order,fc,ap,As", [
    ("cheby2_8_lp", "cheby2", "lowpass", 8, 0.0075, 0.7, 60.0),
    ("butter_5_hp", "butter", "highpass", 5, 0.1, 0.7, 60.0),
    ("cheby1_4_lp", "cheby1", "lowpass", 4, 0.1, 1.0, 60.0),
    ("butter_2_lp", "butter", "lowpass", 2, 0.2, 0.7, 60.0)])
def test_iirdes_zpk_against_scipy(key, ft, bt, order, fc, ap, As):
    z, p, k = O.iirdes_dzpk(ft, bt, order, fc, 0.3, ap, As)
    _match_roots(p, KAT[key + "_p"], 1e-4)
    zs = KAT[key + "_z"]
    if len(zs) < len(z):                                  # scipy omits zeros at z = -1 (lowpass) / +1 (highpass)
        zs = np.concatenate([zs, np.full(len(z) - len(zs), -1.0 if bt == "lowpass" else 1.0)])
    _match_roots(z, zs, 1e-4)
    assert abs(k.real / KAT[key + "_k"] - 1) < 2e-3 and abs(k.imag) < 1e-6


@pytest.mark.parametrize("order", [1, 2, 3, 4, 5, 8, 11])
@pytest.mark.parametrize("bt", ["lowpass", "highpass"])
def test_ellip_design_against_scipy(order, bt):
    """liquid's elliptic design (Landen / Orfanidis) against scipy.signal.ellip computed live: the float design
    lands within 1e-4 of the double-precision poles and zeros; the gain follows k0 = 1 (odd) or the ripple floor (even)."""
    import scipy.signal as ss
    z, p, k = O.iirdes_dzpk("ellip", bt, order, 0.1, 0.3, 1.0, 40.0)
    zs, ps, ks = ss.ellip(order, 1.0, 40.0, 0.2, btype=bt, output="zpk")
    _match_roots(p, ps, 1e-4)
    if len(zs) < len(z):
        zs = np.concatenate([zs, np.full(len(z) - len(zs), -1.0 if bt == "lowpass" else 1.0)])
    _match_roots(z, zs, 5e-4)                             # zeros = j / (k cd(u)): float k moves them by up to 1.6e-4
    assert abs(k.real / ks - 1) < 2e-3 and abs(k.imag) < 1e-5


@pytest.mark.parametrize("order", [1, 2, 3, 5, 8, 12, 16])
def test_bessel_design_against_scipy(order):
    """Poles = scipy's delay-normalised Bessel prototype over sqrt((2n-1) ln 2) (liquid's 3 dB approximation), through
    liquid's bilinear map; the magnitude at the design cut-off is then near -3 dB."""
    import scipy.signal as ss
    _, pa, _ = ss.besselap(order, "delay")
    pa = pa / np.sqrt((2 * order - 1) * np.log(2.0))
    m = np.tan(np.pi * 0.1)
    z, p, k = O.iirdes_dzpk("bessel", "lowpass", order, 0.1, 0.3, 1.0, 40.0)
    _match_roots(p, (1 + pa * m) / (1 - pa * m), 1e-5)
    assert np.allclose(z, -1.0)
    f = O.ComplexIIRFilter("bessel", "lowpass", order, 0.1)
    assert abs(abs(f.freqresponse(0.0)) - 1) < 1e-3 and 0.66 < abs(f.freqresponse(0.1)) < 0.78


def test_cheby2_formulas_in_float64_match_scipy():
    n, As, fc = 8, 60.0, 0.0075
    es = 10 ** (-As / 20)
    t0 = np.sqrt(1 + 1 / es ** 2); tp = (t0 + 1 / es) ** (1 / n); tm = (t0 - 1 / es) ** (1 / n)
    b, a = 0.5 * (tp + tm), 0.5 * (tp - tm)
    L = n // 2
    th = (2 * (np.arange(L) + 1) + n - 1) * np.pi / (2 * n)
    pa = 1 / (a * np.cos(th) - 1j * b * np.sin(th)); pa = np.concatenate([pa, pa.conj()])
    thz = 0.5 * np.pi * (2 * (np.arange(L) + 1) - 1) / n
    za = -1 / (1j * np.cos(thz)); za = np.concatenate([za, -za])
    m = np.tan(np.pi * fc)
    pd, zd = (1 + m * pa) / (1 - m * pa), (1 + m * za) / (1 - m * za)
    kd = np.prod((1 - pd) / (1 - zd))                     # digital gain seeded with k0 = 1, NOT the analog gain
    _match_roots(pd, KAT["cheby2_8_lp_p"], 1e-8); _match_roots(zd, KAT["cheby2_8_lp_z"], 1e-8)
    assert abs(kd.real / KAT["cheby2_8_lp_k"] - 1) < 1e-5


def test_sos_pairing_order_and_dc_gain():
    B, A = O.iirdes_sos("cheby2", "lowpass", 8, 0.0075)
    assert B.shape == A.shape == (4, 3)
    assert np.all(np.diff(-A[:, 1]) > 0)                  # pole pairs by increasing real part
    assert np.all(np.diff(-B[:, 1] / B[:, 0]) > 0)        # zero pairs by increasing real part
    assert np.allclose(B[:, 0], B[0, 0])                  # gain spread evenly: k^(1/4) in every section
    f = O.ComplexIIRFilter("cheby2", order=8, Fc=0.0075)
    assert abs(abs(f.freqresponse(0.0)) - 1) < 1e-3       # k0 = 1: unit DC gain
    assert abs(f.freqresponse(0.03)) < 2e-3               # 60 dB stop band (60 kHz interferer at 2 MS/s)
    B3, A3 = O.iirdes_sos("butter", "lowpass", 3, 0.1)    # odd order: last section first-order
    assert B3.shape == (2, 3) and A3[1, 2] == 0 and B3[1, 2] == 0


def test_iir_execution_matches_scipy_sosfilt():
    B, A = O.iirdes_sos("cheby2", "lowpass", 8, 0.0075)
    rng = np.random.default_rng(0)
    x = crandn(rng, 50000)
    y = O.ComplexIIRFilter(_sos=(B, A))(x)
    ref = ss.sosfilt(np.hstack([B, A]).astype(np.float64), x.astype(np.complex128))
    assert rel_l2(y, ref) < 5e-4                          # fp32 DF-II noise of this filter
    assert rel_l2(O.iir_f64_truth(B, A, x), ref) < 1e-12


def test_fma_convention_distance_is_rounding_noise():
    """The two rounding conventions (fused / unfused) differ by the filter's own fp32 noise, ~5e-5."""
    B, A = O.iirdes_sos("cheby2", "lowpass", 8, 0.0075)
    x = am_iq(200000)
    y1 = O.ComplexIIRFilter(_sos=(B, A))(x)
    y2 = O.ComplexIIRFilter(_sos=(B, A), _lib=O.nofma_lib())(x)
    truth = O.iir_f64_truth(B, A, x)
    d12, d1, d2 = rel_l2(y1, y2), rel_l2(y1, truth), rel_l2(y2, truth)
    assert 1e-7 < d12 < 1e-3 and d1 < 1e-3 and d2 < 1e-3


def test_kaiser_design_against_scipy():
    # beta formula and window: liquid's (wlen-1)-normalised Kaiser equals scipy.signal.windows.kaiser
    for As in (60.0, 40.0, 30.0):
        beta = O.lib.orc_kaiser_beta_As(As)
        ref = 0.1102 * (As - 8.7) if As > 50 else 0.5842 * (As - 21) ** 0.4 + 0.07886 * (As - 21)
        assert abs(beta - ref) < 1e-5
        w = np.array([O.lib.orc_kaiser(i, 51, beta) for i in range(51)])
        assert np.max(np.abs(w - ss.windows.kaiser(51, beta))) < 2e-4
    assert O.lib.orc_kaiser_beta_As(20.0) == 0.0
    assert abs(O.lib.orc_besseli0f(2.5) - np.i0(2.5)) < 1e-4
    h = O.firdes_kaiser(64, 0.1, 60.0)
    t = np.arange(64) - 31.5
    ref = np.sinc(0.2 * t) * ss.windows.kaiser(64, 0.1102 * (60 - 8.7))
    assert np.max(np.abs(h - ref)) < 3e-4 and np.allclose(h, h[::-1], atol=1e-6)


def test_dc_blocker_taps():
    h = O.firdes_notch(25, 0.0, 20.0)                     # As <= 21 -> beta 0 -> rectangular: delta - 1/51
    ref = np.full(51, -1 / 51.0); ref[25] += 1
    assert np.max(np.abs(h - ref)) < 1e-6 and abs(h.sum()) < 1e-5


def test_resampler_integers_closed_form():
    r = O.ComplexResampler(48e3 / 2e6, Fc=48e3 / 2e6)
    assert r.step == int(KAT["resamp_step"]) == 0x29AAAAC0
    assert r.bank().shape == (16, 40)                     # npfb 13 -> 16 filters, 2*m = 40 taps each
    assert np.allclose(r.bank().sum(axis=1), 1.0, atol=1e-4)
    rng = np.random.default_rng(1)
    for k in range(16):
        y = r(crandn(rng, 65536))
        assert len(y) == int(KAT["resamp_counts"][k]) and r.phase == int(KAT["resamp_phases"][k])
    assert list(KAT["resamp_counts"][:8]) == [1573] * 7 + [1572]
    r2 = O.ComplexResampler(48e3 / 2e6, Fc=48e3 / 2e6)
    assert len(r2(np.zeros(20_000_000 // 40, np.complex64))) * 40 in range(479_900, 480_100)


def test_resampler_passes_dc_and_tone():
    r = O.ComplexResampler(0.024, Fc=0.024)
    y = r(np.ones(100000, np.complex64))
    assert abs(np.mean(y[50:]) - 1) < 1e-3
    n = np.arange(400000); f = 2000.0 / 2e6
    y = O.ComplexResampler(0.024, Fc=0.024)(np.exp(2j * np.pi * f * n).astype(np.complex64))
    k = np.arange(len(y)); fo = f / (2 ** 24 / r.step)
    ph = np.unwrap(np.angle(y[100:]))
    assert abs(np.polyfit(k[100:], ph, 1)[0] / (2 * np.pi) - fo) < 1e-6


def test_nco_closed_form_and_constrain_edges():
    o = O.NCO(); o.freq = 0.3
    d = o.dtheta_u32
    assert d == int(KAT["nco_dtheta_0p3"])
    assert abs(d / 2 ** 32 * 2 * np.pi - 0.3) < 1e-6
    x = np.ones(12345, np.complex64)
    y = o.mix_up(x)
    assert o.theta_u32 == (12345 * d) % (1 << 32)         # theta_n = n * d_theta mod 2^32
    tab = O.nco_sintab()
    idx = (((np.arange(12345, dtype=np.uint64) * d) % (1 << 32) + (1 << 21)) >> 22) & 0x3ff
    assert np.array_equal(y.imag, tab[idx]) and np.array_equal(y.real, tab[(idx + 256) & 0x3ff])
    # sinf of a float-rounded argument: near 2 pi the argument itself carries 2.4e-7 of rounding
    assert np.max(np.abs(tab - np.sin(2 * np.pi * np.arange(1024) / 1024))) < 3e-7
    c = O.lib.orc_nco_constrain
    assert c(0.0) == 0 and c(-1e-9) == 0                  # tiny negative: fpart rounds to 1.0f -> wraps to 0
    assert abs(c(np.pi) - 2 ** 31) <= 256 and abs(c(-np.pi / 2) - 3 * 2 ** 30) <= 256
    assert abs(c(2 * np.pi + 0.1) / 2 ** 32 * 2 * np.pi - 0.1) < 1e-5            # wraps modulo 2 pi


def test_agc_steady_state_and_squelch_machine():
    rng = np.random.default_rng(2)
    a = O.AGC(); a.scale = 0.01
    y = a(crandn(rng, 20000, scale=0.3))
    assert abs(np.mean(np.abs(y[10000:]) ** 2) / 0.01 ** 2 - 1) < 0.1       # converges to output power scale^2
    a = O.AGC(); a.gain = 2.0; a.lock = True
    x = crandn(rng, 100)
    assert np.allclose(a(x), 2.0 * x) and a.gain == 2.0                       # locked: no scale, gain frozen
    a = O.AGC(); a.bandwidth = 0.05; a.squelch = True; a.threshold = -30.0; a.set_timeout(100)
    env = np.concatenate([np.full(1500, 1e-3), np.full(2000, 1.0), np.full(2500, 1e-3)])
    x = (env * np.exp(1j * np.arange(6000))).astype(np.complex64)
    y = a(x)
    # gain starts at 1 (rssi 0 dB > -30 dB): RISE on the very first sample, SIGNALHI, then the gain climbs on the
    # weak input, rssi drops: FALL -> SIGNALLO (timeout 100) -> TIMEOUT -> ENABLED; the burst at 1500 gives RISE again
    assert len(a.rise_indices) == 2 and a.rise_indices[0] == 0 and 1500 <= a.rise_indices[1] < 1600
    assert np.all(y[:50] != 0) and np.all(y[400:1500] == 0) and np.all(y[1600:3500] != 0) and np.all(y[3800:] == 0)
    assert a.status == 1                                                      # ... TIMEOUT -> ENABLED


def test_ampmodem_recovers_tone_and_locks():
    n = 48000; t = np.arange(n) / 48e3
    m = np.sin(2 * np.pi * 1000 * t)
    x = (0.01 * (1 + 0.5 * m) * np.exp(1j * (2 * np.pi * 3.0 * t + 0.7))).astype(np.complex64)
    o = O.AmpModem(0.5, "dsb", True)
    y = o(x)
    seg = y[24000:]
    ref = 0.01 * m[24000 - 50:n - 50]                     # delay line 25 + dc-blocker group delay 25; gain A*mod/mod_index
    assert rel_l2(seg, ref) < 0.1
    th, dth = o.nco_u32()
    f = dth / 2 ** 32 * 48e3
    assert abs(f - 3.0) < 0.5                             # PLL pulled onto the 3 Hz carrier offset


def test_freqdem_and_deemphasis_analytic():
    n = np.arange(5000)
    y = O.FreqDem(0.1)(np.exp(2j * np.pi * 0.02 * n).astype(np.complex64))
    assert np.allclose(y[1:], 0.2, atol=1e-5)             # f / kf
    b0, a1 = O.DeemphasisFilter.coeffs(48000)
    xx = np.float32(np.exp(-1.0 / (75e-6 * 48000.0)))
    assert a1 == -xx and b0 == np.float32(1.0 - np.float64(xx)) and abs(xx - 0.757465) < 1e-6
    d = O.DeemphasisFilter(48000)
    assert abs(abs(d.freqresponse(0.0)) - 1) < 1e-6
    assert abs(d(np.ones(500, np.float32))[-1] - 1) < 1e-5


def test_bytes_to_iq():
    raw = np.array([32767, -32767, 0, 1], "<i2").tobytes()
    assert np.allclose(O.bytes_to_iq(raw), [1 - 1j, 0 + 1j / 32767])


def test_int16_conversion_without_division_is_exact():
    """The device converts int16 -> float with s*r followed by one Newton step (r = fl(1/32767)) instead of an IEEE
    division; for all 65536 inputs that is the correctly rounded (float)s / 32767.0f of utility.hpp:65-66."""
    s16 = np.arange(-32768, 32768, dtype=np.int32)
    ref = O.bytes_to_iq(np.stack([s16, s16[::-1]], axis=1).astype("<i2").tobytes())
    r = np.float32(1.0 / 32767.0); sf = s16.astype(np.float32)
    q0 = (sf * r).astype(np.float32)
    res = (sf.astype(np.float64) - q0.astype(np.float64) * 32767.0).astype(np.float32)     # fma(-q0, 32767, s), exact in double
    q = (q0.astype(np.float64) + res.astype(np.float64) * np.float64(r)).astype(np.float32)  # fma(res, r, q0)
    assert np.array_equal(q, ref.real) and np.array_equal(q[::-1], ref.imag)
    assert np.array_equal(ref.real, (sf / np.float32(32767.0)).astype(np.float32))


def test_oracle_frozen_vectors():
    """Regression-freeze of the oracle itself (same image on the GPU box -> same libm -> same bits)."""
    x = VEC["x"]
    assert np.array_equal(O.ComplexIIRFilter("cheby2", order=8, Fc=0.0075)(x), VEC["iir"])
    assert np.array_equal(O.FIRFilter(VEC["fir_taps"])(x), VEC["fir"])
    assert np.array_equal(O.ComplexResampler(0.024, Fc=0.024)(x), VEC["resamp"])
    n = O.NCO(); n.freq = 0.3; n.phase = 1.0
    assert np.array_equal(n.mix_down(x), VEC["nco_down"])
    a = O.AGC(); a.scale = 0.01
    assert np.array_equal(a(x), VEC["agc"])
    assert np.array_equal(O.AmpModem(0.5, "dsb", True)(VEC["am_in"]), VEC["am"])
    assert np.array_equal(O.FreqDem(0.1)(VEC["fm_in"]), VEC["fm"])
    assert np.array_equal(O.DeemphasisFilter(48000)(VEC["de_in"]), VEC["de"])
    radio = O.AMRadio(); iq = am_iq(2 * 65536, seed=0xB200)
    assert np.array_equal(np.concatenate([radio(iq[:65536]), radio(iq[65536:])]), VEC["amradio_pcm"])


def test_amradio_object_equals_stagewise_chain():
    iq = am_iq(3 * 65536)
    radio = O.AMRadio()
    bp = O.ComplexIIRFilter("cheby2", order=8, Fc=15000 / 2e6); rs = O.ComplexResampler(0.024, Fc=0.024)
    agc = O.AGC(); agc.scale = 0.01; am = O.AmpModem(0.5, "dsb", True); de = O.DeemphasisFilter(48000)
    for i in range(3):
        blk = iq[i * 65536:(i + 1) * 65536]
        assert np.array_equal(radio(blk), de(am(agc(rs(bp(blk))))))


def test_broadcast_am_recovers_the_modulation():
    """BroadcastAM (demod.hpp:94-153): carrier offset tracked by the arg() PLL, audio out with the DC removed."""
    fs, n = 48000.0, 60000
    t = np.arange(n) / fs
    audio = 0.5 * np.sin(2 * np.pi * 1000 * t)
    x = ((1 + audio) * np.exp(1j * (2 * np.pi * 35.0 * t + 0.7))).astype(np.complex64)
    d = O.BroadcastAM(25)
    y = d(x)
    th, dth = d.nco_u32()
    f_lock = dth / 2.0 ** 32 * fs
    assert abs(f_lock - 35.0) < 0.5                        # PLL frequency word sits on the carrier offset
    tail = y[-24000:]
    ref = audio[-24000 - 25:-25]                           # signal branch is delayed by m = 25
    assert abs(np.mean(tail)) < 0.02 and rel_l2(tail, ref) < 0.08      # the 20 Hz high-pass leads by ~3 degrees at 1 kHz


def test_firhilbf_sideband_selection_and_analytic_signal():
    """SSBDemod (demod.hpp:155-187): a tone above the carrier comes out of "usb" at twice its amplitude and is
    suppressed by > 70 dB in "lsb"; firhilbf_r2c turns a real tone into its analytic signal at half rate."""
    fs, n = 48000.0, 8000
    t = np.arange(n) / fs
    up = np.exp(2j * np.pi * 1500 * t).astype(np.complex64)
    rms = lambda v: float(np.sqrt(np.mean(np.asarray(v, np.float64) ** 2)))
    for band, other in (("usb", "lsb"), ("lsb", "usb")):
        tone = up if band == "usb" else up.conj()
        assert abs(rms(O.SSBDemod(band)(tone)[500:]) - np.sqrt(2)) < 1e-3
        assert rms(O.SSBDemod(other)(tone)[500:]) < 2 * 10 ** (-70 / 20)
    hq = O.SSBDemod("usb").hq()
    assert hq.size == 50 and np.allclose(hq, -hq[::-1], atol=1e-7)        # antisymmetric quadrature filter
    # HilbertTransform's complex64 branch keeps only the delayed imaginary part with alternating sign (utility.hpp:93)
    z = (np.arange(40) + 1j * (np.arange(40) + 100)).astype(np.complex64)
    y = O.HilbertTransform(5, 60.0)(z)
    k = np.arange(5, 40)
    assert np.all(y[:5] == 0) and np.array_equal(y[5:], ((-1.0) ** (k - 5) * (k - 5 + 100)).astype(np.float32))


def test_nco_constrain_single_precision_form_equals_the_double_form(tmp_path):
    """devmath.cuh nco_constrain_dev evaluates liquid's double-precision phase wrap in single precision;
    tools/check_nco_constrain.c proves equality for every finite float (4.28e9 inputs, 45 s on 8 cores) -- here every
    257th bit pattern (16.6 M inputs)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "chk")
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "check_nco_constrain.c")
    subprocess.check_call(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-o", exe, src, "-lm"])
    out = subprocess.run([exe, "257"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "mismatches 0" in out.stdout, out.stdout[-500:]
