"""GPU vs the committed golden fixtures (tests/golden/oracle_vectors.npz, made by make_golden.py from the
CPU oracle -- the reference itself has none).  Inputs are regenerated from the same seeds."""
import os

import numpy as np
import pytest

import liquiddsp as L
from util import rel_l2, am_iq

pytestmark = pytest.mark.gpu
V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))


def test_stages_against_golden(cuda):
    x = V["x"]
    g = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, sos=(V["iir_sos_B"], V["iir_sos_A"]))
    assert np.array_equal(g(x), V["iir"])
    assert rel_l2(L.FIRFilter(V["fir_taps"])(x), V["fir"]) <= 1e-5
    assert np.array_equal(L.ComplexResampler(0.024, Fc=0.024)(x), V["resamp"])
    n = L.NCO(); n.freq = 0.3; n.phase = 1.0
    assert np.array_equal(n.mix_down(x), V["nco_down"])
    a = L.AGC(); a.scale = 0.01
    assert rel_l2(a(x), V["agc"]) <= 1e-5
    assert rel_l2(L.AmpModem(0.5, "dsb", True)(V["am_in"]), V["am"]) <= 1e-5
    assert rel_l2(L.FreqDem(0.1)(V["fm_in"]), V["fm"]) <= 1e-5
    assert np.array_equal(L.DeemphasisFilter(48000)(V["de_in"]), V["de"])


def test_amradio_chain_against_golden(cuda):
    iq = am_iq(2 * 65536, seed=0xB200)
    st = (L.ComplexIIRFilter("cheby2", order=8, Fc=15000 / 2e6, sos=(V["iir_sos_B"], V["iir_sos_A"])),
          L.ComplexResampler(0.024, Fc=0.024), L.AGC(), L.AmpModem(0.5, "dsb", True), L.DeemphasisFilter(48000))
    st[2].scale = 0.01
    chain = L.Chain(*st)
    pcm = np.concatenate([chain(iq[:65536]), chain(iq[65536:])])
    assert rel_l2(pcm, V["amradio_pcm"]) <= 1e-4
