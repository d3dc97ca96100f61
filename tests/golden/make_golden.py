"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference (python-liquiddsp) ships no tests, golden vectors or fixtures, and liquid-dsp -- where all
of its arithmetic lives -- is not installed here, so nothing can be generated from the reference itself.
What is committed instead:
  kat_scipy.npz      float64 zeros/poles/gains from scipy.signal for the IIR families the path uses, and
                     closed-form integer sequences (resampler step/phase/output counts, NCO phase words).
                     These pin the ORACLE.
  oracle_vectors.npz outputs of the CPU oracle for seeded inputs, one per stage plus the README chain.
                     These freeze the oracle (regression) and are what the GPU path is compared with in
                     tests/test_golden_gpu.py.
"""
import os
import sys

import numpy as np
import scipy.signal as ss

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from util import am_iq, fm_iq, crandn  # noqa: E402


def kat():
    out = {}
    for name, args in {
        "cheby2_8_lp": ("cheby2", (8, 60, 2 * 0.0075), dict(btype="lowpass")),
        "butter_5_hp": ("butter", (5, 2 * 0.1), dict(btype="highpass")),
        "cheby1_4_lp": ("cheby1", (4, 1.0, 2 * 0.1), dict(btype="lowpass")),
        "butter_2_lp": ("butter", (2, 2 * 0.2), dict(btype="lowpass")),
    }.items():
        z, p, k = getattr(ss, args[0])(*args[1], output="zpk", **args[2])
        out[name + "_z"], out[name + "_p"], out[name + "_k"] = np.asarray(z), np.asarray(p), np.float64(k)
    # resampler integers: float32 division 2^24 / rate, rounded
    rate = np.float32(48e3 / 2e6)
    step = int(np.round(np.float32(1 << 24) / rate))
    out["resamp_step"] = np.uint32(step)
    counts, phases, phase = [], [], 0
    for _ in range(16):
        n = 65536
        k = 0 if phase > n * (1 << 24) - 1 else (n * (1 << 24) - 1 - phase) // step + 1
        phase = phase + k * step - n * (1 << 24)
        counts.append(k); phases.append(phase)
    out["resamp_counts"], out["resamp_phases"] = np.array(counts, np.int64), np.array(phases, np.uint32)
    # NCO: theta_n = theta_0 + n * d_theta mod 2^32
    out["nco_dtheta_0p3"] = np.uint32(O.lib.orc_nco_constrain(0.3))
    return out


def vectors():
    rng = np.random.default_rng(2026)
    out = {}
    x = crandn(rng, 4096)
    out["x"] = x
    out["iir"] = O.ComplexIIRFilter("cheby2", order=8, Fc=0.0075)(x)
    out["iir_sos_B"], out["iir_sos_A"] = O.iirdes_sos("cheby2", "lowpass", 8, 0.0075)
    h = O.firdes_kaiser(64, 0.1, 60.0)
    out["fir_taps"], out["fir"] = h, O.FIRFilter(h)(x)
    out["resamp"] = O.ComplexResampler(0.024, Fc=0.024)(x)
    out["resamp_bank"] = O.ComplexResampler(0.024, Fc=0.024).bank()
    n = O.NCO(); n.freq = 0.3; n.phase = 1.0
    out["nco_down"] = n.mix_down(x)
    a = O.AGC(); a.scale = 0.01
    out["agc"] = a(x)
    t = np.arange(4096) / 48e3
    xa = (0.01 * (1 + 0.5 * np.sin(2 * np.pi * 1000 * t)) * np.exp(1j * (2 * np.pi * 5 * t + 0.4))).astype(np.complex64)
    out["am_in"], out["am"] = xa, O.AmpModem(0.5, "dsb", True)(xa)
    out["am_lp"], out["am_dc"] = O.AmpModem(0.5, "dsb", True).taps()
    xf = fm_iq(4096)
    out["fm_in"], out["fm"] = xf, O.FreqDem(0.1)(xf)
    xr = rng.standard_normal(4096).astype(np.float32)
    out["de_in"], out["de"] = xr, O.DeemphasisFilter(48000)(xr)
    out["sintab"] = O.nco_sintab()
    radio = O.AMRadio()
    iq = am_iq(2 * 65536, seed=0xB200)
    out["amradio_pcm"] = np.concatenate([radio(iq[:65536]), radio(iq[65536:])])
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "kat_scipy.npz"), **kat())
    np.savez_compressed(os.path.join(HERE, "oracle_vectors.npz"), **vectors())
    print("wrote", os.listdir(HERE))
