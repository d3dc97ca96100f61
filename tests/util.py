"""Shared helpers for the parity tests."""
import numpy as np


def rel_l2(a, b):
    a = np.asarray(a).astype(np.complex128).ravel(); b = np.asarray(b).astype(np.complex128).ravel()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))


def crandn(rng, *shape, scale=1.0):
    return (scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))).astype(np.complex64)


def am_iq(n, fs=2e6, f_off=200.0, phase=0.3, seed=0xB200, n0=0, noise=0.02, amp=0.1):
    """SURVEY 8d AM signal: carrier offset + two audio tones + out-of-band interferer + noise."""
    rng = np.random.default_rng(seed)
    t = (n0 + np.arange(n)) / fs
    a = 0.6 * np.sin(2 * np.pi * 1000 * t) + 0.4 * np.sin(2 * np.pi * 2500 * t)
    x = amp * (1 + 0.5 * a) * np.exp(1j * (2 * np.pi * f_off * t + phase))
    x = x + 0.05 * np.exp(2j * np.pi * 60e3 * t)
    x = x + (noise / np.sqrt(2)) * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


def fm_iq(n, fs=2e6, kf=0.1, amp=1.0, seed=7, noise=0.01):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    m = np.sin(2 * np.pi * 1000 * t)
    ph = 2 * np.pi * kf * np.cumsum(m)
    x = amp * np.exp(1j * ph) + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


def split_points(n, k, rng):
    """k-1 random cut points -> list of (start, stop) covering [0, n)."""
    cuts = sorted(set(int(c) for c in rng.integers(1, n, size=k - 1)))
    edges = [0] + cuts + [n]
    return [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]


def fm_stereo_iq(n, fs=600e3, fl=1000.0, fr=1700.0, dev=75e3, seed=3, noise=0.005, phase=0.0):
    """Broadcast FM stereo multiplex: L+R, 19 kHz pilot, L-R on the 38 kHz suppressed carrier."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    left, right = np.sin(2 * np.pi * fl * t), np.sin(2 * np.pi * fr * t + 0.4)
    mpx = 0.45 * (left + right) + 0.1 * np.sin(2 * np.pi * 19e3 * t) + 0.45 * (left - right) * np.sin(2 * np.pi * 38e3 * t)
    x = np.exp(1j * (2 * np.pi * dev / fs * np.cumsum(mpx) + phase))
    x = x + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)
