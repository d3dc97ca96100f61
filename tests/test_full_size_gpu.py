"""GPU, at BASELINE.json's full sizes: the oracle cannot run 4e9 samples, so these check size-independent
properties -- identical channels give identical outputs wherever they sit in the grid, one long call equals
several short ones bit for bit, integer bookkeeping follows its closed form, FIR linearity -- plus the oracle
on a handful of channels copied back from the device-generated input."""
import numpy as np
import pytest

import liquiddsp as L
from oracle import oracle as O
from util import rel_l2

pytestmark = pytest.mark.gpu


def _radio(C):
    iir = L.ComplexIIRFilter(filter_type="cheby2", order=8, Fc=15000 / 2e6, channels=C)
    rs = L.ComplexResampler(rate=48e3 / 2e6, Fc=48e3 / 2e6, channels=C)
    agc = L.AGC(channels=C); agc.lock = False; agc.scale = 0.01
    return iir, rs, agc, L.AmpModem(0.5, "dsb", True, channels=C), L.DeemphasisFilter(48000, channels=C)


def _rows(buf, C, n, rows, dtype=np.complex64):
    """Copy selected rows of a device [C x n] buffer back (one small D2H per row)."""
    import ctypes
    out = {}
    es = np.dtype(dtype).itemsize
    for r in rows:
        a = np.empty(n, dtype)
        L._ck(L._lib.lqb_memcpy_d2h(a.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(buf.ptr.value + r * n * es), n * es, None))
        out[r] = a
    L.synchronize()
    return out


def test_config5_full_size(cuda):
    """65536 channels x 65536-sample blocks (34 GB per block), two blocks."""
    C, n, half = 65536, 65536, 32768
    xb = L.DeviceBuffer(C * n * 8)
    st_a = _radio(C)
    one = L.Chain(*st_a)
    n_cap = 1600
    ya = L.DeviceBuffer(C * n_cap * 4)
    probes = (0, 1, 31, 777, half - 1)
    orc = {c: (O.ComplexIIRFilter(_sos=st_a[0].sos()), O.ComplexResampler(48e3 / 2e6, Fc=48e3 / 2e6), O.AGC(), O.AmpModem(0.5, "dsb", True), O.DeemphasisFilter(48000)) for c in probes}
    for q in orc.values():
        q[2].scale = 0.01
    for blk in range(2):
        # the two halves of the channel axis carry the same 32768 signals
        L.synth_fill(0, xb.ptr.value, half, n, channel0=0, n0=blk * n)
        L.synth_fill(0, xb.ptr.value + half * n * 8, half, n, channel0=0, n0=blk * n)
        ka = one.execute_dev(xb.ptr.value, n, ya.ptr.value, n_cap)
        L.synchronize()
        y = ya.download((C, ka), np.float32)
        assert np.array_equal(y[:half].view(np.uint32), y[half:].view(np.uint32))          # position in the grid is irrelevant
        assert np.all(np.isfinite(y)) and y.std() > 1e-4
        xs = _rows(xb, C, n, probes)
        for c in probes:
            v = xs[c]
            for stage in orc[c]:
                v = stage(v)
            assert v.shape[0] == ka
            assert rel_l2(y[c], v) <= 1e-4, (blk, c)
    assert st_a[1].state()[1] == orc[0][1].phase


def test_config5_streaming_split_at_scale(cuda):
    """One 65536-sample call == calls of 16384 + 49152 samples, bit for bit, on 8192 channels."""
    C, n = 8192, 65536
    xb = L.DeviceBuffer(C * n * 8)
    L.synth_fill(0, xb.ptr.value, C, n)
    x = xb.download((C, n), np.complex64)
    a, b = L.Chain(*_radio(C)), L.Chain(*_radio(C))
    whole = a(x)
    parts = np.concatenate([b(np.ascontiguousarray(x[:, :16384])), b(np.ascontiguousarray(x[:, 16384:]))], axis=1)
    assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))


def test_config2_full_size_fir(cuda):
    """FIRFilter 64 taps on 1024 channels x 1M samples: oracle on two channels, linearity over the whole block."""
    C, n = 1024, 1 << 20
    h = O.firdes_kaiser(64, 0.1, 60.0)
    xb, yb = L.DeviceBuffer(C * n * 8), L.DeviceBuffer(C * n * 8)
    L.synth_fill(1, xb.ptr.value, C, n, seed=0xF12)
    f = L.FIRFilter(h, channels=C)
    assert f.execute_dev(xb.ptr.value, n, yb.ptr.value, n) == n
    L.synchronize()
    xs, ys = _rows(xb, C, n, (0, 1023)), _rows(yb, C, n, (0, 1023))
    for c in (0, 1023):
        assert rel_l2(ys[c], O.FIRFilter(h)(xs[c])) <= 1e-5
    # linearity: filtering 2x - x' equals 2 y - y' (x' = the same block shifted by one channel)
    x = xb.download((64, n), np.complex64); y = yb.download((64, n), np.complex64)
    mix = (2 * x[:32] - x[32:64]).astype(np.complex64)
    g = L.FIRFilter(h, channels=32)
    assert rel_l2(g(mix), 2 * y[:32] - y[32:64]) <= 1e-5


def test_config3_full_size_counts_and_phase(cuda):
    """NCO mix-down + resampler on 4096 channels, 32 blocks of 64K with state carry: output counts, resampler phase
    and oscillator phase follow their closed forms exactly; oracle on three channels for the first blocks."""
    C, n, nblk = 4096, 65536, 32
    nco = L.NCO(channels=C); rs = L.ComplexResampler(0.024, Fc=0.024, channels=C)
    f = (2 * np.pi * (0.05 + 0.4 * np.arange(C) / 4096)).astype(np.float32)
    nco.set_frequencies(f); nco.set_direction(True)
    chain = L.Chain(nco, rs)
    _, dth = nco.u32()
    xb, yb = L.DeviceBuffer(C * n * 8), L.DeviceBuffer(C * 1600 * 8)
    step, phase, total = 0x29AAAAC0, 0, 0
    probes = (0, 2048, 4095)
    orc = {c: (O.NCO(), O.ComplexResampler(0.024, Fc=0.024)) for c in probes}
    for c in probes:
        orc[c][0].freq = float(f[c])
    for blk in range(nblk):
        L.synth_fill(2, xb.ptr.value, C, n, n0=blk * n)
        k = chain.execute_dev(xb.ptr.value, n, yb.ptr.value, 1600)
        expect = 0 if phase > n * (1 << 24) - 1 else (n * (1 << 24) - 1 - phase) // step + 1
        phase = phase + expect * step - n * (1 << 24); total += n
        assert k == expect and rs.state() == (step, phase)
        if blk < 3:
            L.synchronize()
            y = yb.download((C, k), np.complex64); xs = _rows(xb, C, n, probes)
            for c in probes:
                assert rel_l2(y[c], orc[c][1](orc[c][0].mix_down(xs[c]))) <= 1e-5
    th, _ = nco.u32()
    assert np.array_equal(th, ((dth.astype(np.uint64) * total) % (1 << 32)).astype(np.uint32))


def test_config4_full_size(cuda):
    """IIR + AGC + FM demod on 16384 channels x 64K: oracle on three channels, identical channels identical."""
    C, n, half = 16384, 65536, 8192
    iir = L.ComplexIIRFilter("cheby2", order=8, Fc=0.0075, channels=C); agc = L.AGC(channels=C); fm = L.FreqDem(0.1, channels=C)
    chain = L.Chain(iir, agc, fm)
    xb, yb = L.DeviceBuffer(C * n * 8), L.DeviceBuffer(C * n * 4)
    L.synth_fill(3, xb.ptr.value, half, n); L.synth_fill(3, xb.ptr.value + half * n * 8, half, n)
    assert chain.execute_dev(xb.ptr.value, n, yb.ptr.value, n) == n
    L.synchronize()
    probes = (0, 1000, half - 1)
    xs, ys, ys2 = _rows(xb, C, n, probes), _rows(yb, C, n, probes, np.float32), _rows(yb, C, n, [p + half for p in probes], np.float32)
    for c in probes:
        yo = O.FreqDem(0.1)(O.AGC()(O.ComplexIIRFilter(_sos=iir.sos())(xs[c])))
        assert np.linalg.norm(ys[c] - yo) / np.sqrt(n) <= 5e-4
        assert np.array_equal(ys[c], ys2[c + half])
