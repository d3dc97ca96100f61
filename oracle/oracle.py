"""ctypes front-end to the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

PARITY UNPINNED: liquid-dsp (the dependency that holds all arithmetic of the reference path) is
un-vendored, un-pinned and absent here; `liquid_oracle.c` restates its published algorithms and is
pinned only by known-answer tests (tests/test_oracle_kat.py).

The classes below mirror the reference's Python surface (/root/reference/src/wrapper.cpp:10-273):
same class names, constructor arguments, defaults and `obj(ndarray) -> ndarray` behaviour, so the
parity tests can drive the oracle and the CUDA product through identical code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_cf = np.complex64
_f = np.float32


def build(force=False):
    """Compile liboracle.so in place (gcc, seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "liquid_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def _load(name="liboracle.so"):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build(force=True)
    lib = C.CDLL(path)
    P, F, U, I, U32 = C.c_void_p, C.c_float, C.c_uint, C.c_int, C.c_uint32
    sig = {
        "orc_iirdes_sos": (I, [I, I, U, F, F, F, F, P, P]),
        "orc_iirdes_dzpk": (I, [I, I, U, F, F, F, F, P, P, P]),
        "orc_kaiser_beta_As": (F, [F]), "orc_besseli0f": (F, [F]), "orc_lngammaf": (F, [F]),
        "orc_kaiser": (F, [U, U, F]), "orc_sincf": (F, [F]),
        "orc_firdes_kaiser": (I, [U, F, F, F, P]), "orc_firdes_notch": (I, [U, F, F, P]),
        "orc_iirfilt_crcf_create_sos": (P, [P, P, U]),
        "orc_iirfilt_crcf_create_prototype": (P, [I, I, U, F, F, F, F]),
        "orc_iirfilt_crcf_destroy": (None, [P]), "orc_iirfilt_crcf_reset": (None, [P]),
        "orc_iirfilt_crcf_get_sos": (U, [P, P, P]),
        "orc_iirfilt_crcf_execute_block": (None, [P, P, U, P]),
        "orc_iirfilt_crcf_freqresponse": (None, [P, F, P]),
        "orc_iirfilt_crcf_execute_block_f64": (None, [P, P, U, P, P, U, P]),
        "orc_iirfilt_rrrf_create": (P, [P, U, P, U]), "orc_iirfilt_rrrf_destroy": (None, [P]),
        "orc_iirfilt_rrrf_reset": (None, [P]), "orc_iirfilt_rrrf_freqresponse": (None, [P, F, P]),
        "orc_firfilt_create": (P, [P, U]), "orc_firfilt_create_kaiser": (P, [U, F, F, F]),
        "orc_firfilt_create_dc_blocker": (P, [U, F]), "orc_firfilt_destroy": (None, [P]),
        "orc_firfilt_reset": (None, [P]), "orc_firfilt_set_scale": (None, [P, F]),
        "orc_firfilt_get_taps": (U, [P, P]),
        "orc_firfilt_crcf_execute_block": (None, [P, P, U, P]),
        "orc_firfilt_rrrf_execute_block": (None, [P, P, U, P]),
        "orc_firfilt_freqresponse": (None, [P, F, P]),
        "orc_resamp_create": (P, [F, U, F, F, U]), "orc_resamp_destroy": (None, [P]),
        "orc_resamp_reset": (None, [P]), "orc_resamp_set_rate": (None, [P, F]),
        "orc_resamp_get_step": (U32, [P]), "orc_resamp_get_phase": (U32, [P]),
        "orc_resamp_get_npfb": (U, [P]), "orc_resamp_get_sublen": (U, [P]),
        "orc_resamp_get_bank": (None, [P, P]),
        "orc_resamp_execute_block": (U, [P, P, U, P]),
        "orc_nco_create": (P, [I]), "orc_nco_destroy": (None, [P]), "orc_nco_reset": (None, [P]),
        "orc_nco_constrain": (U32, [F]),
        "orc_nco_set_frequency": (None, [P, F]), "orc_nco_adjust_frequency": (None, [P, F]),
        "orc_nco_set_phase": (None, [P, F]), "orc_nco_adjust_phase": (None, [P, F]),
        "orc_nco_get_frequency": (F, [P]), "orc_nco_get_phase": (F, [P]),
        "orc_nco_get_theta_u32": (U32, [P]), "orc_nco_get_dtheta_u32": (U32, [P]),
        "orc_nco_set_u32": (None, [P, U32, U32]),
        "orc_nco_pll_set_bandwidth": (None, [P, F]), "orc_nco_pll_step": (None, [P, F]),
        "orc_nco_step": (None, [P]),
        "orc_nco_mix_block_up": (None, [P, P, P, U]), "orc_nco_mix_block_down": (None, [P, P, P, U]),
        "orc_nco_sintab": (C.POINTER(C.c_float), []),
        "orc_agc_create": (P, []), "orc_agc_destroy": (None, [P]), "orc_agc_reset": (None, [P]),
        "orc_agc_lock": (None, [P]), "orc_agc_unlock": (None, [P]),
        "orc_agc_set_bandwidth": (None, [P, F]), "orc_agc_get_bandwidth": (F, [P]),
        "orc_agc_get_signal_level": (F, [P]), "orc_agc_set_signal_level": (None, [P, F]),
        "orc_agc_get_rssi": (F, [P]), "orc_agc_set_rssi": (None, [P, F]),
        "orc_agc_get_gain": (F, [P]), "orc_agc_set_gain": (None, [P, F]),
        "orc_agc_get_scale": (F, [P]), "orc_agc_set_scale": (None, [P, F]),
        "orc_agc_squelch_enable": (None, [P]), "orc_agc_squelch_disable": (None, [P]),
        "orc_agc_squelch_set_threshold": (None, [P, F]), "orc_agc_squelch_get_threshold": (F, [P]),
        "orc_agc_squelch_set_timeout": (None, [P, U]), "orc_agc_squelch_get_status": (I, [P]),
        "orc_agc_get_y2_prime": (F, [P]),
        "orc_wrap_agc_execute": (U, [P, P, U, P, P, P, U]),
        "orc_ampmodem_create": (P, [F, I, I]), "orc_ampmodem_destroy": (None, [P]),
        "orc_ampmodem_reset": (None, [P]), "orc_ampmodem_demodulate_block": (None, [P, P, U, P]),
        "orc_ampmodem_get_lowpass_taps": (U, [P, P]), "orc_ampmodem_get_dcblock_taps": (U, [P, P]),
        "orc_ampmodem_get_nco": (None, [P, P, P]),
        "orc_freqdem_create": (P, [F]), "orc_freqdem_destroy": (None, [P]),
        "orc_freqdem_reset": (None, [P]), "orc_freqdem_demodulate_block": (None, [P, P, U, P]),
        "orc_wrap_deemph_coeffs": (None, [F, P, P]), "orc_wrap_deemph_create": (P, [F]),
        "orc_wrap_deemph_execute": (None, [P, P, U, P]),
        "orc_wrap_bytes_to_iq": (None, [P, U, P]),
        "orc_firhilbf_create": (P, [U, F]), "orc_firhilbf_destroy": (None, [P]), "orc_firhilbf_reset": (None, [P]),
        "orc_firhilbf_get_hq": (U, [P, P]), "orc_wrap_ssb_execute": (None, [P, I, P, U, P]),
        "orc_wrap_hilbert_c2r": (None, [P, P, U, P]), "orc_wrap_hilbert_r2c": (None, [P, P, U, P]),
        "orc_resamp_set_real_taps": (None, [P, I]), "orc_resamp_create_default": (P, [F]),
        "orc_wrap_fmstereo_create": (P, [F, F]), "orc_wrap_fmstereo_destroy": (None, [P]), "orc_wrap_fmstereo_reset": (None, [P]),
        "orc_wrap_fmstereo_get_state": (None, [P, P, P, P]), "orc_wrap_fmstereo_get_deemph": (None, [P, P, P]),
        "orc_wrap_fmstereo_execute": (U, [P, P, U, P]),
        "orc_wrap_bam_create": (P, [I]), "orc_wrap_bam_destroy": (None, [P]), "orc_wrap_bam_reset": (None, [P]),
        "orc_wrap_bam_get_nco": (None, [P, P, P]), "orc_wrap_bam_get_design": (U, [P, P, P, P]),
        "orc_wrap_bam_execute": (None, [P, P, U, P]), "orc_wrap_bam_set_dcblock": (None, [P, P, P, U]),
        "orc_amradio_create": (P, [F, F, F]), "orc_amradio_destroy": (None, [P]),
        "orc_amradio_execute": (U, [P, P, U, P]),
    }
    for name_, (res, args) in sig.items():
        fn = getattr(lib, name_)
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
_nofma = None


def nofma_lib():
    """The unfused-rounding build of the same source (convention-distance reporting only)."""
    global _nofma
    if _nofma is None:
        _nofma = _load("liboracle_nofma.so")
    return _nofma


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c64(x):
    return np.ascontiguousarray(x, dtype=_cf)   # pybind11 forcecast


def _f32(x):
    return np.ascontiguousarray(x, dtype=_f)


FILTER_TYPES = {"butter": 0, "cheby1": 1, "cheby2": 2, "ellip": 3, "bessel": 4}   # iirfilter.hpp:5-12
BAND_TYPES = {"lowpass": 0, "highpass": 1, "bandpass": 2, "bandstop": 3}          # iirfilter.hpp:14-20
AMPMODEM_TYPES = {"dsb": 0, "usb": 1, "lsb": 2}                                    # demod.hpp:221-226


# ---- design helpers -------------------------------------------------------------------------
def iirdes_sos(filter_type, band_type, order, fc, f0=0.3, ap=0.7, As=60.0, L=lib):
    B = np.zeros(3 * 16, _f); A = np.zeros(3 * 16, _f)
    n = L.orc_iirdes_sos(FILTER_TYPES[filter_type], BAND_TYPES[band_type], order, fc, f0, ap, As, _p(B), _p(A))
    if n < 0:
        raise ValueError("orc_iirdes_sos failed (%d)" % n)
    return B[:3 * n].reshape(n, 3).copy(), A[:3 * n].reshape(n, 3).copy()


def iirdes_dzpk(filter_type, band_type, order, fc, f0=0.3, ap=0.7, As=60.0):
    z = np.zeros(32, _cf); p = np.zeros(32, _cf); k = np.zeros(1, _cf)
    n = lib.orc_iirdes_dzpk(FILTER_TYPES[filter_type], BAND_TYPES[band_type], order, fc, f0, ap, As, _p(z), _p(p), _p(k))
    if n < 0:
        raise ValueError("orc_iirdes_dzpk failed (%d)" % n)
    return z[:n].copy(), p[:n].copy(), k[0]


def firdes_kaiser(n, fc, As, mu=0.0):
    h = np.zeros(n, _f)
    if lib.orc_firdes_kaiser(n, fc, As, mu, _p(h)) != 0:
        raise ValueError("bad firdes_kaiser arguments")
    return h


def firdes_notch(m, f0, As):
    h = np.zeros(2 * m + 1, _f)
    if lib.orc_firdes_notch(m, f0, As, _p(h)) != 0:
        raise ValueError("bad firdes_notch arguments")
    return h


def nco_sintab():
    return np.ctypeslib.as_array(lib.orc_nco_sintab(), shape=(1024,)).copy()


# ---- objects mirroring wrapper.cpp ------------------------------------------------------------
class ComplexIIRFilter:
    """wrapper.cpp:134-152, iirfilter.hpp:244-299."""

    def __init__(self, filter_type="butter", band_type="lowpass", order=2, Fc=0.2, F0=0.3, Ap=0.7, As=60.0,
                 _sos=None, _lib=None):
        self._L = _lib or lib
        self.filter_type = filter_type if filter_type in FILTER_TYPES else ""
        self.band_type = band_type if band_type in BAND_TYPES else ""
        self.order, self.Fc, self.F0, self.Ap, self.As = order, Fc, F0, Ap, As
        if _sos is not None:
            B, A = (np.ascontiguousarray(a, _f) for a in _sos)
            self._q = self._L.orc_iirfilt_crcf_create_sos(_p(B), _p(A), B.size // 3)
        else:
            self._q = self._L.orc_iirfilt_crcf_create_prototype(
                FILTER_TYPES.get(filter_type, 0), BAND_TYPES.get(band_type, 0), order, Fc, F0, Ap, As)
        if not self._q:
            raise ValueError("iirfilt create failed")

    def __del__(self):
        if getattr(self, "_q", None):
            self._L.orc_iirfilt_crcf_destroy(self._q); self._q = None

    def sos(self):
        B = np.zeros(48, _f); A = np.zeros(48, _f)
        n = self._L.orc_iirfilt_crcf_get_sos(self._q, _p(B), _p(A))
        return B[:3 * n].reshape(n, 3).copy(), A[:3 * n].reshape(n, 3).copy()

    def reset(self):
        self._L.orc_iirfilt_crcf_reset(self._q)

    def freqresponse(self, f):
        H = np.zeros(1, _cf); self._L.orc_iirfilt_crcf_freqresponse(self._q, f, _p(H)); return complex(H[0])

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _cf)
        self._L.orc_iirfilt_crcf_execute_block(self._q, _p(x), x.shape[0], _p(y))
        return y


class RealIIRFilter(ComplexIIRFilter):
    """wrapper.cpp:154-172, iirfilter.hpp:301-356.  iirfilt_rrrf runs the same section arithmetic on float samples;
    with real coefficients the real lane of the crcf restatement IS that arithmetic, so it is run on x + 0j."""

    def __call__(self, x):
        return np.ascontiguousarray(ComplexIIRFilter.__call__(self, _f32(x).astype(_cf)).real)


def _band_class(base, band, has_f0, name):
    # iirfilter.hpp:61-131 (C*) and :171-241 (R*): band fixed, f0 = 0.1f for low/highpass, Ap = 0.5, As = 20 defaults
    if has_f0:
        def __init__(self, filter_type="butter", order=2, Fc=0.2, F0=0.3, Ap=0.5, As=20.0):
            base.__init__(self, filter_type, band, order, Fc, F0, Ap, As)
    else:
        def __init__(self, filter_type="butter", order=2, Fc=0.2, Ap=0.5, As=20.0):
            base.__init__(self, filter_type, band, order, Fc, 0.1, Ap, As)
    return type(name, (base,), {"__init__": __init__})


CLowpassIIR = _band_class(ComplexIIRFilter, "lowpass", False, "CLowpassIIR")
CHighpassIIR = _band_class(ComplexIIRFilter, "highpass", False, "CHighpassIIR")
CBandpassIIR = _band_class(ComplexIIRFilter, "bandpass", True, "CBandpassIIR")
CBandstopIIR = _band_class(ComplexIIRFilter, "bandstop", True, "CBandstopIIR")
RLowpassIIR = _band_class(RealIIRFilter, "lowpass", False, "RLowpassIIR")
RHighpassIIR = _band_class(RealIIRFilter, "highpass", False, "RHighpassIIR")
RBandpassIIR = _band_class(RealIIRFilter, "bandpass", True, "RBandpassIIR")
RBandstopIIR = _band_class(RealIIRFilter, "bandstop", True, "RBandstopIIR")


class RIIRFilter:
    """wrapper.cpp:82-86, iirfilter.hpp:133-168: iirfilt_rrrf_create(b, nb, a, na), transfer-function form."""

    def __init__(self, Bc, Ac):
        b, a = _f32(Bc), _f32(Ac)
        self._q = lib.orc_iirfilt_rrrf_create(_p(b), b.size, _p(a), a.size)
        if not self._q:
            raise ValueError("iirfilt_rrrf_create failed")

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_iirfilt_rrrf_destroy(self._q); self._q = None

    def reset(self):
        lib.orc_iirfilt_rrrf_reset(self._q)

    def freqresponse(self, f):
        H = np.zeros(1, _cf); lib.orc_iirfilt_rrrf_freqresponse(self._q, f, _p(H)); return complex(H[0])

    def __call__(self, x):
        x = _f32(x); y = np.empty(x.shape[0], _f)
        lib.orc_wrap_deemph_execute(self._q, _p(x), x.shape[0], _p(y))      # the per-sample iirfilt_rrrf_execute loop
        return y


class CIIRFilter:
    """wrapper.cpp:30-34, iirfilter.hpp:23-58.  Real coefficients on complex samples: each lane is the rrrf recurrence."""

    def __init__(self, Bc, Ac):
        self._re, self._im = RIIRFilter(Bc, Ac), RIIRFilter(Bc, Ac)

    def reset(self):
        self._re.reset(); self._im.reset()

    def freqresponse(self, f):
        return self._re.freqresponse(f)

    def __call__(self, x):
        x = _c64(x)
        return (self._re(x.real) + 1j * self._im(x.imag)).astype(_cf)


def iir_f64_truth(B, A, x):
    """Same recurrence in double with the same float32 coefficients."""
    B = np.ascontiguousarray(B, _f).ravel(); A = np.ascontiguousarray(A, _f).ravel()
    x = _c64(x); st = np.zeros(4 * (B.size // 3)); y = np.zeros(2 * x.size)
    lib.orc_iirfilt_crcf_execute_block_f64(_p(B), _p(A), B.size // 3, _p(st), _p(x), x.size, _p(y))
    return y[0::2] + 1j * y[1::2]


class FIRFilter:
    """New crcf class modelled on RealFIRFilter (wrapper.cpp:244-247, firfilter.hpp:5-36)."""

    def __init__(self, h):
        h = _f32(h); self._q = lib.orc_firfilt_create(_p(h), h.size)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_firfilt_destroy(self._q); self._q = None

    def reset(self):
        lib.orc_firfilt_reset(self._q)

    def freqresponse(self, f):
        H = np.zeros(1, _cf); lib.orc_firfilt_freqresponse(self._q, f, _p(H)); return complex(H[0])

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _cf)
        lib.orc_firfilt_crcf_execute_block(self._q, _p(x), x.shape[0], _p(y))
        return y


class RealFIRFilter(FIRFilter):
    """wrapper.cpp:244-247."""

    def __call__(self, x):
        x = _f32(x); y = np.empty(x.shape[0], _f)
        lib.orc_firfilt_rrrf_execute_block(self._q, _p(x), x.shape[0], _p(y))
        return y


class RealDCBlocker(RealFIRFilter):
    """wrapper.cpp:249-252, firfilter.hpp:39-50."""

    def __init__(self, slen=25, As=20.0):
        self._q = lib.orc_firfilt_create_dc_blocker(slen, As)
        if not self._q:
            raise ValueError("firdes_notch failed")


class RealKaiserBessel(RealFIRFilter):
    """wrapper.cpp:254-257, firfilter.hpp:52-67: scale = 1.0 / abs(H(0)) (double division, narrowed to float)."""

    def __init__(self, flen=25, Fc=0.25, As=20.0, offset=0.0):
        self._q = lib.orc_firfilt_create_kaiser(flen, Fc, As, offset)
        if not self._q:
            raise ValueError("firdes_kaiser failed")
        H0 = self.freqresponse(0.0)
        mag = _f(np.hypot(_f(H0.real), _f(H0.imag)))
        lib.orc_firfilt_set_scale(self._q, float(_f(1.0 / float(mag))))

    def taps(self):
        h = np.zeros(4096, _f); n = lib.orc_firfilt_get_taps(self._q, _p(h)); return h[:n].copy()


class ComplexResampler:
    """wrapper.cpp:221-226, resampler.hpp:127-173."""

    def __init__(self, rate, len=20, Fc=None, As=60.0, nfilter=13):
        if Fc is None:
            raise TypeError("Fc is required")
        self._rate = rate
        self._q = lib.orc_resamp_create(rate, len, Fc, As, nfilter)
        if not self._q:
            raise ValueError("resamp create failed")

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_resamp_destroy(self._q); self._q = None

    def reset(self):
        lib.orc_resamp_reset(self._q)

    @property
    def rate(self):
        return self._rate

    @rate.setter
    def rate(self, r):
        self._rate = r; lib.orc_resamp_set_rate(self._q, r)

    step = property(lambda s: lib.orc_resamp_get_step(s._q))
    phase = property(lambda s: lib.orc_resamp_get_phase(s._q))

    def bank(self):
        n, m = lib.orc_resamp_get_npfb(self._q), lib.orc_resamp_get_sublen(self._q)
        b = np.zeros(n * m, _f); lib.orc_resamp_get_bank(self._q, _p(b)); return b.reshape(n, m)

    def __call__(self, x):
        x = _c64(x)
        cap = int(np.ceil(x.shape[0] * max(self._rate, 1e-3) * 1.01)) + 8
        y = np.empty(cap, _cf)
        nw = lib.orc_resamp_execute_block(self._q, _p(x), x.shape[0], _p(y))
        assert nw <= cap
        return y[:nw].copy()


class CResampler(ComplexResampler):
    """wrapper.cpp:20-23, resampler.hpp:38-70: resamp_crcf_create_default(rate) + execute_block."""

    def __init__(self, rate):
        self._rate = rate
        self._q = lib.orc_resamp_create_default(rate)
        if not self._q:
            raise ValueError("resamp create failed")


class RResampler(CResampler):
    """wrapper.cpp:15-18, resampler.hpp:4-36: the same on float samples (resamp_rrrf)."""

    def __call__(self, x):
        return np.ascontiguousarray(ComplexResampler.__call__(self, _f32(x).astype(_cf)).real)


class RealResampler(ComplexResampler):
    """wrapper.cpp:214-219, resampler.hpp:72-125: resamp_rrrf_create(rate, len, Fc, As, nfilter), per-sample loop."""

    def __init__(self, rate, len=20, Fc=None, As=60.0, nfilter=13):
        ComplexResampler.__init__(self, rate, len, Fc, As, nfilter)
        lib.orc_resamp_set_real_taps(self._q, 1)

    def __call__(self, x):
        return np.ascontiguousarray(ComplexResampler.__call__(self, _f32(x).astype(_cf)).real)


class Delay:
    """wrapper.cpp:25-28, utility.hpp:5-59: wdelay read-then-push, separate float and complex lines; the property
    setter rebuilds (and so clears) both."""

    def __init__(self, nd=1):
        self.delay = nd

    @property
    def delay(self):
        return self._nd

    @delay.setter
    def delay(self, nd):
        self._nd = int(nd)
        self._buf = {np.dtype(np.float32): np.zeros(self._nd + 1, _f), np.dtype(np.complex64): np.zeros(self._nd + 1, _cf)}
        self._ri = {k: 0 for k in self._buf}

    def __call__(self, x):
        x = np.asarray(x)
        if x.dtype not in self._buf:
            return None
        v, ri, y = self._buf[x.dtype], self._ri[x.dtype], np.empty_like(x)
        for n in range(x.shape[0]):                     # wdelay_read, then wdelay_push (wdelay.proto.c)
            y[n] = v[ri]; v[ri] = x[n]; ri = (ri + 1) % (self._nd + 1)
        self._ri[x.dtype] = ri
        return y


class NCO:
    """wrapper.cpp:201-212, nco.hpp:10-81."""

    def __init__(self, type="nco"):
        self.type = "nco" if type == "nco" else "vco"
        self._q = lib.orc_nco_create(0 if type == "nco" else 1)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_nco_destroy(self._q); self._q = None

    freq = property(lambda s: lib.orc_nco_get_frequency(s._q), lambda s, v: lib.orc_nco_set_frequency(s._q, v))
    phase = property(lambda s: lib.orc_nco_get_phase(s._q), lambda s, v: lib.orc_nco_set_phase(s._q, v))
    theta_u32 = property(lambda s: lib.orc_nco_get_theta_u32(s._q))
    dtheta_u32 = property(lambda s: lib.orc_nco_get_dtheta_u32(s._q))

    def adjust_frequency(self, df): lib.orc_nco_adjust_frequency(self._q, df)
    def adjust_phase(self, dp): lib.orc_nco_adjust_phase(self._q, dp)
    def set_pll_bandwidth(self, bw): lib.orc_nco_pll_set_bandwidth(self._q, bw)
    def pll_step(self, dphi): lib.orc_nco_pll_step(self._q, dphi)

    def mix_up(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _cf)
        lib.orc_nco_mix_block_up(self._q, _p(x), _p(y), x.shape[0]); return y

    def mix_down(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _cf)
        lib.orc_nco_mix_block_down(self._q, _p(x), _p(y), x.shape[0]); return y

    __call__ = mix_up


class AGC:
    """wrapper.cpp:228-242, agc.hpp:4-128.  state_last is per-object here (the reference's is a
    function-level static shared by all AGC objects, agc.hpp:110)."""

    def __init__(self):
        self._q = lib.orc_agc_create(); self._squelch = False; self._lock = False
        self.onRise = None; self._state_last = C.c_int(0); self.rise_indices = []

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_agc_destroy(self._q); self._q = None

    def _set_squelch(self, v):
        self._squelch = bool(v)
        (lib.orc_agc_squelch_enable if v else lib.orc_agc_squelch_disable)(self._q)

    def _set_lock(self, v):
        self._lock = bool(v)
        (lib.orc_agc_lock if v else lib.orc_agc_unlock)(self._q)

    squelch = property(lambda s: s._squelch, _set_squelch)
    lock = property(lambda s: s._lock, _set_lock)
    threshold = property(lambda s: lib.orc_agc_squelch_get_threshold(s._q), lambda s, v: lib.orc_agc_squelch_set_threshold(s._q, v))
    bandwidth = property(lambda s: lib.orc_agc_get_bandwidth(s._q), lambda s, v: lib.orc_agc_set_bandwidth(s._q, v))
    level = property(lambda s: lib.orc_agc_get_signal_level(s._q), lambda s, v: lib.orc_agc_set_signal_level(s._q, v))
    level_dB = property(lambda s: lib.orc_agc_get_rssi(s._q), lambda s, v: lib.orc_agc_set_rssi(s._q, v))
    gain = property(lambda s: lib.orc_agc_get_gain(s._q), lambda s, v: lib.orc_agc_set_gain(s._q, v))
    scale = property(lambda s: lib.orc_agc_get_scale(s._q), lambda s, v: lib.orc_agc_set_scale(s._q, v))
    status = property(lambda s: lib.orc_agc_squelch_get_status(s._q))
    y2_prime = property(lambda s: lib.orc_agc_get_y2_prime(s._q))

    def set_timeout(self, t): lib.orc_agc_squelch_set_timeout(self._q, t)
    def reset(self): lib.orc_agc_reset(self._q)

    def __call__(self, x):
        x = _c64(x); n = x.shape[0]; y = np.empty(n, _cf)
        idx = np.zeros(max(n, 1), np.uint32)
        r = lib.orc_wrap_agc_execute(self._q, _p(x), n, _p(y), C.byref(self._state_last), _p(idx), idx.size)
        self.rise_indices = idx[:r].tolist()
        if self.onRise is not None:
            for _ in range(r):
                self.onRise()
        return y


class AmpModem:
    """wrapper.cpp:189-199, demod.hpp:228-307."""

    def __init__(self, modulation=0.75, type="dsb", carrier=False):
        self._mod, self._car = modulation, carrier
        self._type = type if type in AMPMODEM_TYPES else ""
        self._q = None; self._make()

    def _make(self):
        if self._q:
            lib.orc_ampmodem_destroy(self._q)
        self._q = lib.orc_ampmodem_create(self._mod, AMPMODEM_TYPES.get(self._type, 0), 0 if self._car else 1)
        if not self._q:
            raise NotImplementedError("oracle restates the DSB modes only")

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_ampmodem_destroy(self._q); self._q = None

    def _set_mod(self, v): self._mod = v; self._make()
    def _set_car(self, v): self._car = bool(v); self._make()

    def _set_type(self, v):
        if v in AMPMODEM_TYPES:
            self._type = v; self._make()

    modulation = property(lambda s: s._mod, _set_mod)
    carrier = property(lambda s: s._car, _set_car)
    type = property(lambda s: s._type, _set_type)

    def reset(self): lib.orc_ampmodem_reset(self._q)

    def taps(self):
        lp = np.zeros(64, _f); dc = np.zeros(64, _f)
        n1 = lib.orc_ampmodem_get_lowpass_taps(self._q, _p(lp)); n2 = lib.orc_ampmodem_get_dcblock_taps(self._q, _p(dc))
        return lp[:n1].copy(), dc[:n2].copy()

    def nco_u32(self):
        t = C.c_uint32(); d = C.c_uint32()
        lib.orc_ampmodem_get_nco(self._q, C.byref(t), C.byref(d)); return t.value, d.value

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _f)
        lib.orc_ampmodem_demodulate_block(self._q, _p(x), x.shape[0], _p(y)); return y


class FreqDem:
    """wrapper.cpp:183-187, demod.hpp:189-219."""

    def __init__(self, kd):
        self._q = lib.orc_freqdem_create(kd)
        if not self._q:
            raise ValueError("freqdem create failed")

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_freqdem_destroy(self._q); self._q = None

    def reset(self): lib.orc_freqdem_reset(self._q)

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _f)
        lib.orc_freqdem_demodulate_block(self._q, _p(x), x.shape[0], _p(y)); return y


class DeemphasisFilter:
    """wrapper.cpp:178-181, iirfilter.hpp:358-392."""

    def __init__(self, sample_rate=48000):
        self._q = lib.orc_wrap_deemph_create(sample_rate)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_iirfilt_rrrf_destroy(self._q); self._q = None

    def reset(self): lib.orc_iirfilt_rrrf_reset(self._q)

    @staticmethod
    def coeffs(sample_rate):
        b0 = C.c_float(); a1 = C.c_float()
        lib.orc_wrap_deemph_coeffs(sample_rate, C.byref(b0), C.byref(a1)); return b0.value, a1.value

    def freqresponse(self, f):
        H = np.zeros(1, _cf); lib.orc_iirfilt_rrrf_freqresponse(self._q, f, _p(H)); return complex(H[0])

    def __call__(self, x):
        x = _f32(x); y = np.empty(x.shape[0], _f)
        lib.orc_wrap_deemph_execute(self._q, _p(x), x.shape[0], _p(y)); return y


class SSBDemod:
    """wrapper.cpp:269-272, demod.hpp:155-187: firhilbf(25, 60 dB), band == "usb" picks the upper side-band."""

    def __init__(self, band, _m=25, _As=60.0):
        self.usb = band == "usb"
        self._q = lib.orc_firhilbf_create(_m, _As)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_firhilbf_destroy(self._q); self._q = None

    def reset(self): lib.orc_firhilbf_reset(self._q)

    def hq(self):
        h = np.zeros(4096, _f); n = lib.orc_firhilbf_get_hq(self._q, _p(h)); return h[:n].copy()

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _f)
        lib.orc_wrap_ssb_execute(self._q, int(self.usb), _p(x), x.shape[0], _p(y))
        return y


class HilbertTransform:
    """wrapper.cpp:174-176, utility.hpp:71-108: complex64 in -> float32 out, float32 in -> complex64 out, else None."""

    def __init__(self, m=5, As=60.0):
        self._c2r, self._r2c = lib.orc_firhilbf_create(m, As), lib.orc_firhilbf_create(m, As)
        if not self._c2r:
            raise ValueError("firhilbf: m must be >= 2")

    def __del__(self):
        for q in (getattr(self, "_c2r", None), getattr(self, "_r2c", None)):
            if q:
                lib.orc_firhilbf_destroy(q)
        self._c2r = self._r2c = None

    def __call__(self, x):
        x = np.asarray(x)
        if x.dtype == np.complex64:
            x = np.ascontiguousarray(x); y = np.empty(x.shape[0], _f)
            lib.orc_wrap_hilbert_c2r(self._c2r, _p(x), x.shape[0], _p(y)); return y
        if x.dtype == np.float32:
            x = np.ascontiguousarray(x); y = np.empty(x.shape[0], _cf)
            lib.orc_wrap_hilbert_r2c(self._r2c, _p(x), x.shape[0], _p(y)); return y
        return None


class FMStereo:
    """wrapper.cpp:264-267, demod.hpp:4-85.  Output: interleaved [L0, R0, L1, R1, ...] float32."""

    def __init__(self, iq_rate=600000.0, pcm_rate=48000.0):
        self._q = lib.orc_wrap_fmstereo_create(iq_rate, pcm_rate)
        if not self._q:
            raise ValueError("FMStereo: bad rates")

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_wrap_fmstereo_destroy(self._q); self._q = None

    def reset(self): lib.orc_wrap_fmstereo_reset(self._q)

    def state(self):
        t = np.zeros(1, np.uint32); d = np.zeros(1, np.uint32); pe = np.zeros(1, _f)
        lib.orc_wrap_fmstereo_get_state(self._q, _p(t), _p(d), _p(pe)); return int(t[0]), int(d[0]), float(pe[0])

    def deemph(self):
        b = np.zeros(1, _f); a = np.zeros(1, _f); lib.orc_wrap_fmstereo_get_deemph(self._q, _p(b), _p(a)); return float(b[0]), float(a[0])

    def __call__(self, x):
        x = _c64(x); y = np.empty(2 * x.shape[0] + 8, _f)
        nw = lib.orc_wrap_fmstereo_execute(self._q, _p(x), x.shape[0], _p(y))
        return y[:nw].copy()


class BroadcastAM:
    """wrapper.cpp:259-262, demod.hpp:94-153."""

    def __init__(self, slen=25, _dcblock=None):
        self._q = lib.orc_wrap_bam_create(slen)
        if not self._q:
            raise ValueError("BroadcastAM: slen must be >= 1")
        if _dcblock is not None:
            B, A = (np.ascontiguousarray(a, _f).ravel() for a in _dcblock)
            lib.orc_wrap_bam_set_dcblock(self._q, _p(B), _p(A), B.size // 3)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_wrap_bam_destroy(self._q); self._q = None

    def reset(self): lib.orc_wrap_bam_reset(self._q)

    def design(self):
        lp = np.zeros(4096, _f); B = np.zeros(48, _f); A = np.zeros(48, _f)
        n = lib.orc_wrap_bam_get_design(self._q, _p(lp), _p(B), _p(A))
        return lp[:n].copy(), B[:6].reshape(2, 3).copy(), A[:6].reshape(2, 3).copy()

    def nco_u32(self):
        t = np.zeros(1, np.uint32); d = np.zeros(1, np.uint32)
        lib.orc_wrap_bam_get_nco(self._q, _p(t), _p(d)); return int(t[0]), int(d[0])

    def __call__(self, x):
        x = _c64(x); y = np.empty(x.shape[0], _f)
        lib.orc_wrap_bam_execute(self._q, _p(x), x.shape[0], _p(y))
        return y


def bytes_to_iq(b):
    """wrapper.cpp:13, utility.hpp:61-69."""
    a = np.frombuffer(b, dtype="<i2"); n = a.size // 2
    y = np.empty(n, _cf); a = np.ascontiguousarray(a[:2 * n])
    lib.orc_wrap_bytes_to_iq(_p(a), n, _p(y)); return y


class AMRadio:
    """README.md:41-58 as one native object (CPU baseline); one channel."""

    def __init__(self, bandwidth=15000, iq_rate=2000000, pcm_rate=48000):
        self._q = lib.orc_amradio_create(bandwidth, iq_rate, pcm_rate)

    def __del__(self):
        if getattr(self, "_q", None):
            lib.orc_amradio_destroy(self._q); self._q = None

    def __call__(self, iq):
        iq = _c64(iq); pcm = np.empty(iq.shape[0], _f)
        n = lib.orc_amradio_execute(self._q, _p(iq), iq.shape[0], _p(pcm)); return pcm[:n].copy()
