"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/liboracle.so (the plain-C restatement of the reference's
CPU arithmetic, oracle_c.c) and of oracle/_ref/libref_firdes.so (the reference's own
lib/window.cc + lib/firdes.cc compiled from /root/reference by oracle/Makefile).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under gr_clenabled_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")
REF = os.path.join(_HERE, "_ref", "libref_firdes.so")
REF_FILTERS = os.path.join(_HERE, "_ref", "libref_filters.so")

_lib = None
_ref = None
_ref_filters = None

_fp = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")
_up = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        L = C.CDLL(LIB)
        L.orc_num_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_rng_f32.argtypes = [_fp, C.c_long, C.c_uint64, C.c_uint64]
        L.orc_rng_i8.argtypes = [_bp, C.c_long, C.c_uint64, C.c_uint64]
        L.orc_mathconst_c32.argtypes = [_fp, _fp, C.c_long, C.c_float, C.c_int]
        L.orc_mathconst_f32.argtypes = [_fp, _fp, C.c_long, C.c_float, C.c_int]
        L.orc_mathconst_i32.argtypes = [_ip, _ip, C.c_long, C.c_float, C.c_int]
        L.orc_mathop_c32.argtypes = [_fp, _fp, _fp, C.c_long, C.c_int]
        L.orc_mathop_f32.argtypes = [_fp, _fp, _fp, C.c_long, C.c_int]
        L.orc_mathop_i32.argtypes = [_ip, _ip, _ip, C.c_long, C.c_int]
        L.orc_log10.argtypes = [_fp, _fp, C.c_long, C.c_float, C.c_float]
        L.orc_snr.argtypes = [_fp, _fp, _fp, C.c_long, C.c_float, C.c_float]
        L.orc_complex_to_mag.argtypes = [_fp, _fp, C.c_long]
        L.orc_complex_to_arg.argtypes = [_fp, _fp, C.c_long]
        L.orc_complex_to_magphase.argtypes = [_fp, _fp, _fp, C.c_long]
        L.orc_magphase_to_complex.argtypes = [_fp, _fp, _fp, C.c_long]
        L.orc_window_blackman.argtypes = [_fp, C.c_int]
        L.orc_window_hamming.argtypes = [_fp, C.c_int]
        L.orc_firdes_ntaps_hamming.argtypes = [C.c_double, C.c_double]
        L.orc_firdes_ntaps_hamming.restype = C.c_int
        L.orc_firdes_low_pass_hamming.argtypes = [_fp, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_fft_c32.argtypes = [_fp, _fp, C.c_int, C.c_long, C.c_int, C.c_void_p, C.c_int]
        L.orc_fft_c32.restype = C.c_int
        L.orc_fft_r32.argtypes = [_fp, _fp, C.c_int, C.c_long, C.c_void_p]
        L.orc_fft_r32.restype = C.c_int
        L.orc_fir_ccf.argtypes = [_fp, _fp, C.c_long, _fp, C.c_int, C.c_int]
        L.orc_fir_ccf.restype = C.c_long
        L.orc_fftfilt_sizes.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_fftfilt_create.argtypes = [_fp, C.c_int, C.c_int]
        L.orc_fftfilt_create.restype = C.c_void_p
        L.orc_fftfilt_destroy.argtypes = [C.c_void_p]
        L.orc_fftfilt_filter.argtypes = [C.c_void_p, C.c_long, _fp, _fp]
        L.orc_fftfilt_filter.restype = C.c_int
        L.orc_pfb.argtypes = [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, _ip, C.c_int, C.c_long]
        L.orc_pfb.restype = C.c_int
        L.orc_xengine_i8_exact.argtypes = [_bp, _ip, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_xengine_f32.argtypes = [C.c_void_p, C.c_void_p, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_unpack4.argtypes = [_up, _bp, C.c_long]
        L.orc_xc_mag.argtypes = [_fp, _fp, C.c_long]
        L.orc_xc_max_shift.argtypes = [C.c_int, C.c_int]
        L.orc_xc_max_shift.restype = C.c_int
        L.orc_xc_factors.argtypes = [_fp, _fp, C.c_int, C.c_int, _fp]
        L.orc_xc_find_max.argtypes = [_fp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
        L.orc_xcorr_fft_vcf.argtypes = [_fp, _fp, _fp, C.c_int, C.c_long, C.c_int]
        L.orc_xcorr_fft_vcf.restype = C.c_int
        L.orc_fir_ccc.argtypes = [_fp, _fp, C.c_long, _fp, C.c_int, C.c_int]
        L.orc_fir_ccc.restype = C.c_long
        L.orc_quad_demod.argtypes = [_fp, _fp, C.c_long, C.c_float]
        L.orc_sig_source.argtypes = [_fp, C.c_long, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        L.orc_sig_source_advance.argtypes = [C.c_double, C.c_double, C.c_long]
        L.orc_sig_source_advance.restype = C.c_double
        _lib = L
    return _lib


def ref():
    """The reference's own window/firdes code (None if it was never built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF):
            return None
        R = C.CDLL(REF)
        R.ref_window_build.argtypes = [C.c_int, C.c_int, C.c_double, _fp]
        R.ref_window_build.restype = C.c_int
        R.ref_firdes_low_pass.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, _fp, C.c_int]
        R.ref_firdes_low_pass.restype = C.c_int
        R.ref_firdes_high_pass.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, _fp, C.c_int]
        R.ref_firdes_high_pass.restype = C.c_int
        R.ref_firdes_band_pass.argtypes = [C.c_double] * 5 + [C.c_int, C.c_double, _fp, C.c_int]
        R.ref_firdes_band_pass.restype = C.c_int
        R.ref_firdes_root_raised_cosine.argtypes = [C.c_double] * 4 + [C.c_int, _fp, C.c_int]
        R.ref_firdes_root_raised_cosine.restype = C.c_int
        _ref = R
    return _ref


def ref_filters():
    """The reference's own fft_filter_ccf / fir_filter_ccf / fft_complex (lib/fft_filter.cc, fir_filter.cc,
    fft.cc compiled from /root/reference against oracle/shim/; None if never built)."""
    global _ref_filters
    if _ref_filters is None:
        if not os.path.exists(REF_FILTERS):
            return None
        R = C.CDLL(REF_FILTERS)
        R.ref_fftfilt_create.argtypes = [C.c_int, _fp, C.c_int]
        R.ref_fftfilt_create.restype = C.c_void_p
        R.ref_fftfilt_destroy.argtypes = [C.c_void_p]
        R.ref_fftfilt_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        R.ref_fftfilt_set_taps.argtypes = [C.c_void_p, _fp, C.c_int]
        R.ref_fftfilt_set_taps.restype = C.c_int
        R.ref_fftfilt_filter.argtypes = [C.c_void_p, C.c_int, _fp, _fp]
        R.ref_fftfilt_filter.restype = C.c_int
        R.ref_fftfilt_xformed_taps.argtypes = [C.c_void_p, _fp]
        R.ref_fir_ccf.argtypes = [_fp, C.c_int, _fp, C.c_long, C.c_int, _fp]
        R.ref_fir_ccf.restype = C.c_int
        R.ref_fft_complex.argtypes = [C.c_int, C.c_int, _fp, _fp, C.c_long]
        R.ref_fft_complex.restype = C.c_int
        _ref_filters = R
    return _ref_filters


class RefFftFilter:
    """The reference's fft_filter_ccf object itself (oracle/_ref/libref_filters.so)."""

    def __init__(self, taps, decim=1):
        self.R = ref_filters()
        assert self.R is not None, "oracle/_ref/libref_filters.so missing (make -C oracle where /root/reference exists)"
        t = np.ascontiguousarray(taps, np.float32)
        self.decim = decim
        self.h = self.R.ref_fftfilt_create(decim, t, t.size)
        assert self.h
        a, b = C.c_int(), C.c_int()
        self.R.ref_fftfilt_sizes(self.h, C.byref(a), C.byref(b))
        self.fftsize, self.nsamples = a.value, b.value

    def set_taps(self, taps):
        t = np.ascontiguousarray(taps, np.float32)
        self.nsamples = self.R.ref_fftfilt_set_taps(self.h, t, t.size)
        a, b = C.c_int(), C.c_int()
        self.R.ref_fftfilt_sizes(self.h, C.byref(a), C.byref(b))
        self.fftsize = a.value

    def xformed_taps(self):
        out = np.zeros(self.fftsize, np.complex64)
        self.R.ref_fftfilt_xformed_taps(self.h, _f(out))
        return out

    def filter(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.size % self.nsamples == 0 and x.size % self.decim == 0
        nout = x.size // self.decim
        out = np.zeros(nout + self.nsamples, np.complex64)
        self.R.ref_fftfilt_filter(self.h, nout, _f(x), _f(out))
        return out[:nout]

    def __del__(self):
        if getattr(self, "h", None):
            self.R.ref_fftfilt_destroy(self.h)
            self.h = None


def ref_fir(x_with_history, taps, decim=1):
    """The reference's fir_filter_ccf::filterN / filterNdec on K-1 history samples + new samples."""
    R = ref_filters()
    taps = np.ascontiguousarray(taps, np.float32)
    x = np.ascontiguousarray(x_with_history, np.complex64)
    nin = x.size - (taps.size - 1)
    n = (nin + decim - 1) // decim
    out = np.zeros(n, np.complex64)
    assert R.ref_fir_ccf(taps, taps.size, _f(x), n, decim, _f(out)) == 0
    return out


def ref_fft(x, n, direction=-1):
    """The reference's fft_complex plan wrapper (over the shim transform) on nvec vectors."""
    R = ref_filters()
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    assert R.ref_fft_complex(n, int(direction < 0), _f(x), _f(out), x.size // n) == 0
    return out


def _f(a):
    return np.ascontiguousarray(a).view(np.float32).reshape(-1)


# ---- inputs -------------------------------------------------------------------
SEED_M, SEED_F, SEED_L, SEED_P, SEED_X = 1001, 1002, 1003, 1004, 1005


def rng_c32(n, seed, first=0):
    out = np.empty(2 * n, np.float32)
    lib().orc_rng_f32(out, 2 * n, seed, 2 * first)
    return out.view(np.complex64)


def rng_f32(n, seed, first=0):
    out = np.empty(n, np.float32)
    lib().orc_rng_f32(out, n, seed, first)
    return out


def rng_i8(n, seed, first=0):
    out = np.empty(n, np.int8)
    lib().orc_rng_i8(out, n, seed, first)
    return out


def tone(n):
    """test_clenabled.cc:835-851 -- in[i] = (sin 2 pi i/N, cos 2 pi i/N)"""
    i = np.arange(n, dtype=np.float64)
    return (np.sin(2 * np.pi * i / n) + 1j * np.cos(2 * np.pi * i / n)).astype(np.complex64)


# ---- blocks ---------------------------------------------------------------------
def mathconst(x, k, op):
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    if x.dtype == np.complex64:
        lib().orc_mathconst_c32(_f(x), _f(out), x.size, k, op)
    elif x.dtype == np.float32:
        lib().orc_mathconst_f32(x, out, x.size, k, op)
    else:
        lib().orc_mathconst_i32(x, out, x.size, k, op)
    return out


def mathop(a, b, op):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = np.zeros_like(a)
    if a.dtype == np.complex64:
        lib().orc_mathop_c32(_f(a), _f(b), _f(out), a.size, op)
    elif a.dtype == np.float32:
        lib().orc_mathop_f32(a, b, out, a.size, op)
    else:
        lib().orc_mathop_i32(a, b, out, a.size, op)
    return out


def log10(a, n, k):
    out = np.zeros(a.size, np.float32)
    lib().orc_log10(np.ascontiguousarray(a, np.float32), out, a.size, n, k)
    return out


def snr(a, b, n, k):
    out = np.zeros(a.size, np.float32)
    lib().orc_snr(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32), out, a.size, n, k)
    return out


def complex_to_mag(x):
    out = np.zeros(x.size, np.float32)
    lib().orc_complex_to_mag(_f(x), out, x.size)
    return out


def complex_to_arg(x):
    out = np.zeros(x.size, np.float32)
    lib().orc_complex_to_arg(_f(x), out, x.size)
    return out


def magphase_to_complex(m, p):
    out = np.zeros(m.size, np.complex64)
    lib().orc_magphase_to_complex(np.ascontiguousarray(m, np.float32), np.ascontiguousarray(p, np.float32),
                                  _f(out), m.size)
    return out


def window_blackman(n):
    w = np.zeros(n, np.float32)
    lib().orc_window_blackman(w, n)
    return w


def window_hamming(n):
    w = np.zeros(n, np.float32)
    lib().orc_window_hamming(w, n)
    return w


def firdes_low_pass_hamming(gain, fs, fc, tw):
    n = lib().orc_firdes_ntaps_hamming(fs, tw)
    t = np.zeros(n, np.float32)
    lib().orc_firdes_low_pass_hamming(t, n, gain, fs, fc)
    return t


def fft(x, n, direction=-1, window=None, shift=False):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    w = None if window is None else np.ascontiguousarray(window, np.float32)
    rc = lib().orc_fft_c32(_f(x), _f(out), n, x.size // n, direction,
                           None if w is None else w.ctypes.data, int(shift))
    assert rc == 0
    return out


def fft_real(x, n, window=None):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size, np.complex64)
    w = None if window is None else np.ascontiguousarray(window, np.float32)
    rc = lib().orc_fft_r32(x, _f(out), n, x.size // n, None if w is None else w.ctypes.data)
    assert rc == 0
    return out


def fir(x_with_history, taps, decim=1):
    """x_with_history = K-1 history samples followed by the new samples."""
    taps = np.ascontiguousarray(taps, np.float32)
    x = np.ascontiguousarray(x_with_history, np.complex64)
    nin = x.size - (taps.size - 1)
    out = np.zeros((nin + decim - 1) // decim, np.complex64)
    n = lib().orc_fir_ccf(_f(x), _f(out), nin, taps, taps.size, decim)
    return out[:n]


def fftfilt_sizes(ntaps):
    a, b = C.c_int(), C.c_int()
    lib().orc_fftfilt_sizes(ntaps, C.byref(a), C.byref(b))
    return a.value, b.value


class FftFilter:
    """fft_filter_ccf (lib/fft_filter.cc): stateful overlap-add filter."""

    def __init__(self, taps, decim=1):
        self.taps = np.ascontiguousarray(taps, np.float32)
        self.decim = decim
        self.h = lib().orc_fftfilt_create(self.taps, self.taps.size, decim)
        self.fftsize, self.nsamples = fftfilt_sizes(self.taps.size)

    def filter(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.size % self.nsamples == 0 and x.size % self.decim == 0
        nout = x.size // self.decim
        out = np.zeros(nout + self.nsamples, np.complex64)
        lib().orc_fftfilt_filter(self.h, nout, _f(x), _f(out))
        return out[:nout]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fftfilt_destroy(self.h)
            self.h = None


def pfb(x, taps, M, R, ch_map, niter):
    taps = np.ascontiguousarray(taps, np.float32)
    m = np.ascontiguousarray(ch_map, np.int32)
    x = np.ascontiguousarray(x, np.complex64)
    assert x.size >= (niter - 1) * R + taps.size
    out = np.zeros(niter * m.size, np.complex64)
    rc = lib().orc_pfb(_f(x), _f(out), taps, taps.size, M, R, m, m.size, niter)
    assert rc == 0
    return out


def xengine_exact(buf_i8, A, F, T, npol):
    nbl = A * (A + 1) // 2
    out = np.zeros(F * nbl * npol * npol * 2, np.int32)
    lib().orc_xengine_i8_exact(np.ascontiguousarray(buf_i8, np.int8).reshape(-1), out, A, F, T, npol)
    return out.reshape(-1, 2)


def xengine_f32(buf, A, F, T, npol, accumulate_into=None):
    nbl = A * (A + 1) // 2
    out = np.zeros(F * nbl * npol * npol, np.complex64) if accumulate_into is None else accumulate_into
    buf = np.ascontiguousarray(buf)
    if buf.dtype == np.int8:
        lib().orc_xengine_f32(buf.ctypes.data, None, _f(out), A, F, T, npol, int(accumulate_into is not None))
    else:
        buf = buf.astype(np.complex64, copy=False)
        lib().orc_xengine_f32(None, buf.ctypes.data, _f(out), A, F, T, npol, int(accumulate_into is not None))
    return out


def unpack4(b):
    b = np.ascontiguousarray(b, np.uint8)
    out = np.zeros(2 * b.size, np.int8)
    lib().orc_unpack4(b, out, b.size)
    return out


# ---- SURVEY 8(f) "next" rows ---------------------------------------------------
XC_GROUP = 1024      # find_max work-group size on a device whose kernel work-group limit is 1024


def xc_max_shift(signal_length, max_search_index):
    return lib().orc_xc_max_shift(signal_length, max_search_index)


def xcorrelate(signals, signal_length, max_search_index=0, complex_in=False):
    """clXCorrelate (lib/clXCorrelate_impl.cc:1528-1594): signals[0] is the reference.
    Returns (corr float32[n-1], lag int32[n-1], factors float32[n-1][2*max_shift])."""
    L = signal_length
    ms = xc_max_shift(L, max_search_index)
    mags = []
    for x in signals:
        if complex_in:
            x = np.ascontiguousarray(x[:L], np.complex64)
            m = np.zeros(L, np.float32)
            lib().orc_xc_mag(_f(x), m, L)
        else:
            m = np.ascontiguousarray(x[:L], np.float32)
        mags.append(m)
    n = len(signals) - 1
    corr, lag = np.zeros(n, np.float32), np.zeros(n, np.int32)
    fac = np.zeros((n, 2 * ms), np.float32)
    for k in range(n):
        lib().orc_xc_factors(mags[0], mags[k + 1], L, ms, fac[k])
        c, i = C.c_float(), C.c_int()
        lib().orc_xc_find_max(fac[k], 2 * ms, min(2 * ms, XC_GROUP), C.byref(c), C.byref(i))
        corr[k], lag[k] = c.value, i.value - ms
    return corr, lag, fac


def xcorr_fft_vcf(ref_sig, sig, n, input_type=1):
    ref_sig = np.ascontiguousarray(ref_sig, np.complex64)
    sig = np.ascontiguousarray(sig, np.complex64)
    out = np.zeros(ref_sig.size, np.float32)
    rc = lib().orc_xcorr_fft_vcf(_f(ref_sig), _f(sig), out, n, ref_sig.size // n, input_type)
    assert rc == 0
    return out


def fir_ccc(x_with_history, taps, decim=1):
    x = np.ascontiguousarray(x_with_history, np.complex64)
    t = np.ascontiguousarray(taps, np.complex64)
    out = np.zeros(max(0, (x.size - (t.size - 1) + decim - 1) // decim), np.complex64)
    n = lib().orc_fir_ccc(_f(x), _f(out), x.size, _f(t), t.size, decim)
    return out[:n]


def quad_demod(x_with_history, gain):
    x = np.ascontiguousarray(x_with_history, np.complex64)
    out = np.zeros(x.size - 1, np.float32)
    lib().orc_quad_demod(_f(x), out, out.size, gain)
    return out


def sig_source(n, complex_out, waveform_sin, phase, phase_inc, ampl):
    out = np.zeros(n, np.complex64 if complex_out else np.float32)
    lib().orc_sig_source(_f(out), n, int(complex_out), int(waveform_sin), phase, phase_inc, ampl)
    return out


def sig_source_advance(phase, phase_inc, n):
    return lib().orc_sig_source_advance(phase, phase_inc, n)
