"""CPU tests: the C-ABI library loads and exports every symbol include/clenabled_b200.h
declares; argument checks that need no GPU behave like the reference's ctor checks."""
import ctypes as C
import os

import pytest

from gr_clenabled_b200 import capi


def test_library_is_built_in_tree():
    assert os.path.exists(capi.LIB_PATH), "run __graft_entry__.build()"


def test_every_header_symbol_is_exported_and_bound():
    names = capi.header_symbols()
    assert len(names) >= 40
    lib = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in the header but not exported: %s" % missing
    unbound = [n for n in names if n not in capi.SIGNATURES]
    assert not unbound, "no ctypes signature for: %s" % unbound
    extra = [n for n in capi.SIGNATURES if n not in names]
    assert not extra, "bound but not declared in the header: %s" % extra


def test_version_and_error_string():
    lib = capi.load()
    assert b"sm_100a" in lib.clb200_version()
    assert isinstance(capi.last_error(), str)


def test_no_gpu_means_loud_failure_not_fallback():
    lib = capi.load()
    if capi.device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.clb200_mathconst_create(capi.DTYPE_COMPLEX, 0, 2.0, capi.OP_MULTIPLY, C.byref(h))
    assert rc == capi.ECUDA and "no CUDA device" in capi.last_error()
    with pytest.raises(capi.Clb200Error):
        capi.require_gpu()


def test_argument_checks_before_device_use():
    lib = capi.load()
    h = C.c_void_p()
    # clFFT ctor: window length must be 0 or fft_size (lib/clFFT_impl.cc:74-76)
    w = (C.c_float * 4)()
    rc = lib.clb200_fft_create(8, -1, C.cast(w, C.c_void_p), 4, capi.DTYPE_COMPLEX, 0, 0, C.byref(h))
    assert rc == capi.EINVAL and "window not the same length" in capi.last_error()
    for bad_size in (1, (1 << 22) + 1, 1 << 23):      # any length 2 .. 2 Mi, powers of two to 4 Mi
        rc = lib.clb200_fft_create(bad_size, -1, None, 0, capi.DTYPE_COMPLEX, 0, 0, C.byref(h))
        assert rc == capi.EINVAL and "fft size" in capi.last_error()
    # clXEngine ctor: at least 2 inputs (lib/clXEngine_impl.cc:106-109)
    rc = lib.clb200_xengine_create(0, capi.DTYPE_BYTE, 1, 1, 16, 16, C.byref(h))
    assert rc == capi.EINVAL and "at least 2 inputs" in capi.last_error()
    # clPolyphaseChannelizer ctor (lib/clPolyphaseChannelizer_impl.cc:59-62)
    t = (C.c_float * 8)()
    m = (C.c_int * 2)(0, 1)
    rc = lib.clb200_pfb_create(0, C.cast(t, C.c_void_p), 8, 10, 4, 4, C.cast(m, C.c_void_p), 2, C.byref(h))
    assert rc == capi.EINVAL and "multiple of num_channels" in capi.last_error()
    a, b = C.c_int(), C.c_int()
    assert lib.clb200_filter_ref_sizes(256, C.byref(a), C.byref(b)) == 0
    assert (a.value, b.value) == (512, 257)           # fft_filter.cc:77-78
    assert lib.clb200_destroy(None) == 0
