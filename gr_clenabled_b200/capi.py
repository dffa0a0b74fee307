"""ctypes binding of libclenabled_b200.so (include/clenabled_b200.h).

This is the same stub a maintainer of the reference would write for its python
layer (python/bindings/*_python.cc binds the C++ classes with pybind11; here the
boundary is the C ABI, so the binding is ctypes).  Nothing in this module
computes: every call goes to the CUDA library, and loading fails loudly when the
library has not been built -- there is no CPU fallback.
"""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# CLB200_LIB_PATH: an alternative build of the same library (kernel A/B measurements in tools/)
LIB_PATH = os.environ.get("CLB200_LIB_PATH") or os.path.join(_HERE, "lib", "libclenabled_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "clenabled_b200.h")

OK, EINVAL, ECUDA, ENOMEM, ESTATE = 0, -1, -2, -3, -4

# include/clenabled/GRCLBase.h:57-62
DTYPE_COMPLEX, DTYPE_FLOAT, DTYPE_INT, DTYPE_SHORT, DTYPE_BYTE, DTYPE_PACKEDXY = 1, 2, 3, 4, 5, 6
OP_MULTIPLY, OP_ADD, OP_SUBTRACT, OP_COMPLEX_CONJ, OP_MULTIPLY_CONJ = 1, 2, 3, 4, 5
OP_EMPTY, OP_EMPTY_W_COPY = 255, 254
FFT_FORWARD, FFT_BACKWARD = -1, 1
UNARY_LOG10, UNARY_COMPLEX_TO_MAG, UNARY_COMPLEX_TO_ARG = 1, 2, 3


class Clb200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("clenabled_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None

_vp, _i, _l, _f = C.c_void_p, C.c_int, C.c_long, C.c_float
_pf, _pi = C.POINTER(C.c_float), C.POINTER(C.c_int)
_ph = C.POINTER(C.c_void_p)
_pl = C.POINTER(C.c_long)
_pu64 = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); must list every CLB200_API symbol of the header
SIGNATURES = {
    "clb200_version": (C.c_char_p, []),
    "clb200_last_error": (C.c_char_p, []),
    "clb200_device_count": (_i, []),
    "clb200_device_name": (_i, [_i, C.c_char_p, _i]),
    "clb200_device_sm_count": (_i, [_i]),
    "clb200_select_device": (_i, [_i, _i, _i, _i]),
    "clb200_destroy": (_i, [_vp]),
    "clb200_get_counters": (_i, [_vp, _pu64, _pu64, _pu64]),
    "clb200_register_host_buffer": (_i, [_vp, C.c_size_t]),
    "clb200_unregister_host_buffer": (_i, [_vp]),
    "clb200_mathconst_create": (_i, [_i, _i, _f, _i, _ph]),
    "clb200_mathconst_set_k": (_i, [_vp, _f]),
    "clb200_mathconst_k": (_f, [_vp]),
    "clb200_mathconst_work": (_i, [_vp, _vp, _vp, _l]),
    "clb200_mathconst_launch_device": (_i, [_vp, _vp, _vp, _l, _vp]),
    "clb200_mathop_create": (_i, [_i, _i, _i, _ph]),
    "clb200_mathop_work": (_i, [_vp, _vp, _vp, _vp, _l]),
    "clb200_mathop_launch_device": (_i, [_vp, _vp, _vp, _vp, _l, _vp]),
    "clb200_unary_create": (_i, [_i, _i, _f, _f, _ph]),
    "clb200_unary_work": (_i, [_vp, _vp, _vp, _l]),
    "clb200_unary_launch_device": (_i, [_vp, _vp, _vp, _l, _vp]),
    "clb200_snr_create": (_i, [_i, _f, _f, _ph]),
    "clb200_snr_work": (_i, [_vp, _vp, _vp, _vp, _l]),
    "clb200_c2magphase_create": (_i, [_i, _ph]),
    "clb200_c2magphase_work": (_i, [_vp, _vp, _vp, _vp, _l]),
    "clb200_magphase2c_create": (_i, [_i, _ph]),
    "clb200_magphase2c_work": (_i, [_vp, _vp, _vp, _vp, _l]),
    "clb200_fft_create": (_i, [_i, _i, _vp, _i, _i, _i, _i, _ph]),
    "clb200_fft_work": (_i, [_vp, _vp, _vp, _l]),
    "clb200_fft_work_streams": (_i, [_vp, _ph, _ph, _i, _l]),
    "clb200_fft_launch_device": (_i, [_vp, _vp, _vp, _l, _vp]),
    "clb200_filter_create": (_i, [_i, _i, _vp, _i, _i, _ph]),
    "clb200_filter_set_taps": (_i, [_vp, _vp, _i]),
    "clb200_filter_ntaps": (_i, [_vp]),
    "clb200_filter_get_taps": (_i, [_vp, _vp, _i]),
    "clb200_filter_reset": (_i, [_vp]),
    "clb200_filter_ref_sizes": (_i, [_i, _pi, _pi]),
    "clb200_filter_work": (_i, [_vp, _vp, _l, _vp, _pl]),
    "clb200_filter_launch_device": (_i, [_vp, _vp, _l, _vp, _pl, _vp]),
    "clb200_pfb_create": (_i, [_i, _vp, _i, _i, _i, _i, _vp, _i, _ph]),
    "clb200_pfb_work": (_i, [_vp, _vp, _vp, _l]),
    "clb200_pfb_launch_device": (_i, [_vp, _vp, _vp, _l, _vp]),
    "clb200_probe_fp32": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "clb200_describe": (_i, [_vp, C.c_char_p, _i]),
    "clb200_set_debug": (_i, [_vp, _i]),
    "clb200_xengine_create": (_i, [_i, _i, _i, _i, _i, _i, _ph]),
    "clb200_xengine_input_bytes": (_l, [_vp]),
    "clb200_xengine_output_items": (_l, [_vp]),
    "clb200_xengine_work": (_i, [_vp, _vp, _vp, _i]),
    "clb200_xengine_work_i32": (_i, [_vp, _vp, _vp]),
    "clb200_xengine_launch_device": (_i, [_vp, _vp, _vp, _i, _vp]),
    "clb200_xengine_launch_device_i32": (_i, [_vp, _vp, _vp, _vp]),
    "clb200_xengine_launch_device_batch": (_i, [_vp, _vp, _vp, _i, _vp]),
    "clb200_xengine_stream_begin": (_i, [_vp, _i, _i]),
    "clb200_xengine_push_timesteps": (_i, [_vp, _ph, _i, _l]),
    "clb200_xengine_poll_result": (_i, [_vp, _vp, _i, _pi]),
    "clb200_xengine_stream_state": (_i, [_vp, _pl, _pl, _pl, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "clb200_xengine_stream_ports_stable": (_i, [_vp, _i]),
    "clb200_xengine_stream_end": (_i, [_vp]),
    "clb200_xengine_set_shard": (_i, [_vp, _i, _i]),
    "clb200_xengine_set_gather": (_i, [_vp, _i, _ph]),
    "clb200_xengine_launch_device_gather": (_i, [_vp, _vp, _vp]),
    "clb200_xengine_set_gather_sync": (_i, [_vp, _i, _ph, _vp, _vp]),
    "clb200_xengine_gather_wait": (_i, [_vp, _vp]),
    "clb200_mem_alloc": (_i, [_i, C.c_size_t, _ph]),
    "clb200_mem_free": (_i, [_i, _vp]),
    "clb200_mem_copy_to_host": (_i, [_i, _vp, _vp, C.c_size_t]),
    "clb200_ipc_export": (_i, [_i, _vp, _vp]),
    "clb200_ipc_open": (_i, [_i, _vp, _ph]),
    "clb200_ipc_close": (_i, [_i, _vp]),
    "clb200_xcorrelate_create": (_i, [_i, _i, _i, _i, _i, _ph]),
    "clb200_xcorrelate_max_shift": (_i, [_vp]),
    "clb200_xcorrelate_work": (_i, [_vp, _ph, _vp, _vp]),
    "clb200_xcorrelate_launch_device": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "clb200_xcorrelate_factors": (_i, [_vp, _i, _vp, _i]),
    "clb200_xcorr_fft_create": (_i, [_i, _i, _i, _i, _ph]),
    "clb200_xcorr_fft_work": (_i, [_vp, _ph, _ph, _l]),
    "clb200_xcorr_fft_launch_device": (_i, [_vp, _ph, _ph, _l, _vp]),
    "clb200_cfilter_create": (_i, [_i, _i, _vp, _i, _ph]),
    "clb200_cfilter_set_taps": (_i, [_vp, _vp, _i]),
    "clb200_cfilter_work": (_i, [_vp, _vp, _l, _vp, _pl]),
    "clb200_cfilter_launch_device": (_i, [_vp, _vp, _l, _vp, _pl, _vp]),
    "clb200_quaddemod_create": (_i, [_i, _f, _ph]),
    "clb200_quaddemod_work": (_i, [_vp, _vp, _vp, _l]),
    "clb200_quaddemod_launch_device": (_i, [_vp, _vp, _vp, _l, _vp]),
    "clb200_sigsource_create": (_i, [_i, _i, C.c_double, _i, C.c_double, C.c_double, _ph]),
    "clb200_sigsource_work": (_i, [_vp, _vp, _l]),
    "clb200_sigsource_launch_device": (_i, [_vp, _vp, _l, _vp]),
    "clb200_sigsource_phase": (C.c_double, [_vp]),
}


SIG_COS, SIG_SIN = 1, 2          # lib/clSignalSource_impl.h:27-28


def header_symbols(path=HEADER_PATH):
    """Every entry point declared CLB200_API in the C header."""
    text = open(path).read()
    return sorted(set(re.findall(r"CLB200_API\s+[\w\s\*]+?\b(clb200_\w+)\s*\(", text)))


def load():
    """Load the CUDA library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C gr_clenabled_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().clb200_last_error().decode(errors="replace")


def check(rc):
    if rc != OK:
        raise Clb200Error(rc, last_error())
    return rc


def device_count():
    n = load().clb200_device_count()
    if n < 0:
        raise Clb200Error(n, last_error())
    return n


def require_gpu():
    if device_count() < 1:
        raise Clb200Error(ECUDA, "no CUDA device present (this library has no CPU path)")
