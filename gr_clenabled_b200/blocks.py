"""Python mirror of the reference's block API over the C ABI.

Class names, constructor argument order/names and setters follow the reference's
pybind11 module `clenabled` (python/bindings/*_python.cc, python/__init__.py) so
a flowgraph's `clenabled.clFFT(...)` call keeps its shape; `work()` takes and
returns numpy arrays the way a gr::block::work() call takes item buffers.
Every method forwards to libclenabled_b200.so -- no arithmetic happens here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, load

c64 = np.complex64


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _in(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


class _Block:
    """Owns one clb200 handle."""

    def __init__(self):
        self._h = C.c_void_p()
        self._lib = load()

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.clb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def counters(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(self._lib.clb200_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value, "launches": c.value}

    @staticmethod
    def _device(openCLPlatformType, devSelector, platformId, devId):
        d = load().clb200_select_device(openCLPlatformType, devSelector, platformId, devId)
        if d < 0:
            raise capi.Clb200Error(d, capi.last_error())
        return d


_NP = {capi.DTYPE_COMPLEX: np.complex64, capi.DTYPE_FLOAT: np.float32, capi.DTYPE_INT: np.int32}


class clMathConst(_Block):
    """clenabled.clMathConst (include/clenabled/clMathConst.h:51; python/bindings/clMathConst_python.cc)."""

    def __init__(self, idataType, openCLPlatformType, devSelector, platformId, devId, fValue,
                 operatorType, setDebug=0):
        super().__init__()
        self.dtype = idataType
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_mathconst_create(idataType, dev, float(fValue), operatorType,
                                                C.byref(self._h)))

    def k(self):
        return self._lib.clb200_mathconst_k(self._h)

    def set_k(self, k):
        check(self._lib.clb200_mathconst_set_k(self._h, float(k)))

    def work(self, x, out=None):
        x = _in(x, _NP[self.dtype])
        if out is None:
            out = np.zeros_like(x)
        check(self._lib.clb200_mathconst_work(self._h, _ptr(x), _ptr(out), x.size))
        return out

    def launch_device(self, d_in, d_out, nitems, stream=0):
        check(self._lib.clb200_mathconst_launch_device(self._h, d_in, d_out, nitems, stream))


class clMathOp(_Block):
    """clenabled.clMathOp (include/clenabled/clMathOp.h:42)."""

    def __init__(self, idataType, openCLPlatformType, devSelector, platformId, devId, operatorType,
                 setDebug=0):
        super().__init__()
        self.dtype = idataType
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_mathop_create(idataType, dev, operatorType, C.byref(self._h)))

    def work(self, a, b, out=None):
        a = _in(a, _NP[self.dtype])
        b = _in(b, _NP[self.dtype])
        if a.size != b.size:
            raise ValueError("clMathOp: inputs differ in length")
        if out is None:
            out = np.zeros_like(a)
        check(self._lib.clb200_mathop_work(self._h, _ptr(a), _ptr(b), _ptr(out), a.size))
        return out

    def launch_device(self, d_a, d_b, d_c, nitems, stream=0):
        check(self._lib.clb200_mathop_launch_device(self._h, d_a, d_b, d_c, nitems, stream))


class _Unary(_Block):
    def __init__(self, kind, in_dtype, openCLPlatformType, devSelector, platformId, devId,
                 nValue=0.0, kValue=0.0):
        super().__init__()
        self._in_dtype = in_dtype
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_unary_create(kind, dev, float(nValue), float(kValue), C.byref(self._h)))

    def work(self, x):
        x = _in(x, self._in_dtype)
        out = np.zeros(x.size, np.float32)
        check(self._lib.clb200_unary_work(self._h, _ptr(x), _ptr(out), x.size))
        return out


class clLog(_Unary):
    """clenabled.clLog(openCLPlatformType, devSelector, platformId, devId, nValue, kValue, setDebug)."""

    def __init__(self, openCLPlatformType, devSelector, platformId, devId, nValue, kValue, setDebug=0):
        super().__init__(capi.UNARY_LOG10, np.float32, openCLPlatformType, devSelector, platformId,
                         devId, nValue, kValue)


class clComplexToMag(_Unary):
    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug=0):
        super().__init__(capi.UNARY_COMPLEX_TO_MAG, c64, openCLPlatformType, devSelector, platformId, devId)


class clComplexToArg(_Unary):
    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug=0):
        super().__init__(capi.UNARY_COMPLEX_TO_ARG, c64, openCLPlatformType, devSelector, platformId, devId)


class clSNR(_Block):
    def __init__(self, openCLPlatformType, devSelector, platformId, devId, nValue, kValue, setDebug=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_snr_create(dev, float(nValue), float(kValue), C.byref(self._h)))

    def work(self, a, b):
        a = _in(a, np.float32)
        b = _in(b, np.float32)
        out = np.zeros(a.size, np.float32)
        check(self._lib.clb200_snr_work(self._h, _ptr(a), _ptr(b), _ptr(out), a.size))
        return out


class clComplexToMagPhase(_Block):
    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_c2magphase_create(dev, C.byref(self._h)))

    def work(self, x):
        x = _in(x, c64)
        mag = np.zeros(x.size, np.float32)
        ph = np.zeros(x.size, np.float32)
        check(self._lib.clb200_c2magphase_work(self._h, _ptr(x), _ptr(mag), _ptr(ph), x.size))
        return mag, ph


class clMagPhaseToComplex(_Block):
    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_magphase2c_create(dev, C.byref(self._h)))

    def work(self, mag, phase):
        mag = _in(mag, np.float32)
        phase = _in(phase, np.float32)
        out = np.zeros(mag.size, c64)
        check(self._lib.clb200_magphase2c_work(self._h, _ptr(mag), _ptr(phase), _ptr(out), mag.size))
        return out


class clFFT(_Block):
    """clenabled.clFFT -- argument order of lib/clFFT_impl.cc:35-36 / grc/clenabled_clFFT.block.yml:86-88."""

    def __init__(self, fftSize, clFFTDir, window, idataType, openCLPlatformType, devSelector,
                 platformId, devId, setDebug=0, num_streams=1, shift=False):
        super().__init__()
        self.fft_size = int(fftSize)
        self.dtype = idataType
        self.num_streams = int(num_streams)
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        w = np.ascontiguousarray(window, np.float32) if window is not None and len(window) else None
        check(self._lib.clb200_fft_create(self.fft_size, clFFTDir, _ptr(w) if w is not None else None,
                                          0 if w is None else w.size, idataType, dev, int(bool(shift)),
                                          C.byref(self._h)))

    def work(self, x, out=None):
        """x: one stream, nvec*fft_size items.  Returns complex64[nvec*fft_size]."""
        x = _in(x, np.float32 if self.dtype == capi.DTYPE_FLOAT else c64)
        if x.size % self.fft_size:
            raise ValueError("clFFT: input is not a whole number of vectors")
        nvec = x.size // self.fft_size
        if out is None:
            out = np.zeros(x.size, c64)
        check(self._lib.clb200_fft_work(self._h, _ptr(x), _ptr(out), nvec))
        return out

    def work_streams(self, xs):
        xs = [_in(x, np.float32 if self.dtype == capi.DTYPE_FLOAT else c64) for x in xs]
        nvec = xs[0].size // self.fft_size
        outs = [np.zeros(x.size, c64) for x in xs]
        n = len(xs)
        pin = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        pout = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        check(self._lib.clb200_fft_work_streams(self._h, pin, pout, n, nvec))
        return outs

    def work_ptr(self, in_ptr, out_ptr, nvec):
        """host pointers (ints), e.g. pinned torch tensors' data_ptr()"""
        check(self._lib.clb200_fft_work(self._h, in_ptr, out_ptr, nvec))

    def launch_device(self, d_in, d_out, nvec, stream=0):
        check(self._lib.clb200_fft_launch_device(self._h, d_in, d_out, nvec, stream))


class clFilter(_Block):
    """clenabled.clFilter (include/clenabled/clFilter.h:52-53): complex in, real taps."""

    def __init__(self, openCLPlatformType, devSelector, platformId, devId, decimation, taps,
                 nthreads=1, setDebug=0, use_time=False):
        super().__init__()
        self.decimation = int(decimation)
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        t = np.ascontiguousarray(taps, np.float32)
        check(self._lib.clb200_filter_create(dev, self.decimation, _ptr(t), t.size, int(bool(use_time)),
                                             C.byref(self._h)))

    def set_taps2(self, taps):
        t = np.ascontiguousarray(taps, np.float32)
        check(self._lib.clb200_filter_set_taps(self._h, _ptr(t), t.size))

    def taps(self):
        n = self._lib.clb200_filter_ntaps(self._h)
        t = np.zeros(n, np.float32)
        check(self._lib.clb200_filter_get_taps(self._h, _ptr(t), n))
        return t

    def set_nthreads(self, n):      # clFilter.h:58-61: accepted, unused (reference passes 1)
        pass

    def reset(self):
        check(self._lib.clb200_filter_reset(self._h))

    def work(self, x):
        """Feed the next x.size stream samples; returns the decimated outputs produced."""
        x = _in(x, c64)
        out = np.zeros(x.size // self.decimation + 2, c64)
        n_out = C.c_long(0)
        check(self._lib.clb200_filter_work(self._h, _ptr(x), x.size, _ptr(out), C.byref(n_out)))
        return out[:n_out.value]

    def launch_device(self, d_in, n_in, d_out, stream=0):
        n_out = C.c_long(0)
        check(self._lib.clb200_filter_launch_device(self._h, d_in, n_in, d_out, C.byref(n_out), stream))
        return n_out.value


def filter_ref_sizes(ntaps):
    a, b = C.c_int(), C.c_int()
    check(load().clb200_filter_ref_sizes(ntaps, C.byref(a), C.byref(b)))
    return a.value, b.value


class clPolyphaseChannelizer(_Block):
    """clenabled.clPolyphaseChannelizer (include/clenabled/clPolyphaseChannelizer.h:48-49)."""

    def __init__(self, openCLPlatformType, devSelector, platformId, devId, taps, buf_items,
                 num_channels, ninputs_per_iter, ch_map, setDebug=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        t = np.ascontiguousarray(taps, np.float32)
        m = np.ascontiguousarray(ch_map, np.int32)
        self.ntaps, self.M, self.R, self.nmap = t.size, int(num_channels), int(ninputs_per_iter), m.size
        self.buf_items = int(buf_items)
        check(self._lib.clb200_pfb_create(dev, _ptr(t), t.size, self.buf_items, self.M, self.R,
                                          _ptr(m), m.size, C.byref(self._h)))

    def input_items_for(self, niter):
        return (niter - 1) * self.R + self.ntaps

    def work(self, x, niter):
        """x: history layout (x[0] is ntaps-1 samples in the past), >= input_items_for(niter) items."""
        x = _in(x, c64)
        if x.size < self.input_items_for(niter):
            raise ValueError("clPolyphaseChannelizer: %d input items, %d needed" %
                             (x.size, self.input_items_for(niter)))
        out = np.zeros(niter * self.nmap, c64)
        check(self._lib.clb200_pfb_work(self._h, _ptr(x), _ptr(out), niter))
        return out

    def launch_device(self, d_in, d_out, niter, stream=0):
        check(self._lib.clb200_pfb_launch_device(self._h, d_in, d_out, niter, stream))


class clXEngine(_Block):
    """clenabled.clXEngine (include/clenabled/clXEngine.h:48-52, python/bindings/clXEngine_python.cc:40-62).

    The stream-port marshal, PDU/file outputs and ATA synchroniser of the reference
    live above the C ABI (gr_clenabled_b200/host); this class exposes one
    integration at a time: work(buffer[t][station][chan][pol]) -> visibilities.
    """

    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug, data_type,
                 polarization, num_inputs, output_format, first_channel, num_channels, integration,
                 antenna_list=(), output_file=False, file_base="", rollover_size_mb=0,
                 internal_synchronizer=False, sync_timestamp=0, object_name="",
                 starting_chan_center_freq=0.0, channel_width=0.0, disable_output=False,
                 pipeline_integration=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        self.data_type, self.npol = data_type, int(polarization)
        self.num_inputs, self.num_channels, self.integration = int(num_inputs), int(num_channels), int(integration)
        self.first_channel, self.output_format = first_channel, output_format
        check(self._lib.clb200_xengine_create(dev, data_type, self.npol, self.num_inputs,
                                              self.num_channels, self.integration, C.byref(self._h)))

    @property
    def num_baselines(self):
        return self.num_inputs * (self.num_inputs + 1) // 2

    def input_bytes(self):
        return self._lib.clb200_xengine_input_bytes(self._h)

    def output_items(self):
        return self._lib.clb200_xengine_output_items(self._h)

    def set_shard(self, total_channels, chan_first):
        check(self._lib.clb200_xengine_set_shard(self._h, total_channels, chan_first))

    def work(self, buf, accumulate=False, out=None):
        buf = np.ascontiguousarray(buf)
        if out is None:
            out = np.zeros(self.output_items(), c64)
        check(self._lib.clb200_xengine_work(self._h, _ptr(buf), _ptr(out), int(accumulate)))
        return out

    def work_i32(self, buf):
        buf = np.ascontiguousarray(buf)
        out = np.zeros((self.output_items(), 2), np.int32)
        check(self._lib.clb200_xengine_work_i32(self._h, _ptr(buf), _ptr(out)))
        return out

    # ---- streaming ingest: the general_work() shape of the reference block (work_processor / runThread) ----
    def stream_begin(self, pipeline_integration=0, result_slots=4):
        check(self._lib.clb200_xengine_stream_begin(self._h, int(pipeline_integration), int(result_slots)))

    def push(self, ports, ntime):
        """ports: one array (or raw address) per input stream holding `ntime` items each"""
        keep = [p if isinstance(p, int) else np.ascontiguousarray(p) for p in ports]
        a = (C.c_void_p * len(keep))(*[p if isinstance(p, int) else p.ctypes.data for p in keep])
        check(self._lib.clb200_xengine_push_timesteps(self._h, a, len(keep), int(ntime)))

    def poll(self, wait=False, out=None):
        """the oldest finished visibility matrix, or None if none is ready"""
        if out is None:
            out = np.zeros(self.output_items(), c64)
        ready = C.c_int(0)
        check(self._lib.clb200_xengine_poll_result(self._h, _ptr(out), int(wait), C.byref(ready)))
        return out if ready.value else None

    def stream_state(self):
        t, n, r = C.c_long(), C.c_long(), C.c_long()
        p, b = C.c_uint64(), C.c_uint64()
        check(self._lib.clb200_xengine_stream_state(self._h, C.byref(t), C.byref(n), C.byref(r), C.byref(p), C.byref(b)))
        return {"tracker": t.value, "integrations": n.value, "results_pending": r.value,
                "pushes": p.value, "pushes_blocked": b.value}

    def stream_ports_stable(self, stable=True):
        check(self._lib.clb200_xengine_stream_ports_stable(self._h, int(stable)))

    def stream_end(self):
        check(self._lib.clb200_xengine_stream_end(self._h))

    def launch_device(self, d_in, d_out, accumulate=False, stream=0):
        check(self._lib.clb200_xengine_launch_device(self._h, d_in, d_out, int(accumulate), stream))

    def launch_device_batch(self, d_in, d_out, nbatch, stream=0):
        check(self._lib.clb200_xengine_launch_device_batch(self._h, d_in, d_out, int(nbatch), stream))

    def launch_device_i32(self, d_in, d_out, stream=0):
        check(self._lib.clb200_xengine_launch_device_i32(self._h, d_in, d_out, stream))

    def set_gather(self, full_out_ptrs):
        a = (C.c_void_p * len(full_out_ptrs))(*full_out_ptrs)
        check(self._lib.clb200_xengine_set_gather(self._h, len(full_out_ptrs), a))

    def launch_device_gather(self, d_in, stream=0):
        check(self._lib.clb200_xengine_launch_device_gather(self._h, d_in, stream))

    def set_gather_sync(self, my_rank, flag_ptrs, multicast_out=None, multicast_flags=None):
        a = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        check(self._lib.clb200_xengine_set_gather_sync(self._h, int(my_rank), a, multicast_out, multicast_flags))

    def gather_wait(self, stream=0):
        check(self._lib.clb200_xengine_gather_wait(self._h, stream))


# ------------------------------------------------------------------------------------------
# SURVEY 8(f) "next" rows
# ------------------------------------------------------------------------------------------
def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


class clXCorrelate(_Block):
    """clenabled.clXCorrelate (include/clenabled/clXCorrelate.h:56-57).  work() takes the frame of
    every input and returns what the reference publishes on message port "corr": the dict
    {"corrvect": float32[n-1], "corrective_lags": int32[n-1]} (lib/clXCorrelate_impl.cc:1585-1593),
    or None for a frame dropped by decim_frames (:1540-1547)."""

    def __init__(self, openCLPlatformType, devSelector, platformId, devId, setDebug, num_inputs, signal_length,
                 data_type, data_size, max_search_index, decim_frames, async_=False):
        super().__init__()
        if data_size == 0:
            raise ValueError("Unknown data type.")                               # :710-714
        self.num_inputs, self.signal_length, self.dtype = int(num_inputs), int(signal_length), int(data_type)
        self.decim_frames, self._frame = int(decim_frames), 1
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_xcorrelate_create(dev, self.num_inputs, self.signal_length, self.dtype,
                                                 int(max_search_index), C.byref(self._h)))

    def max_shift(self):
        return self._lib.clb200_xcorrelate_max_shift(self._h)

    def work(self, inputs):
        if self.decim_frames > 1:
            keep = (self._frame % self.decim_frames) == 0
            self._frame += 1
            if keep:
                self._frame = 1
            else:
                return None
        arrs = [_in(x, _NP[self.dtype])[:self.signal_length] for x in inputs]
        assert len(arrs) == self.num_inputs and all(a.size == self.signal_length for a in arrs)
        arrs = [np.ascontiguousarray(a) for a in arrs]
        corr = np.zeros(self.num_inputs - 1, np.float32)
        lag = np.zeros(self.num_inputs - 1, np.int32)
        check(self._lib.clb200_xcorrelate_work(self._h, _ptr_array(arrs), _ptr(corr), _ptr(lag)))
        return {"corrvect": corr, "corrective_lags": lag}

    def factors(self, signal):
        out = np.zeros(2 * self.max_shift(), np.float32)
        check(self._lib.clb200_xcorrelate_factors(self._h, signal, _ptr(out), out.size))
        return out

    def launch_device(self, d_in, d_corr, d_lag, stream=0):
        check(self._lib.clb200_xcorrelate_launch_device(self._h, d_in, d_corr, d_lag, stream))


class clxcorrelate_fft_vcf(_Block):
    """clenabled.clxcorrelate_fft_vcf (include/clenabled/clxcorrelate_fft_vcf.h:49)."""

    def __init__(self, fftSize, num_inputs, openCLPlatformType, devSelector, platformId, devId, input_type=1):
        super().__init__()
        self.n, self.num_inputs = int(fftSize), int(num_inputs)
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_xcorr_fft_create(self.n, self.num_inputs, int(input_type), dev, C.byref(self._h)))

    def work(self, inputs):
        """inputs: num_inputs arrays of nvec*fftSize complex; returns num_inputs-1 float arrays."""
        arrs = [_in(x, c64) for x in inputs]
        nvec = arrs[0].size // self.n
        outs = [np.zeros(nvec * self.n, np.float32) for _ in range(self.num_inputs - 1)]
        check(self._lib.clb200_xcorr_fft_work(self._h, _ptr_array(arrs), _ptr_array(outs), nvec))
        return outs

    def launch_device(self, d_ins, d_outs, nvec, stream=0):
        a = (C.c_void_p * len(d_ins))(*d_ins)
        o = (C.c_void_p * len(d_outs))(*d_outs)
        check(self._lib.clb200_xcorr_fft_launch_device(self._h, a, o, nvec, stream))


class clComplexFilter(_Block):
    """clenabled.clComplexFilter (include/clenabled/clComplexFilter.h:706): complex taps, time domain."""

    def __init__(self, openclPlatform, devSelector, platformId, devId, decimation, taps, nthreads=1, setDebug=0):
        super().__init__()
        self.decimation = int(decimation)
        dev = self._device(openclPlatform, devSelector, platformId, devId)
        t = np.ascontiguousarray(taps, c64)
        check(self._lib.clb200_cfilter_create(dev, self.decimation, _ptr(t), t.size, C.byref(self._h)))

    def set_taps2(self, taps):
        t = np.ascontiguousarray(taps, c64)
        check(self._lib.clb200_cfilter_set_taps(self._h, _ptr(t), t.size))

    def work(self, x):
        x = _in(x, c64)
        out = np.zeros(x.size // self.decimation + 2, c64)
        n_out = C.c_long(0)
        check(self._lib.clb200_cfilter_work(self._h, _ptr(x), x.size, _ptr(out), C.byref(n_out)))
        return out[:n_out.value]

    def launch_device(self, d_in, n_in, d_out, stream=0):
        n_out = C.c_long(0)
        check(self._lib.clb200_cfilter_launch_device(self._h, d_in, n_in, d_out, C.byref(n_out), stream))
        return n_out.value


class clQuadratureDemod(_Block):
    """clenabled.clQuadratureDemod (include/clenabled/clQuadratureDemod.h:49)."""

    def __init__(self, gain, openCLPlatformType, devSelector, platformId, devId, setDebug=0):
        super().__init__()
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_quaddemod_create(dev, float(gain), C.byref(self._h)))

    def work(self, x):
        x = _in(x, c64)
        out = np.zeros(x.size, np.float32)
        check(self._lib.clb200_quaddemod_work(self._h, _ptr(x), _ptr(out), x.size))
        return out

    def launch_device(self, d_in, d_out, nitems, stream=0):
        check(self._lib.clb200_quaddemod_launch_device(self._h, d_in, d_out, nitems, stream))


class clSignalSource(_Block):
    """clenabled.clSignalSource (include/clenabled/clSignalSource.h:49-50)."""

    def __init__(self, idataType, openCLPlatformType, devSelector, platformId, devId, samp_rate, waveform, freq,
                 amplitude, setDebug=0):
        super().__init__()
        self.dtype = int(idataType)
        dev = self._device(openCLPlatformType, devSelector, platformId, devId)
        check(self._lib.clb200_sigsource_create(dev, self.dtype, float(samp_rate), int(waveform), float(freq),
                                                float(amplitude), C.byref(self._h)))

    def phase(self):
        return self._lib.clb200_sigsource_phase(self._h)

    def work(self, nitems):
        out = np.zeros(nitems, _NP[self.dtype])
        check(self._lib.clb200_sigsource_work(self._h, _ptr(out), nitems))
        return out

    def launch_device(self, d_out, nitems, stream=0):
        check(self._lib.clb200_sigsource_launch_device(self._h, d_out, nitems, stream))
