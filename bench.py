#!/usr/bin/env python
"""bench.py -- the headline benchmark: clFFT forward 8192-pt gr_complex (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the clFFT hot path over one batch of synthetic vectors
(`--nvec` vectors of 8192 gr_complex per GPU; default 8192 vectors = 512 MiB in + 512 MiB
out, larger than the 126 MB L2, so no L2 flush is needed between steps).

  value     device-resident throughput (inputs in HBM), Msamples/s over all ranks, CUDA events
  e2e       same metric through the C ABI with pinned HOST buffers: clb200_fft_work() does
            H2D -> kernel -> D2H inside the timed region; `copy_ceiling` is the bare pinned
            H2D + D2H copy rate of the same bytes on the same ranks (no kernel), measured in the run
  roofline  the FFT kernel against MEASURED_PEAKS.json's HBM copy bandwidth, 16 B/sample
  cpu_baseline  the reference's CPU path on the host cores: the oracle port of clFFT_impl::testCPU and the
            same loop over pocketfft (scipy.fft, complex64) as a stand-in for FFTW3f; `value` is the faster
  blocks    the other hot-path blocks, same run: device-resident kernels with their rooflines, the X-engine
            through its host entry points, scheduler-sized calls (BASELINE configs[0])

Multi-GPU (N > 1, one process per GPU): clFFT vectors are sharded with no collective ("weak" scaling).  The
X-engine legs (BASELINE config 5) are under `blocks`: channel-sharded with the visibility gather fused into the
kernel (peer-memory / NVSwitch multicast stores + device-side completion flags) against NCCL all_gather, both
timed to consumable and checked against the unsharded result; host-fed sharded ingest through
clb200_xengine_push_timesteps; weak-scaled device-resident batches.  torch is used only for device memory,
streams/events and torch.distributed; every kernel is ours.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FFT_N = 8192
BYTES_PER_SAMPLE = 16          # 8 B read + 8 B written per gr_complex sample
METRIC = "Msamples/sec per block (clFFT forward 8192-pt)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nvec", type=int, default=8192, help="8192-pt vectors per GPU per step")
    ap.add_argument("--no-blocks", action="store_true", help="skip the secondary per-block numbers")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def workload_config(nvec, world):
    """the same dict in both arms (the driver compares them)"""
    return {"workload": "clFFT forward 8192-pt gr_complex, 1 stream per GPU (BASELINE configs[1])",
            "fft_size": FFT_N, "vectors_per_step_per_gpu": nvec, "gpus": world,
            "l2": "inputs larger than L2 (2 x %d MiB per step), no flush needed" % (nvec * FFT_N * 8 >> 20),
            "parallelism": "vectors sharded over %d GPU(s), no collective" % world}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(nvec):
    """dram bytes per launch of the FFT kernel from the newest committed `ncu --set full` capture
    (profiles/r*_traffic.json); it cannot be measured inside an unprofiled run, and is only valid for the
    launch shape that was captured."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f)["k_fft_8192pt_x8192vec"]
            return (t["bytes"] if nvec == 8192 else None), "profiles/" + name + " (ncu --set full capture of this launch shape)"
        except Exception:
            continue
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The default run's timed region is a
    few milliseconds, far below nvidia-smi's own start-up time, so the samples come from NVML directly
    (nvidia_ml_py, ~1 ms period, a thread in this process); the sampler starts before the last warm-up
    steps so that it also sees the clocks under load just before the region.  `sm_mhz` is the median of
    the samples taken inside the timed region when there are any (else of the under-load ones)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.thr, self.run = index, [], None, False
        self.t0 = self.t1 = None
        self.nv = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES renumbers CUDA ordinals, NVML does not
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None
            return
        self.run = True

        def rd():
            nv = self.nv
            while self.run:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((time.perf_counter(), sm, rs))
                except Exception:
                    pass
                time.sleep(0.001)
        self.thr = threading.Thread(target=rd, daemon=True)
        self.thr.start()
        t_end = time.perf_counter() + 2.0                 # NVML start-up can outlast the whole default run
        while not self.rows and time.perf_counter() < t_end:
            time.sleep(0.001)

    def mark(self, which):
        if which == 0:
            self.t0 = time.perf_counter()
        else:
            self.t1 = time.perf_counter()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self.run = False
        self.thr.join(timeout=1)
        try:
            mx = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        inside = [r for r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1]
        use = inside if inside else self.rows
        sm = sorted(r[1] for r in use)
        bits = 0
        for r in use:
            bits |= r[2]
        reasons = sorted(nm for b, nm in self.REASONS.items() if bits & b)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(self.rows),
                "samples_in_timed_region": len(inside), "reasons": reasons,
                "how": "NVML, 1 ms period, started 3 warm-up steps before the timed region"}


# --------------------------------------------------------------------------------
# reference arm / cpu baseline: clFFT_impl::testCPU (lib/clFFT_impl.cc:464-518) on the host cores.
# FFTW3f is not in the image, so the transform inside that loop is (a) the oracle's radix-2 restatement
# (OpenMP over vectors) and (b) pocketfft through scipy.fft on complex64 (workers = threads) -- the
# stronger, FFTW-class baseline; the faster of the two is the arm's value.
# --------------------------------------------------------------------------------
def _cpu_fft_step(impl, x2d, threads):
    if impl == "pocketfft":
        import scipy.fft as sf
        return sf.fft(x2d, axis=1, workers=threads)
    from oracle import oracle as orc
    return orc.fft(x2d.reshape(-1), FFT_N, -1)


def cpu_fft_rate(impl, nvec, threads, min_s, x=None):
    """Msamples/s of `impl` over nvec vectors per call, repeated for about min_s seconds"""
    import numpy as np
    from oracle import oracle as orc
    orc.lib().orc_set_threads(threads)
    if x is None:
        x = orc.rng_c32(FFT_N * nvec, orc.SEED_F).reshape(nvec, FFT_N)
    _cpu_fft_step(impl, x[:min(nvec, 64)], threads)              # warm (plan / thread pool)
    reps, t0 = 0, time.perf_counter()
    while True:
        _cpu_fft_step(impl, x, threads)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_s:
            return FFT_N * nvec * reps / dt / 1e6, dt, reps


def run_reference(args, rank):
    """--impl reference: the reference's CPU path on all host cores, same batch per step as our arm"""
    if rank != 0:
        return
    import numpy as np
    from oracle import oracle as orc
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:                        # noqa: BLE001
        cores = os.cpu_count() or 1
    nvec = args.nvec
    x = orc.rng_c32(FFT_N * nvec, orc.SEED_F).reshape(nvec, FFT_N)
    # pick the fastest (implementation, thread count) on a short trial, then time exactly `steps` steps of it
    trial = {}
    for impl in ("pocketfft", "oracle_radix2"):
        for th in sorted({cores, max(1, cores // 2)}, reverse=True):
            try:
                trial["%s_%dthreads" % (impl, th)] = cpu_fft_rate(impl, min(nvec, 1024), th, 0.5, x[:min(nvec, 1024)])[0]
            except Exception:                # noqa: BLE001
                trial["%s_%dthreads" % (impl, th)] = 0.0
    best = max(trial, key=trial.get)
    impl, cores = best.rsplit("_", 1)[0], int(best.rsplit("_", 1)[1].replace("threads", ""))
    orc.lib().orc_set_threads(cores)
    for _ in range(max(1, min(args.warmup, 3))):
        _cpu_fft_step(impl, x, cores)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        _cpu_fft_step(impl, x, cores)
    dt = time.perf_counter() - t0
    val = FFT_N * nvec * steps / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC,
        "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(nvec, args.gpus),
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": "port",
                         "implementation": impl, "trial_Msamples_s": trial,
                         "sample": "%d vectors of 8192 per step x %d steps; clFFT_impl::testCPU's loop with the "
                                   "transform done by %s (FFTW3f of the reference is not in the image), all host "
                                   "threads" % (nvec, steps, "pocketfft (scipy.fft, complex64)" if impl == "pocketfft"
                                                else "the oracle's radix-2 FFT (OpenMP over vectors)")},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------
def _timeit(torch, fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1e3


def fp32_peaks(capi, dev):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    capi.check(capi.load().clb200_probe_fp32(dev, C.byref(a), C.byref(b), C.byref(c)))
    return {"ffma_tflops": a.value, "fadd_tflops": b.value, "fadd2_tflops": c.value,
            "how": "clb200_probe_fp32: 16 independent chains per thread, 8 x 256-thread CTAs per SM, best of 3"}


def secondary_blocks(torch, blocks, capi, dev, stream_ptr, hbm_peak):
    """device-resident throughput of the other blocks, a few launches each (not the headline)"""
    import numpy as np
    from oracle import oracle as orc
    out = {}
    gpu = (1, 2, 0, dev)

    def hbm(nbytes, t, nsamp):
        return {"Msamples_s": nsamp / t / 1e6, "GBps": nbytes / t / 1e9, "frac_hbm": nbytes / t / 1e9 / hbm_peak}

    try:
        fp = fp32_peaks(capi, dev)
        out["fp32_peak_measured"] = fp
    except Exception as e:                           # noqa: BLE001
        fp = None
        out["fp32_peak_measured"] = {"error": str(e)}

    n = 1 << 26                                     # 64 Mi complex samples = 512 MiB
    a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    try:
        blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *gpu, 0.7071, capi.OP_MULTIPLY)
        t = _timeit(torch, lambda: blk.launch_device(a.data_ptr(), b.data_ptr(), n, stream_ptr))
        out["clMultiplyConst"] = hbm(16 * n, t, n)
        # 2 -> 1 (clMathOp multiply): 24 B per output sample; half-size streams keep the three buffers in 1 GiB
        h = n // 2
        blk = blocks.clMathOp(capi.DTYPE_COMPLEX, *gpu, capi.OP_MULTIPLY)
        t = _timeit(torch, lambda: blk.launch_device(a.data_ptr(), a.data_ptr() + 8 * h, b.data_ptr(), h, stream_ptr))
        out["clMultiply_2to1"] = hbm(24 * h, t, h)
        # row M5: complex -> magnitude (12 B per sample), 10*log10 (8 B per float)
        lib = capi.load()
        blk = blocks.clComplexToMag(*gpu)
        t = _timeit(torch, lambda: capi.check(lib.clb200_unary_launch_device(blk._h, a.data_ptr(), b.data_ptr(), n, stream_ptr)))
        out["clComplexToMag"] = hbm(12 * n, t, n)
        a.abs_().add_(1e-3)
        blk = blocks.clLog(*gpu, 10.0, 0.0)
        t = _timeit(torch, lambda: capi.check(lib.clb200_unary_launch_device(blk._h, a.data_ptr(), b.data_ptr(), 2 * n, stream_ptr)))
        out["clLog10"] = hbm(8 * 2 * n, t, 2 * n)
        a.uniform_(-1, 1)
    except Exception as e:                           # noqa: BLE001
        out["elementwise"] = {"error": str(e)}
    try:
        taps = np.zeros(256, np.float32)
        taps[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
        for name, ut in (("clFilter_fft_256tap", False), ("clFilter_fir_256tap", True)):
            blk = blocks.clFilter(*gpu, 1, taps, 1, 0, ut)
            t = _timeit(torch, lambda: blk.launch_device(a.data_ptr(), n, b.data_ptr(), stream_ptr), 3)
            out[name] = hbm(16 * n, t, n)
            if fp and "ffma_tflops" in fp:
                if ut:
                    # 256 taps x (re, im) = 512 FFMA per output sample
                    fl = 2.0 * 512 * n / t / 1e12
                    out[name]["roofline_fp32"] = {"bound": "fp32 FFMA pipe", "achieved": fl, "peak": fp["ffma_tflops"],
                                                  "unit": "TFLOP/s", "frac": fl / fp["ffma_tflops"],
                                                  "flop_per_sample": 1024}
                else:
                    # executed FP32 lane operations per sample, counted from the SASS histogram of the ncu capture of
                    # this launch shape (profiles/r2_fftfilt_sass_hist.txt: 2 x FADD2 + FMUL + FFMA + FADD warp
                    # instructions x 32 lanes / 64 Mi samples); the nominal count (two 1024-point FFTs at 5 N log2 N
                    # + 1024 complex multiplies per 769 new samples) is given beside it
                    lane_ops = (2 * 55851520 + 40143280 + 30369264 + 9774016) * 32.0 / (1 << 26)
                    tl = lane_ops * n / t / 1e12
                    out[name]["roofline_fp32"] = {"bound": "fp32 lanes (one add, multiply or fma per lane and clock)",
                                                  "achieved": tl, "peak": fp["fadd2_tflops"], "unit": "T lane-ops/s",
                                                  "frac": tl / fp["fadd2_tflops"], "lane_ops_per_sample": lane_ops,
                                                  "lane_ops_source": "profiles/r2_fftfilt_sass_hist.txt",
                                                  "nominal_flop_per_sample": (2 * 5 * 1024 * 10 + 6 * 1024) / 769.0}
    except Exception as e:                           # noqa: BLE001
        out["clFilter"] = {"error": str(e)}
    try:
        M = 64
        ptaps = np.zeros(128, np.float32)
        ptaps[:127] = orc.firdes_low_pass_hamming(1.0, 64.0, 0.5, 1.21)
        niter = (n - 128) // M
        blk = blocks.clPolyphaseChannelizer(*gpu, ptaps, 65536, M, M, list(range(M)))
        t = _timeit(torch, lambda: blk.launch_device(a.data_ptr(), b.data_ptr(), niter, stream_ptr), 3)
        out["clPolyphaseChannelizer_64ch"] = hbm(16 * niter * M, t, niter * M)
    except Exception as e:                           # noqa: BLE001
        out["clPolyphaseChannelizer_64ch"] = {"error": str(e)}
    try:
        # clFFT outside the one-kernel sizes: 65536 points (two passes of column transforms), 1000 points (not a power
        # of two: fused chirp-z over 2048-point transforms) and 10000 points (chirp-z over 32768-point two-pass plans);
        # 16 B/sample algorithmic like the headline
        for name, N in (("clFFT_65536pt_two_pass", 65536), ("clFFT_1000pt_chirpz_two_kernels", 1000),
                        ("clFFT_10000pt_chirpz_five_kernels", 10000)):
            nv = (1 << 24) // N if N != 65536 else n // N
            blk = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *gpu)
            t = _timeit(torch, lambda: blk.launch_device(a.data_ptr(), b.data_ptr(), nv, stream_ptr), 3)
            out[name] = hbm(16 * nv * N, t, nv * N)
    except Exception as e:                           # noqa: BLE001
        out["clFFT_other_sizes"] = {"error": str(e)}
    del a, b
    try:
        A, F, T, K = 32, 1024, 1024, 16
        nbl = A * (A + 1) // 2
        per = T * A * F * 2
        buf = torch.randint(-127, 128, (per * K,), dtype=torch.int8, device="cuda")        # 1 GiB > L2
        vis = torch.empty(K * F * nbl * 2, dtype=torch.float32, device="cuda")
        blk = blocks.clXEngine(*gpu, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
        bytes_ = per + F * nbl * 8

        def xe_entry(t):
            return {"Msamples_s": A * F * T / t / 1e6, "us_per_integration": t * 1e6, "GBps": bytes_ / t / 1e9,
                    "frac_hbm": bytes_ / t / 1e9 / hbm_peak, "int8_TOPS": 2.0 * 64 * 64 * T * F / t / 1e12}

        def single():
            for k in range(K):
                blk.launch_device(buf.data_ptr() + k * per, vis.data_ptr() + k * F * nbl * 8, False, stream_ptr)
        out["clXEngine_32st_1024ch_int1024"] = xe_entry(_timeit(torch, single, 4) / K)
        out["clXEngine_32st_1024ch_int1024"]["launch"] = "one launch per integration (PDL overlaps consecutive launches)"
        e = xe_entry(_timeit(torch, lambda: blk.launch_device_batch(buf.data_ptr(), vis.data_ptr(), K, stream_ptr), 4) / K)
        e["launch"] = "clb200_xengine_launch_device_batch: %d integrations per grid, persistent CTAs" % K
        out["clXEngine_32st_1024ch_int1024_batch%d" % K] = e
        del buf, vis
        # complex-float input (the reference's default DTYPE_COMPLEX): 8 B per sample in, FP32 arithmetic
        Fc = 256
        xc = torch.empty(T * A * Fc * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
        visc = torch.empty(Fc * nbl * 2, dtype=torch.float32, device="cuda")
        blkc = blocks.clXEngine(*gpu, False, capi.DTYPE_COMPLEX, 1, A, 1, 0, Fc, T, [])
        t = _timeit(torch, lambda: blkc.launch_device(xc.data_ptr(), visc.data_ptr(), False, stream_ptr), 4)
        bc = T * A * Fc * 8 + Fc * nbl * 8
        out["clXEngine_32st_256ch_int1024_complex"] = {
            "Msamples_s": A * Fc * T / t / 1e6, "us_per_integration": t * 1e6, "GBps": bc / t / 1e9,
            "frac_hbm": bc / t / 1e9 / hbm_peak, "fp32_TFLOPs": 8.0 * nbl * T * Fc / t / 1e12}
        del xc, visc
    except Exception as e:                           # noqa: BLE001
        out["clXEngine_32st_1024ch_int1024"] = {"error": str(e)}
    return out


def cpu_blocks(cores):
    """The reference CPU path of the other BASELINE configs, restated by the oracle, each on a bounded
    sample (~1 s): the Msamples/s that sit beside the device-resident `blocks` figures."""
    import numpy as np
    from oracle import oracle as orc
    res = {}

    def rate(fn, nsamp, min_s=1.0):
        fn()
        reps, t0 = 0, time.perf_counter()
        while True:
            fn()
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_s:
                return nsamp * reps / dt / 1e6

    try:
        # config 1: MultiplyConst c32[8192] per call (1 thread, as the reference) and a 64 Mi-sample stream on all threads
        x = orc.rng_c32(8192, orc.SEED_M)
        orc.lib().orc_set_threads(1)
        res["clMultiplyConst_8192_per_call"] = {"Msamples_s": rate(lambda: orc.mathconst(x, 2.0, 1), 8192, 0.5), "cores": 1,
                                                "what": "clMathConst_impl::testCPU loop, one 8192-sample buffer per call"}
        xl = orc.rng_c32(1 << 24, orc.SEED_M)
        for th in (1, cores):
            orc.lib().orc_set_threads(th)
            res["clMultiplyConst_stream_%dthread" % th] = {"Msamples_s": rate(lambda: orc.mathconst(xl, 2.0, 1), xl.size, 0.5),
                                                           "cores": th, "what": "same loop over 16 Mi samples per call"}
        del xl
        orc.lib().orc_set_threads(1)                       # clFilter's CPU paths are single-threaded (nthreads ignored)
        taps = np.zeros(256, np.float32)
        taps[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
        n = 257 * 2048
        x = orc.rng_c32(n, orc.SEED_L)
        f = orc.lib().orc_fftfilt_create(taps, 256, 1)
        out = np.zeros(n, np.complex64)
        res["clFilter_fft_256tap"] = {"Msamples_s": rate(lambda: orc.lib().orc_fftfilt_filter(f, n, orc._f(x), orc._f(out)), n),
                                      "cores": 1, "what": "fft_filter_ccf::filter restated (fftsize 512 / 257 new samples)"}
        orc.lib().orc_fftfilt_destroy(f)
        xh = orc.rng_c32(1 << 16, orc.SEED_L)
        res["clFilter_fir_256tap"] = {"Msamples_s": rate(lambda: orc.fir(xh, taps, 1), xh.size - 255), "cores": 1,
                                      "what": "fir_filter_ccf::filterN restated"}
        orc.lib().orc_set_threads(cores)
        M = 64
        ptaps = np.zeros(128, np.float32)
        ptaps[:127] = orc.firdes_low_pass_hamming(1.0, 64.0, 0.5, 1.21)
        niter = 8192
        xp = orc.rng_c32((niter - 1) * M + 128, orc.SEED_P)
        res["clPolyphaseChannelizer_64ch"] = {"Msamples_s": rate(lambda: orc.pfb(xp, ptaps, M, M, list(range(M)), niter), niter * M),
                                              "cores": cores, "what": "filterpfb2 + DFT + map restated (the reference has no CPU path)"}
        A, F, T = 32, 16, 1024
        buf = orc.rng_i8(T * A * F * 2, orc.SEED_X)
        res["clXEngine_32st_int1024"] = {"Msamples_s": rate(lambda: orc.xengine_f32(buf, A, F, T, 1), A * F * T),
                                         "cores": cores, "what": "CharToComplex + XCorrelate restated on a 16-channel slab "
                                                                  "(the reference has no CPU X-engine)"}
        A2 = 16
        buf2 = orc.rng_i8(T * A2 * F * 2 * 2, orc.SEED_X + 1)
        res["clXEngine_16st_2pol_int1024"] = {"Msamples_s": rate(lambda: orc.xengine_f32(buf2, A2, F, T, 2), A2 * 2 * F * T),
                                              "cores": cores, "what": "same, two polarisations (4 products per baseline), "
                                                                       "samples = stations x pols x channels x time"}
    except Exception as e:                               # noqa: BLE001
        res["error"] = str(e)
    return res


def copy_ceiling(torch, dist, nbytes, reps=4):
    """bare pinned H2D + D2H copies of nbytes each way, both directions at once, all ranks at once (no kernel):
    the host-link ceiling the e2e figure is judged against.  GB/s per direction per rank (slowest rank)."""
    n = int(nbytes)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def once():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)
    once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    return n * reps / dt / 1e9


def xengine_host(blocks, capi, dev, torch, dist, world, F, what):
    """clXEngine (32 stations, `F` channels on this rank, integration 1024, IChar) through its HOST entry points:
      whole_buffer  clb200_xengine_work: one pinned integration buffer in, visibilities out, one blocking call
      stream        clb200_xengine_stream_begin / push_timesteps / poll_result: 256 time steps per push from 32
                    page-locked port buffers (DMA'd in place), results picked up asynchronously, NI integrations
                    pipelined -- the general_work() shape of the reference block
    Wall clock around whole calls, max over ranks; every matrix is checked against a device launch on the same data."""
    import numpy as np
    A, T, NI, PUSH = 32, 1024, 6, 256
    nbl = A * (A + 1) // 2
    res = {"what": what, "channels_per_rank": F, "ranks": world}
    row = F * 2
    # per-station streams, NI integrations each: [NI*T][F][re,im] int8
    ports = [torch.randint(-127, 128, (NI * T * row,), dtype=torch.int8).pin_memory() for _ in range(A)]
    lib = capi.load()

    def sync_max(dt):
        if dist is None:
            return dt
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # reference result of the LAST integration: the same data gathered into [t][station][chan] on the device
    last = torch.stack([p.view(NI, T, row)[NI - 1] for p in ports], dim=1).contiguous()          # [T][A][row]
    d_last = last.cuda()
    d_vis = torch.empty(F * nbl * 2, dtype=torch.float32, device="cuda")
    chk = blocks.clXEngine(1, 2, 0, dev, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    chk.launch_device(d_last.data_ptr(), d_vis.data_ptr(), False, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = d_vis.cpu().numpy().view(np.complex64)

    # ---- streaming ingest ----
    blk = blocks.clXEngine(1, 2, 0, dev, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    blk.stream_begin(0, 4)
    out = np.zeros(F * nbl, np.complex64)
    ready = C.c_int(0)
    arr = (C.c_void_p * A)()

    def run_stream():
        got = 0
        for pos in range(0, NI * T, PUSH):
            for s in range(A):
                arr[s] = ports[s].data_ptr() + pos * row
            capi.check(lib.clb200_xengine_push_timesteps(blk._h, arr, A, PUSH))
            while True:                                                  # pick up what has finished; wait only when the
                behind = (pos + PUSH) // T - got >= 3                    # 4-slot result ring is about to fill
                capi.check(lib.clb200_xengine_poll_result(blk._h, out.ctypes.data_as(C.c_void_p), int(behind), C.byref(ready)))
                if not ready.value:
                    break
                got += 1
        while got < NI:
            capi.check(lib.clb200_xengine_poll_result(blk._h, out.ctypes.data_as(C.c_void_p), 1, C.byref(ready)))
            got += ready.value
    for key, stable in (("stream", True), ("stream_ports_returned_each_push", False)):
        blk.stream_ports_stable(stable)
        run_stream()                                                     # warm-up (allocations, first touch)
        if dist is not None:
            dist.barrier()
        t0 = time.perf_counter()
        run_stream()
        dt = sync_max(time.perf_counter() - t0)
        ok = bool(np.array_equal(out, want))
        st = blk.stream_state()
        res[key] = {"us_per_integration": dt / NI * 1e6, "Msamples_s": world * A * F * T * NI / dt / 1e6,
                    "h2d_GBps_per_rank": T * A * row * NI / dt / 1e9, "integrations": NI, "timesteps_per_push": PUSH,
                    "pushes_that_waited_for_the_gpu": st["pushes_blocked"], "matches_device_launch": ok,
                    "api": "clb200_xengine_push_timesteps / poll_result, page-locked ports DMA'd in place; " +
                           ("the ports stay untouched until their result is polled (stream_ports_stable): uploads overlap the next push"
                            if stable else "push returns once the DMA has read the ports (scheduler-owned buffers)")}
    blk.stream_end()
    # the same with pageable ports (marshal -> pinned staging -> DMA), one integration's worth
    pports = [np.ascontiguousarray(p.view(NI, T * row)[NI - 1].numpy()).copy() for p in ports]
    blk2 = blocks.clXEngine(1, 2, 0, dev, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    blk2.stream_begin(0, 4)

    def run_pageable(n):
        for _ in range(n):
            for pos in range(0, T, PUSH):
                for s in range(A):
                    arr[s] = pports[s].ctypes.data + pos * row
                capi.check(lib.clb200_xengine_push_timesteps(blk2._h, arr, A, PUSH))
            capi.check(lib.clb200_xengine_poll_result(blk2._h, out.ctypes.data_as(C.c_void_p), 1, C.byref(ready)))
    run_pageable(1)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    run_pageable(3)
    dt = sync_max(time.perf_counter() - t0)
    res["stream_pageable_ports"] = {"us_per_integration": dt / 3 * 1e6, "Msamples_s": world * A * F * T * 3 / dt / 1e6,
                                    "matches_device_launch": bool(np.array_equal(out, want)),
                                    "api": "same, pageable ports: host marshal into pinned staging (the reference's memcpy)"}
    blk2.stream_end()
    # ---- whole-buffer call ----
    hin = last.pin_memory()
    hout = torch.empty(F * nbl * 2, dtype=torch.float32).pin_memory()
    ip, op = C.c_void_p(hin.data_ptr()), C.c_void_p(hout.data_ptr())
    capi.check(lib.clb200_xengine_work(chk._h, ip, op, 0))
    if dist is not None:
        dist.barrier()
    reps, t0 = 4, time.perf_counter()
    for _ in range(reps):
        capi.check(lib.clb200_xengine_work(chk._h, ip, op, 0))
    dt = sync_max(time.perf_counter() - t0) / reps
    res["whole_buffer"] = {"us_per_integration": dt * 1e6, "Msamples_s": world * A * F * T / dt / 1e6,
                           "h2d_bytes": T * A * row, "d2h_bytes": F * nbl * 8,
                           "matches_device_launch": bool(np.array_equal(hout.numpy().view(np.complex64), want)),
                           "api": "clb200_xengine_work (pinned host in/out, blocking)"}
    return res


def per_call(blocks, capi, dev, n=8192, iters=300):
    """BASELINE configs[0]: scheduler-sized work() calls (8192 gr_complex, pageable host buffers,
    1 warm-up + N timed calls like lib/test_clenabled.cc:1237-1251), through the C ABI."""
    import numpy as np
    from oracle import oracle as orc
    lib = capi.load()
    x = orc.rng_c32(n, orc.SEED_M)
    out = np.zeros(n, np.complex64)
    xp, op = C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data)
    res = {}
    mc = blocks.clMathConst(capi.DTYPE_COMPLEX, 1, 2, 0, dev, 2.0, capi.OP_MULTIPLY)
    ff = blocks.clFFT(n, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 2, 0, dev)
    for name, fn in (("clMultiplyConst", lambda: lib.clb200_mathconst_work(mc._h, xp, op, n)),
                     ("clFFT_8192", lambda: lib.clb200_fft_work(ff._h, xp, op, 1))):
        fn()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        dt = (time.perf_counter() - t0) / iters
        res[name] = {"us_per_call": dt * 1e6, "Msamples_s": n / dt / 1e6}
    # the reference CPU loop on the same buffer (clMathConst_impl::testCPU, 1 thread)
    orc.lib().orc_set_threads(1)
    orc.mathconst(x, 2.0, 1)
    t0 = time.perf_counter()
    for _ in range(iters):
        orc.mathconst(x, 2.0, 1)
    dt = (time.perf_counter() - t0) / iters
    res["cpu_clMultiplyConst_1thread"] = {"us_per_call": dt * 1e6, "Msamples_s": n / dt / 1e6}
    return res


def xengine_multi_gpu(torch, dist, blocks, capi, multigpu, rank, world, local, sp, stream):
    """BASELINE config 5 on N GPUs (32 stations x 1024 channels x 1024 time steps, IChar, channels sharded)"""
    import numpy as np
    extra = {}
    A, F, T = 32, 1024, 1024
    f0, fc = multigpu.shard_channels(F, rank, world)
    nbl = A * (A + 1) // 2
    g = torch.Generator(device="cuda")
    g.manual_seed(4321 + rank)
    bufs = [torch.randint(-127, 128, (T * A * fc * 2,), dtype=torch.int8, device="cuda", generator=g) for _ in range(4)]
    slab = torch.empty(fc * nbl * 2, dtype=torch.float32, device="cuda")
    xe = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def timed(step, n=20):
        for i in range(4):
            step(i)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(stream)
        for i in range(n):
            step(i)
        x1.record(stream)
        barrier()
        t = torch.tensor([x0.elapsed_time(x1) / n], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3                         # us

    # the unsharded truth of integration 0: every rank's slab of bufs[0], gathered by NCCL once
    xe.launch_device(bufs[0].data_ptr(), slab.data_ptr(), False, sp)
    want = multigpu.gather_visibilities(slab, F, nbl * 2).cpu().numpy().view(np.complex64)

    def nccl_step(i):
        xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp)
        return multigpu.gather_visibilities(slab, F, nbl * 2)
    us = timed(nccl_step)
    extra["clXEngine_sharded_nccl_all_gather"] = {
        "us_per_integration_to_consumable": us, "Msamples_s": A * F * T / us, "gpus": world,
        "collective": "all_gather of %d B visibility slabs per rank (NCCL)" % (slab.numel() * 4)}
    us = timed(lambda i: xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp))
    extra["clXEngine_sharded_no_gather"] = {"us_per_integration": us, "Msamples_s": A * F * T / us, "gpus": world,
                                            "collective": "none: every rank keeps its own sub-band (the reference's first_channel/num_channels deployment)"}
    for name, cls in (("peer_memory", multigpu.PeerGather), ("nvswitch_multicast", multigpu.MulticastGather)):
        key = "clXEngine_sharded_fused_gather_" + name
        try:
            xg = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
            xg.set_shard(F, f0)
            pg = cls(xg, local, F, nbl)
            xg.launch_device_gather(bufs[0].data_ptr(), sp)
            xg.gather_wait(sp)
            torch.cuda.synchronize()                         # no barrier: the device-side flags are the completion signal
            ok = torch.tensor([1 if np.array_equal(pg.result(), want) else 0], device="cuda")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            dist.barrier()

            def fstep(i):
                xg.launch_device_gather(bufs[i % 4].data_ptr(), sp)
                xg.gather_wait(sp)
            us = timed(fstep)
            us_nw = timed(lambda i: xg.launch_device_gather(bufs[i % 4].data_ptr(), sp))
            extra[key] = {"us_per_integration_to_consumable": us, "us_per_integration_pipelined": us_nw,
                          "Msamples_s": A * F * T / us, "gpus": world, "full_matrix_on_every_rank_equals_unsharded": bool(ok.item()),
                          "collective": "none: the epilogue stores the slab into every rank's matrix (%s), completion through "
                                        "device-side flags (clb200_xengine_gather_wait)" %
                                        ("16 B peer stores over NVLink" if name == "peer_memory" else "one multimem.st per 16 B, replicated by the switch")}
            pg.close()
        except Exception as e:                           # noqa: BLE001
            extra[key] = {"error": str(e)[:300]}
    del bufs, slab
    torch.cuda.empty_cache()
    # weak scaling, device-resident: 1024 channels PER GPU, batches of 8 integrations, no gather
    try:
        K = 8
        per = T * A * F * 2
        buf = torch.randint(-127, 128, (per * K,), dtype=torch.int8, device="cuda")
        vis = torch.empty(K * F * nbl * 2, dtype=torch.float32, device="cuda")
        xw = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
        us = timed(lambda i: xw.launch_device_batch(buf.data_ptr(), vis.data_ptr(), K, sp), 10) / K
        extra["clXEngine_weak_1024ch_per_gpu_batch%d" % K] = {
            "us_per_integration_per_gpu": us, "Msamples_s": world * A * F * T / us, "gpus": world, "scaling": "weak",
            "channels_total": world * F, "collective": "none (sub-bands stay on their GPU)"}
        del buf, vis
        torch.cuda.empty_cache()
    except Exception as e:                               # noqa: BLE001
        extra["clXEngine_weak_1024ch_per_gpu"] = {"error": str(e)[:300]}
    # host-fed: every rank ingests ITS sub-band (F/N channels) through push_timesteps from its own page-locked ports
    try:
        extra["clXEngine_host_fed_sharded"] = xengine_host(blocks, capi, local, torch, dist, world, fc,
                                                           "32 stations x %d of 1024 channels per rank, integration 1024" % fc)
    except Exception as e:                               # noqa: BLE001
        extra["clXEngine_host_fed_sharded"] = {"error": str(e)[:300]}
    return extra


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm is meant to use all host threads (measured:
        # pocketfft drops from 430 to 100 Msamples/s on 8 cores with it set), so it is dropped before numpy / scipy /
        # the OpenMP runtime of the oracle are loaded
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
            os.environ.pop(k, None)
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    from gr_clenabled_b200 import blocks, capi

    capi.require_gpu()                       # no CPU fallback: fail loudly without the CUDA path
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = peaks()

    nvec, N = args.nvec, FFT_N
    nsamp = nvec * N
    # device-resident batch (synthetic uniform [-1,1) gr_complex), in != out, 2 x 512 MiB > L2
    x = torch.empty(nsamp * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    y = torch.empty_like(x)
    fft = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, 1, 2, 0, local)
    stream = torch.cuda.current_stream()
    sp = stream.cuda_stream

    def step():
        fft.launch_device(x.data_ptr(), y.data_ptr(), nvec, sp)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    l0 = fft.counters()["launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark(0)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.mark(1)
    ms = e0.elapsed_time(e1)
    launches = fft.counters()["launches"] - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * nsamp / (ms_per_step / 1e3) / 1e6          # Msamples/s, all ranks
    kernel_s = ms_per_step / 1e3                                # one launch per step
    achieved = BYTES_PER_SAMPLE * nsamp / kernel_s / 1e9

    # parity spot check of what was just timed (first and last vector) against numpy's FFT
    yh = y.view(torch.complex64)
    for v in (0, nvec - 1):
        got = yh[v * N:(v + 1) * N].cpu().numpy()
        want = np.fft.fft(x.view(torch.complex64)[v * N:(v + 1) * N].cpu().numpy().astype(np.complex128))
        err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        if not err < 1e-5:
            raise SystemExit("bench: FFT output does not match (rel err %g)" % err)

    # ---- e2e: pinned host buffers through clb200_fft_work (H2D + kernel + D2H timed) ----
    hx = torch.empty(nsamp * 2, dtype=torch.float32).pin_memory()
    hy = torch.empty(nsamp * 2, dtype=torch.float32).pin_memory()
    hx.uniform_(-1, 1)
    for _ in range(2):
        fft.work_ptr(hx.data_ptr(), hy.data_ptr(), nvec)
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fft.work_ptr(hx.data_ptr(), hy.data_ptr(), nvec)       # returns with outputs on the host
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_val = world * nsamp * e2e_steps / dt / 1e6
    got = hy.view(torch.complex64)[:N].numpy()
    want = np.fft.fft(hx.view(torch.complex64)[:N].numpy().astype(np.complex128))
    if not float(np.max(np.abs(got - want)) / np.max(np.abs(want))) < 1e-5:
        raise SystemExit("bench: e2e FFT output does not match")
    del hx, hy
    # the host-link ceiling for the same bytes on the same ranks: bare pinned copies both ways at once, no kernel
    try:
        ceil_gbs = copy_ceiling(torch, dist, nsamp * 8)
        ceil_msps = world * ceil_gbs * 1e9 / 8 / 1e6
        ceiling = {"GBps_per_direction_per_rank": ceil_gbs, "Msamples_s_all_ranks": ceil_msps,
                   "frac_of_copy_ceiling": e2e_val / ceil_msps,
                   "how": "pinned cudaMemcpyAsync H2D + D2H of the step's bytes, both directions and all ranks at once, no kernel"}
    except Exception as e:                               # noqa: BLE001
        ceiling = {"error": str(e)[:200]}

    extra = {}
    if world > 1 and not args.no_blocks:
        from gr_clenabled_b200 import multigpu
        del x, y
        torch.cuda.empty_cache()
        extra = xengine_multi_gpu(torch, dist, blocks, capi, multigpu, rank, world, local, sp, stream)
    if rank == 0 and world == 1 and not args.no_blocks:
        del x, y
        torch.cuda.empty_cache()
        extra = secondary_blocks(torch, blocks, capi, local, sp, hbm_peak)
        extra["per_call_8192_pageable"] = per_call(blocks, capi, local)
        try:
            extra["clXEngine_host"] = xengine_host(blocks, capi, local, torch, None, 1, 1024,
                                                   "BASELINE config 5 on one GPU: 32 stations x 1024 channels, integration 1024")
        except Exception as e:                           # noqa: BLE001
            extra["clXEngine_host"] = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        rates = {}
        for impl in ("pocketfft", "oracle_radix2"):
            try:
                rates[impl + "_1thread"] = cpu_fft_rate(impl, 512, 1, 2.0)[0]
                rates[impl + "_%dthreads" % cores] = cpu_fft_rate(impl, 2048, cores, 4.0)[0]
            except Exception as e:                       # noqa: BLE001
                rates[impl + "_error"] = str(e)[:200]
        best = max((k for k in rates if k.endswith("threads")), key=lambda k: rates[k], default=None)
        cpu = {"value": rates.get(best), "unit": "Msamples/s", "cores": cores, "kind": "port",
               "implementation": best, "all": rates,
               "sample": "2048 vectors of 8192 per call for ~4 s per implementation (512 vectors, ~2 s on one thread); "
                         "clFFT_impl::testCPU's loop with the transform by pocketfft (scipy.fft complex64) and by the "
                         "oracle's radix-2 FFT -- FFTW3f of the reference is absent from the image; value = the faster"}

    if rank == 0:
        traffic, traffic_src = ncu_traffic(nvec)
        line = {
            "metric": METRIC,
            "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(nvec, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "k_fft<13,...>", "bytes_per_launch": BYTES_PER_SAMPLE * nsamp},
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": nsamp * 8,
                    "d2h_bytes_per_step": nsamp * 8, "steps": e2e_steps,
                    "api": "clb200_fft_work (pinned host in/out)", "copy_ceiling": ceiling},
            "gpu_launches": int(launches) * world,
            "clocks": clocks,
        }
        if cpu:
            line["cpu_baseline"] = cpu
            if extra is not None:
                extra["cpu_reference_paths"] = cpu_blocks(cpu["cores"])
        if extra:
            line["blocks"] = extra
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
