#!/usr/bin/env python
"""bench.py -- the headline benchmark: clFFT forward 8192-pt gr_complex (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the clFFT hot path over one batch of synthetic vectors
(`--nvec` vectors of 8192 gr_complex per GPU; default 8192 vectors = 512 MiB in + 512 MiB
out, larger than the 126 MB L2, so no L2 flush is needed between steps).

  value     device-resident throughput (inputs in HBM), Msamples/s over all ranks, CUDA events
  e2e       same metric through the C ABI with pinned HOST buffers: clb200_fft_work() does
            H2D -> kernel -> D2H inside the timed region
  roofline  the FFT kernel against MEASURED_PEAKS.json's HBM copy bandwidth, 16 B/sample
  cpu_baseline  the oracle (C restatement of the reference CPU path) on the host cores
  blocks    device-resident throughput of the other hot-path blocks (secondary, same run)

Multi-GPU: clFFT streams/vectors are independent (lib/clFFT_impl.cc:537-541) so ranks
shard vectors with no collective ("weak" scaling: per-GPU work fixed).  torch is used only for
device memory, streams/events and torch.distributed; every kernel is ours.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FFT_N = 8192
BYTES_PER_SAMPLE = 16          # 8 B read + 8 B written per gr_complex sample


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


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(nvec):
    """dram bytes per launch of the FFT kernel from the committed ncu --set full capture
    (profiles/r1_traffic.json); only valid for the launch shape that was captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)["k_fft_8192pt_x8192vec"]
        return t["bytes"] if nvec == 8192 else None
    except Exception:
        return None


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
# reference arm / cpu baseline: the oracle's restatement of clFFT_impl::testCPU
# (lib/clFFT_impl.cc:464-518; FFTW3f is not in the image) on the host cores
# --------------------------------------------------------------------------------
def cpu_fft_run(nvec, reps, threads):
    import numpy as np
    from oracle import oracle as orc
    orc.lib().orc_set_threads(threads)
    x = orc.rng_c32(FFT_N * nvec, orc.SEED_F)
    orc.fft(x[:FFT_N * min(nvec, 64)], FFT_N, -1)           # warm
    t0 = time.perf_counter()
    for _ in range(reps):
        orc.fft(x, FFT_N, -1)
    dt = time.perf_counter() - t0
    return FFT_N * nvec * reps / dt / 1e6, dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (restated; kind 'port') on all host cores."""
    if rank != 0:
        return
    from oracle import oracle as orc
    cores = os.cpu_count() or 1
    orc.lib().orc_set_threads(cores)
    threads = orc.lib().orc_num_threads()
    nvec = args.nvec                                       # the same batch per step as our arm (8192 vectors)
    import numpy as np
    x = orc.rng_c32(FFT_N * nvec, orc.SEED_F)
    for _ in range(max(1, min(args.warmup, 3))):
        orc.fft(x, FFT_N, -1)
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.fft(x, FFT_N, -1)
    dt = time.perf_counter() - t0
    val = FFT_N * nvec * steps / dt / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/sec per block (clFFT forward 8192-pt)",
        "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "clFFT forward 8192-pt gr_complex, 1 stream (BASELINE configs[1])",
                   "fft_size": FFT_N, "vectors_per_step": nvec},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port",
                         "sample": "%d vectors of 8192 per step, oracle radix-2 FFT (FFTW3f of the "
                                   "reference is not in the image), OpenMP over vectors" % nvec},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------
def secondary_blocks(torch, blocks, capi, dev, stream_ptr, hbm_peak):
    """device-resident throughput of the other blocks, a few launches each (not the headline)"""
    import numpy as np
    from oracle import oracle as orc
    out = {}
    gpu = (1, 2, 0, dev)

    def timeit(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters / 1e3

    n = 1 << 26                                     # 64 Mi complex samples = 512 MiB
    a = torch.empty(n * 2, dtype=torch.float32, device="cuda").uniform_(-1, 1)
    b = torch.empty(n * 2, dtype=torch.float32, device="cuda")
    try:
        blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *gpu, 0.7071, capi.OP_MULTIPLY)
        t = timeit(lambda: blk.launch_device(a.data_ptr(), b.data_ptr(), n, stream_ptr))
        out["clMultiplyConst"] = {"Msamples_s": n / t / 1e6, "GBps": 16 * n / t / 1e9, "frac_hbm": 16 * n / t / 1e9 / hbm_peak}
    except Exception as e:                           # noqa: BLE001
        out["clMultiplyConst"] = {"error": str(e)}
    try:
        taps = np.zeros(256, np.float32)
        taps[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
        for name, ut in (("clFilter_fft_256tap", False), ("clFilter_fir_256tap", True)):
            blk = blocks.clFilter(*gpu, 1, taps, 1, 0, ut)
            t = timeit(lambda: blk.launch_device(a.data_ptr(), n, b.data_ptr(), stream_ptr), 3)
            out[name] = {"Msamples_s": n / t / 1e6, "GBps": 16 * n / t / 1e9, "frac_hbm": 16 * n / t / 1e9 / hbm_peak}
    except Exception as e:                           # noqa: BLE001
        out["clFilter"] = {"error": str(e)}
    try:
        M = 64
        ptaps = np.zeros(128, np.float32)
        ptaps[:127] = orc.firdes_low_pass_hamming(1.0, 64.0, 0.5, 1.21)
        niter = (n - 128) // M
        blk = blocks.clPolyphaseChannelizer(*gpu, ptaps, 65536, M, M, list(range(M)))
        t = timeit(lambda: blk.launch_device(a.data_ptr(), b.data_ptr(), niter, stream_ptr), 3)
        out["clPolyphaseChannelizer_64ch"] = {"Msamples_s": niter * M / t / 1e6, "GBps": 16 * niter * M / t / 1e9,
                                              "frac_hbm": 16 * niter * M / t / 1e9 / hbm_peak}
    except Exception as e:                           # noqa: BLE001
        out["clPolyphaseChannelizer_64ch"] = {"error": str(e)}
    del a, b
    try:
        A, F, T = 32, 1024, 1024
        nb = T * A * F * 2
        bufs = [torch.randint(-127, 128, (nb,), dtype=torch.int8, device="cuda") for _ in range(4)]   # 256 MiB > L2
        vis = torch.empty(F * (A * (A + 1) // 2) * 2, dtype=torch.float32, device="cuda")
        blk = blocks.clXEngine(*gpu, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
        it = [0]
        def f():
            blk.launch_device(bufs[it[0] % 4].data_ptr(), vis.data_ptr(), False, stream_ptr)
            it[0] += 1
        t = timeit(f, 32)
        bytes_ = nb + vis.numel() * 4
        out["clXEngine_32st_1024ch_int1024"] = {
            "Msamples_s": A * F * T / t / 1e6, "us_per_integration": t * 1e6, "GBps": bytes_ / t / 1e9,
            "frac_hbm": bytes_ / t / 1e9 / hbm_peak, "int8_TOPS": 2.0 * 64 * 64 * T * F / t / 1e12}
    except Exception as e:                           # noqa: BLE001
        out["clXEngine_32st_1024ch_int1024"] = {"error": str(e)}
    return out


def cpu_blocks(cores):
    """The reference CPU path of the other BASELINE configs, restated by the oracle, each on a bounded
    sample (~1-2 s): the Msamples/s that sit beside the device-resident `blocks` figures."""
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
    except Exception as e:                               # noqa: BLE001
        res["error"] = str(e)
    return res


def xengine_e2e(blocks, capi, dev, torch):
    """clXEngine through the host entry point (clb200_xengine_work: pinned host integration buffer in,
    visibilities out), BASELINE config 5 on one GPU."""
    import numpy as np
    A, F, T = 32, 1024, 1024
    nb = T * A * F * 2
    src = torch.randint(-127, 128, (nb,), dtype=torch.int8).pin_memory()
    out = torch.empty(F * (A * (A + 1) // 2) * 2, dtype=torch.float32).pin_memory()
    blk = blocks.clXEngine(1, 2, 0, dev, False, capi.DTYPE_BYTE, 1, A, 1, 0, F, T, [])
    lib = capi.load()
    ip, op = C.c_void_p(src.data_ptr()), C.c_void_p(out.data_ptr())
    capi.check(lib.clb200_xengine_work(blk._h, ip, op, 0))
    reps, t0 = 5, time.perf_counter()
    for _ in range(reps):
        capi.check(lib.clb200_xengine_work(blk._h, ip, op, 0))
    dt = (time.perf_counter() - t0) / reps
    return {"us_per_integration": dt * 1e6, "Msamples_s": A * F * T / dt / 1e6, "h2d_bytes": nb, "d2h_bytes": out.numel() * 4,
            "api": "clb200_xengine_work (pinned host in/out)"}


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


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
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

    extra = {}
    if world > 1 and not args.no_blocks:
        # clXEngine sharded by channel: every rank correlates its slab, the slabs meet
        # in one NCCL all_gather; device-resident slabs, 4 rotating buffers per rank
        from gr_clenabled_b200 import multigpu
        del x, y
        torch.cuda.empty_cache()
        A, F, T = 32, 1024, 1024
        f0, fc = multigpu.shard_channels(F, rank, world)
        nbl = A * (A + 1) // 2
        bufs = [torch.randint(-127, 128, (T * A * fc * 2,), dtype=torch.int8, device="cuda") for _ in range(4)]
        slab = torch.empty(fc * nbl * 2, dtype=torch.float32, device="cuda")
        xe = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
        def xstep(i):
            xe.launch_device(bufs[i % 4].data_ptr(), slab.data_ptr(), False, sp)
            return multigpu.gather_visibilities(slab, F, nbl * 2)
        for i in range(3):
            full = xstep(i)
        barrier()
        x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        x0.record(stream)
        nx = 10
        for i in range(nx):
            full = xstep(i)
        x1.record(stream)
        barrier()
        t = torch.tensor([x0.elapsed_time(x1) / nx], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        us = float(t.item()) * 1e3
        extra["clXEngine_32st_1024ch_int1024_sharded"] = {
            "us_per_integration": us, "Msamples_s": A * F * T / us, "gpus": world,
            "collective": "all_gather of %d B visibility slabs per rank (NCCL)" % (slab.numel() * 4),
            "gathered_items": int(full.numel() // 2)}
        # the same step with the gather fused into the kernel: the epilogue stores this rank's slab into every
        # rank's full matrix over NVLink peer memory, no collective call
        try:
            xg = blocks.clXEngine(1, 2, 0, local, False, capi.DTYPE_BYTE, 1, A, 1, 0, fc, T, [])
            xg.set_shard(F, f0)
            pg = multigpu.PeerGather(xg, local, F, nbl)
            for i in range(3):
                xg.launch_device_gather(bufs[i % 4].data_ptr(), sp)
            barrier()
            x0.record(stream)
            for i in range(nx):
                xg.launch_device_gather(bufs[i % 4].data_ptr(), sp)
            x1.record(stream)
            barrier()
            t = torch.tensor([x0.elapsed_time(x1) / nx], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us2 = float(t.item()) * 1e3
            extra["clXEngine_32st_1024ch_int1024_sharded_fused_gather"] = {
                "us_per_integration": us2, "Msamples_s": A * F * T / us2, "gpus": world,
                "collective": "none: epilogue stores into every rank's matrix (cudaIpc peer memory over NVLink)"}
            pg.close()
        except Exception as e:                           # noqa: BLE001
            extra["clXEngine_32st_1024ch_int1024_sharded_fused_gather"] = {"error": str(e)}
    if rank == 0 and world == 1 and not args.no_blocks:
        del x, y
        torch.cuda.empty_cache()
        extra = secondary_blocks(torch, blocks, capi, local, sp, hbm_peak)
        extra["per_call_8192_pageable"] = per_call(blocks, capi, local)
        try:
            extra["clXEngine_e2e_host"] = xengine_e2e(blocks, capi, local, torch)
        except Exception as e:                           # noqa: BLE001
            extra["clXEngine_e2e_host"] = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        _, d1 = cpu_fft_run(512, 1, 1)
        v1, _ = cpu_fft_run(512, max(1, int(4.0 / max(d1, 1e-3))), 1)      # ~4 s on one thread
        _, dN = cpu_fft_run(2048, 1, cores)
        reps = max(2, int(10.0 / max(dN, 1e-3)))                            # ~10 s on all threads
        vN, dtN = cpu_fft_run(2048, reps, cores)
        cpu = {"value": vN, "unit": "Msamples/s", "cores": cores, "kind": "port",
               "single_thread_value": v1,
               "sample": "%d x %d vectors of 8192 (%.1f s); oracle radix-2 FFT restating "
                         "clFFT_impl::testCPU, FFTW3f absent from the image" % (reps, 2048, dtN)}

    if rank == 0:
        line = {
            "metric": "Msamples/sec per block (clFFT forward 8192-pt)",
            "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "clFFT forward 8192-pt gr_complex, 1 stream per GPU (BASELINE configs[1])",
                       "fft_size": N, "vectors_per_step_per_gpu": nvec,
                       "l2": "inputs larger than L2 (2 x %d MiB per step), no flush needed" % (nsamp * 8 >> 20),
                       "parallelism": "vectors sharded over %d GPU(s), no collective" % world},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": ncu_traffic(nvec), "peak_source": peak_src,
                         "kernel": "k_fft<13,...>", "bytes_per_launch": BYTES_PER_SAMPLE * nsamp},
            "e2e": {"value": e2e_val, "unit": "Msamples/s", "h2d_bytes_per_step": nsamp * 8,
                    "d2h_bytes_per_step": nsamp * 8, "steps": e2e_steps,
                    "api": "clb200_fft_work (pinned host in/out)"},
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
