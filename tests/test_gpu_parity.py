"""GPU parity tests: every block, through the C ABI (gr_clenabled_b200.blocks -> ctypes ->
libclenabled_b200.so), against the oracle on the same seeded inputs, against the golden
vectors, and -- at BASELINE.json sizes -- through size-independent properties.

Tolerances: bit-exact for clMathConst / clMathOp / the X-engine integer accumulators;
1e-5 relative (to the output's max magnitude) for FFT / filter / channelizer results,
the figure BASELINE.json's north_star states.
"""
import ctypes as C

import numpy as np
import pytest

from gr_clenabled_b200 import blocks, capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
c64 = np.complex64
GPU = (1, 1, 0, 0)        # openCLPlatformType=GPU, devSelector=first, platformId, devId
TOL = 1e-5


def rel_err(a, b):
    a = np.asarray(a).astype(np.complex128)
    b = np.asarray(b).astype(np.complex128)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def test_device_is_blackwell():
    assert capi.device_count() >= 1
    buf = C.create_string_buffer(256)
    capi.check(capi.load().clb200_device_name(0, buf, 256))
    assert b"sm_100" in buf.value, buf.value


# ------------------------------------------------------------------- clMathConst --
def test_mathconst_known_answer(golden):
    blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, float(golden["mc_k"]), capi.OP_MULTIPLY)
    assert np.array_equal(blk.work(golden["mc_in"]), golden["mc_out"])
    assert blk.counters()["launches"] >= 1


@pytest.mark.parametrize("op", [capi.OP_MULTIPLY, capi.OP_ADD, capi.OP_SUBTRACT, capi.OP_COMPLEX_CONJ,
                                capi.OP_EMPTY_W_COPY])
@pytest.mark.parametrize("n", [1, 3, 8192, 8192 * 3 + 5, 3_000_001])
def test_mathconst_complex_bit_exact(op, n):
    x = orc.rng_c32(n, orc.SEED_M)
    blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 0.7071, op)
    got = blk.work(x)
    assert np.array_equal(got.view(np.uint32), orc.mathconst(x, 0.7071, op).view(np.uint32))


def test_mathconst_float_int_and_setter():
    x = orc.rng_f32(10007, orc.SEED_M)
    blk = blocks.clMathConst(capi.DTYPE_FLOAT, *GPU, 3.0, capi.OP_ADD)
    assert np.array_equal(blk.work(x), orc.mathconst(x, 3.0, 2))
    blk.set_k(-1.5)
    assert blk.k() == -1.5
    assert np.array_equal(blk.work(x), orc.mathconst(x, -1.5, 2))
    xi = (orc.rng_f32(5001, 3) * 1e6).astype(np.int32)
    blk = blocks.clMathConst(capi.DTYPE_INT, *GPU, 7.9, capi.OP_MULTIPLY)
    assert np.array_equal(blk.work(xi), orc.mathconst(xi, 7.9, 1))
    assert blk.work(np.zeros(0, np.int32)).size == 0          # empty input


def test_mathconst_empty_op_leaves_output_untouched():
    blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 2.0, capi.OP_EMPTY)
    out = np.full(100, 9 + 9j, c64)
    blk.work(orc.rng_c32(100, 1), out=out)
    assert np.all(out == 9 + 9j)


# ---------------------------------------------------------------------- clMathOp --
@pytest.mark.parametrize("op", [capi.OP_MULTIPLY, capi.OP_ADD, capi.OP_SUBTRACT, capi.OP_MULTIPLY_CONJ])
@pytest.mark.parametrize("n", [1, 8192, 2_000_003])
def test_mathop_complex_bit_exact(op, n):
    a, b = orc.rng_c32(n, orc.SEED_M), orc.rng_c32(n, orc.SEED_M + 1)
    blk = blocks.clMathOp(capi.DTYPE_COMPLEX, *GPU, op)
    assert np.array_equal(blk.work(a, b).view(np.uint32), orc.mathop(a, b, op).view(np.uint32))


def test_mathop_float_and_int():
    a, b = orc.rng_f32(4099, 5), orc.rng_f32(4099, 6)
    for op in (1, 2, 3):
        assert np.array_equal(blocks.clMathOp(capi.DTYPE_FLOAT, *GPU, op).work(a, b), orc.mathop(a, b, op))
    ai, bi = (a * 1e5).astype(np.int32), (b * 1e5).astype(np.int32)
    for op in (1, 2, 3):
        assert np.array_equal(blocks.clMathOp(capi.DTYPE_INT, *GPU, op).work(ai, bi), orc.mathop(ai, bi, op))


def test_secondary_elementwise():
    x = orc.rng_c32(100003, orc.SEED_M + 2)
    f = np.abs(orc.rng_f32(100003, orc.SEED_M + 3)) + 0.01
    g = np.abs(orc.rng_f32(100003, orc.SEED_M + 4)) + 0.01
    assert np.allclose(blocks.clLog(*GPU, 10.0, 1.5).work(f), orc.log10(f, 10.0, 1.5), rtol=TOL, atol=TOL)
    assert np.allclose(blocks.clSNR(*GPU, 10.0, 0.0).work(f, g), orc.snr(f, g, 10.0, 0.0), rtol=TOL, atol=TOL)
    assert np.allclose(blocks.clComplexToMag(*GPU).work(x), orc.complex_to_mag(x), rtol=TOL)
    assert np.allclose(blocks.clComplexToArg(*GPU).work(x), orc.complex_to_arg(x), rtol=TOL, atol=TOL)
    m, p = blocks.clComplexToMagPhase(*GPU).work(x)
    assert np.allclose(m, orc.complex_to_mag(x), rtol=TOL) and np.allclose(p, orc.complex_to_arg(x), rtol=TOL, atol=TOL)
    assert np.allclose(blocks.clMagPhaseToComplex(*GPU).work(m, p), x, rtol=1e-4, atol=1e-6)


def test_empty_calls_and_bad_arguments_for_every_block():
    """zero items is a valid call for every block (the scheduler may offer none); constructor checks throw like the
    reference's (clFFT window length :74-76, channelizer buf_items :59-62, X-engine inputs :106-109)"""
    e = np.zeros(0, c64)
    ef = np.zeros(0, np.float32)
    assert blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 2.0, capi.OP_MULTIPLY).work(e).size == 0
    assert blocks.clMathOp(capi.DTYPE_COMPLEX, *GPU, capi.OP_ADD).work(e, e).size == 0
    assert blocks.clLog(*GPU, 10.0, 0.0).work(ef).size == 0
    assert blocks.clComplexToMag(*GPU).work(e).size == 0
    assert blocks.clFFT(1024, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(e).size == 0
    assert blocks.clFFT(65536, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(e).size == 0
    taps = np.ones(16, np.float32) / 16
    for use_time in (False, True):
        f = blocks.clFilter(*GPU, 2, taps, 1, 0, use_time)
        assert f.work(e).size == 0
        assert f.work(np.ones(1, c64)).size == 1 and f.work(np.ones(1, c64)).size == 0     # decimation phase carried
    assert blocks.clQuadratureDemod(1.0, *GPU).work(e).size == 0
    for bad in (lambda: blocks.clFFT(1, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU),
                lambda: blocks.clFFT((1 << 21) + 2, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU),  # non-power-of-two limit
                lambda: blocks.clFFT(1024, capi.FFT_FORWARD, np.ones(100, np.float32), capi.DTYPE_COMPLEX, *GPU),
                lambda: blocks.clFFT(1024, 0, [], capi.DTYPE_COMPLEX, *GPU),
                lambda: blocks.clFFT(1024, capi.FFT_BACKWARD, [], capi.DTYPE_FLOAT, *GPU),             # real input is forward-only
                lambda: blocks.clFFT(1 << 23, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU),
                lambda: blocks.clFilter(*GPU, 0, taps),
                lambda: blocks.clFilter(*GPU, 1, np.zeros(0, np.float32)),
                lambda: blocks.clPolyphaseChannelizer(*GPU, taps, 100, 8, 8, list(range(8))),
                lambda: blocks.clPolyphaseChannelizer(*GPU, taps, 64, 12, 12, list(range(12))),         # not a power of two
                lambda: _xe(capi.DTYPE_BYTE, 1, 1, 8, 8),
                lambda: _xe(capi.DTYPE_BYTE, 3, 4, 8, 8),
                lambda: _xe(capi.DTYPE_SHORT, 1, 4, 8, 8),
                lambda: _xe(capi.DTYPE_BYTE, 2, 33, 8, 8),                                            # > 64 rows
                lambda: blocks.clMathConst(99, *GPU, 1.0, capi.OP_MULTIPLY)):
        with pytest.raises(capi.Clb200Error) as ei:
            bad()
        assert ei.value.code == capi.EINVAL
    xe = _xe(capi.DTYPE_BYTE, 1, 4, 8, 8)
    with pytest.raises(capi.Clb200Error) as ei:                                                      # stream calls before stream_begin
        xe.push([np.zeros(16, np.int8)] * 4, 1)
    assert ei.value.code == capi.ESTATE


# ------------------------------------------------------------------------- clFFT --
@pytest.mark.parametrize("N", [2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_fft_forward_all_sizes(N):
    nvec = 37 if N <= 1024 else 5
    x = orc.rng_c32(N * nvec, orc.SEED_F)
    got = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(x)
    assert rel_err(got, orc.fft(x, N, -1)) < TOL
    want = np.fft.fft(x.astype(np.complex128).reshape(nvec, N), axis=1).reshape(-1)
    assert rel_err(got, want) < 2e-6


@pytest.mark.parametrize("N", [2, 4, 8, 64, 2048, 8192])
def test_fft_backward_window_shift(N, golden):
    x = orc.rng_c32(N * 4, orc.SEED_F + 1)
    w = orc.window_blackman(N)
    for direction in (capi.FFT_FORWARD, capi.FFT_BACKWARD):
        for win in (None, w):
            for shift in (False, True):
                blk = blocks.clFFT(N, direction, [] if win is None else win, capi.DTYPE_COMPLEX, *GPU, 0, 1, shift)
                assert rel_err(blk.work(x), orc.fft(x, N, direction, win, shift)) < TOL, (direction, win is None, shift)


@pytest.mark.parametrize("N", [2, 4, 8])
def test_fft_small_sizes_unaligned_device_pointers(N):
    """4- and 8-point transforms use 128-bit loads/stores when the pointers allow it; 8-byte-aligned device
    pointers must take the scalar path and give the same bits"""
    import torch
    nvec = 1000
    x = orc.rng_c32(N * nvec, orc.SEED_F + 9)
    sp = torch.cuda.current_stream().cuda_stream
    for shift in (False, True):
        blk = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU, 0, 1, shift)
        want = orc.fft(x, N, capi.FFT_FORWARD, None, shift)
        outs = []
        for off in (0, 1):                                   # float2 units
            d_in = torch.zeros(2 * (N * nvec + 2), dtype=torch.float32, device="cuda")
            d_out = torch.zeros_like(d_in)
            d_in[2 * off:2 * off + 2 * N * nvec] = torch.from_numpy(x.view(np.float32)).cuda()
            blk.launch_device(d_in.data_ptr() + 8 * off, d_out.data_ptr() + 8 * off, nvec, sp)
            torch.cuda.synchronize()
            outs.append(d_out[2 * off:2 * off + 2 * N * nvec].cpu().numpy().view(np.complex64))
        assert np.array_equal(outs[0], outs[1])
        assert rel_err(outs[0], want) < TOL


def test_fft_tone_known_answer(golden):
    for N in (2048, 8192):
        got = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(golden["tone_in_%d" % N])
        assert rel_err(got, golden["tone_fft_%d" % N]) < TOL
    # the harness configuration: tone * Blackman(8192) (test_clenabled.cc:812,835-851)
    w = golden["win_blackman_8192"]
    got = blocks.clFFT(8192, capi.FFT_FORWARD, w, capi.DTYPE_COMPLEX, *GPU).work(golden["tone_in_8192"])
    assert rel_err(got, orc.fft(golden["tone_in_8192"], 8192, -1, w)) < TOL


def test_fft_real_input_and_streams():
    N = 1024
    x = orc.rng_f32(N * 3, orc.SEED_F + 2)
    got = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_FLOAT, *GPU).work(x)
    assert rel_err(got, orc.fft_real(x, N)) < TOL
    xs = [orc.rng_c32(N * 2, orc.SEED_F + 10 + s) for s in range(3)]
    outs = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU, 0, 3).work_streams(xs)
    for xi, oi in zip(xs, outs):
        assert rel_err(oi, orc.fft(xi, N, -1)) < TOL


@pytest.mark.parametrize("N", [3, 5, 6, 7, 12, 100, 243, 1000, 1001, 1200, 3125, 4095, 10000, 20000])
def test_fft_sizes_that_are_not_a_power_of_two(N):
    """clFFT plans take any 2^a 3^b 5^c 7^d length (lib/clFFT_impl.cc:97-100); here every length runs as a chirp-z
    transform over the power-of-two kernels: all window / shift / direction / real-input combinations against the oracle
    (direct DFT in double) and pocketfft, plus the round trip"""
    nvec = 5 if N <= 4095 else 2
    x = orc.rng_c32(N * nvec, orc.SEED_F + 30)
    w = orc.window_blackman(N)
    ref = np.fft.fft(x.astype(np.complex128).reshape(nvec, N), axis=1).reshape(-1)
    got = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(x)
    assert rel_err(got, ref) < 3e-6
    assert rel_err(got, orc.fft(x, N, -1)) < TOL
    if N <= 4095:
        for direction in (capi.FFT_FORWARD, capi.FFT_BACKWARD):
            for win in (None, w):
                for shift in (False, True):
                    blk = blocks.clFFT(N, direction, [] if win is None else win, capi.DTYPE_COMPLEX, *GPU, 0, 1, shift)
                    assert rel_err(blk.work(x), orc.fft(x, N, direction, win, shift)) < TOL, (direction, win is None, shift)
        xr = orc.rng_f32(N * nvec, orc.SEED_F + 31)
        assert rel_err(blocks.clFFT(N, capi.FFT_FORWARD, w, capi.DTYPE_FLOAT, *GPU, 0, 1, True).work(xr), orc.fft_real(xr, N, w)) < TOL
    back = blocks.clFFT(N, capi.FFT_BACKWARD, [], capi.DTYPE_COMPLEX, *GPU).work(got) / N
    assert rel_err(back, x) < TOL


@pytest.mark.parametrize("N,nvec", [(65536, 90), (1000, 6500), (1 << 21, 3)])
def test_fft_multi_kernel_sizes_through_the_chunked_host_path(N, nvec):
    """the sizes that need scratch buffers (two-pass, chirp-z, five-pass), with enough vectors for several 32 MiB chunks
    on the three slot streams: every chunk owns its scratch, vectors sampled across all chunks against the oracle"""
    x = orc.rng_c32(N * nvec, orc.SEED_F + 50)
    blk = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU, 0, 1, True)
    got = blk.work(x)
    got2 = blk.work(x)                                   # scratch reused by the second call
    assert np.array_equal(got.view(np.uint32), got2.view(np.uint32))
    pick = sorted(set([0, nvec - 1] + list(range(0, nvec, max(1, nvec // 12)))))
    xs = np.concatenate([x[v * N:(v + 1) * N] for v in pick]).astype(np.complex128).reshape(len(pick), N)
    ref = np.fft.fftshift(np.fft.fft(xs, axis=1), axes=1) if N % 2 == 0 else None
    gs = np.concatenate([got[v * N:(v + 1) * N] for v in pick]).reshape(len(pick), N)
    assert rel_err(gs, ref) < 3e-6


@pytest.mark.parametrize("N", [32768, 65536, 1 << 18, 1 << 20])
def test_fft_sizes_above_16384_four_step(N):
    """sizes beyond one CTA's shared memory (clFFT plans accept them: waterfalls of 32768 / 65536 points): four-step
    decomposition, every window / shift / direction / real-input combination against the oracle and pocketfft"""
    nvec = 3 if N <= 65536 else 1
    x = orc.rng_c32(N * nvec, orc.SEED_F + 20)
    w = orc.window_blackman(N)
    ref = np.fft.fft(x.astype(np.complex128).reshape(nvec, N), axis=1).reshape(-1)
    got = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU).work(x)
    assert rel_err(got, ref) < 2e-6
    assert rel_err(got, orc.fft(x, N, -1)) < TOL
    if N <= 65536:
        for direction in (capi.FFT_FORWARD, capi.FFT_BACKWARD):
            for win in (None, w):
                for shift in (False, True):
                    blk = blocks.clFFT(N, direction, [] if win is None else win, capi.DTYPE_COMPLEX, *GPU, 0, 1, shift)
                    assert rel_err(blk.work(x), orc.fft(x, N, direction, win, shift)) < TOL, (direction, win is None, shift)
        xr = orc.rng_f32(N * nvec, orc.SEED_F + 21)
        assert rel_err(blocks.clFFT(N, capi.FFT_FORWARD, w, capi.DTYPE_FLOAT, *GPU, 0, 1, True).work(xr), orc.fft_real(xr, N, w)) < TOL
    back = blocks.clFFT(N, capi.FFT_BACKWARD, [], capi.DTYPE_COMPLEX, *GPU).work(got) / N
    assert rel_err(back, x) < TOL


@pytest.mark.parametrize("N", [2, 8, 64, 1024, 8192, 16384])
def test_fft_real_input_sizes_window_and_ignored_shift(N, golden):
    """dtype FLOAT: full Hermitian spectrum, window applied, and the shift flag ignored like the reference
    (lib/clFFT_impl.cc:594 shifts complex data only)"""
    nvec = 3
    x = orc.rng_f32(N * nvec, orc.SEED_F + 4)
    w = golden["win_blackman_8192"] if N == 8192 else orc.window_blackman(N)
    for win in (None, w):
        want = orc.fft_real(x, N, win)
        ref = np.fft.fft((x.reshape(nvec, N) * (1.0 if win is None else win)).astype(np.float64), axis=1).reshape(-1)
        assert rel_err(want, ref) < 2e-6
        for shift in (False, True):
            got = blocks.clFFT(N, capi.FFT_FORWARD, [] if win is None else win, capi.DTYPE_FLOAT, *GPU, 0, 1, shift).work(x)
            assert rel_err(got, want) < TOL, (win is None, shift)


def test_fft_full_size_properties():
    """BASELINE config 2 at scale: 4096 vectors of 8192 (256 MiB in) -- Parseval, linearity and
    forward->backward round trip, none of which needs the oracle to transform 32 Mi samples."""
    N, nvec = 8192, 4096
    x = orc.rng_c32(N * nvec, orc.SEED_F + 3)
    fwd = blocks.clFFT(N, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU)
    inv = blocks.clFFT(N, capi.FFT_BACKWARD, [], capi.DTYPE_COMPLEX, *GPU)
    X = fwd.work(x)
    e_in = np.sum(np.abs(x.reshape(nvec, N).astype(np.complex128)) ** 2, axis=1)
    e_out = np.sum(np.abs(X.reshape(nvec, N).astype(np.complex128)) ** 2, axis=1) / N
    assert np.max(np.abs(e_out / e_in - 1)) < 1e-5
    back = inv.work(X) / N
    assert rel_err(back, x) < TOL
    # spot-check 3 vectors against the oracle
    for v in (0, nvec // 2, nvec - 1):
        assert rel_err(X[v * N:(v + 1) * N], orc.fft(x[v * N:(v + 1) * N], N, -1)) < TOL


# ---------------------------------------------------------------------- clFilter --
def _stream_ref(x, taps, decim):
    """zero-state causal convolution of the whole stream, every decim-th sample (oracle FIR)"""
    hist = np.concatenate([np.zeros(taps.size - 1, c64), x])
    return orc.fir(hist, taps, decim)


@pytest.mark.parametrize("use_time", [True, False])
@pytest.mark.parametrize("decim", [1, 4])
def test_filter_lowpass_256_taps_streaming(golden, use_time, decim):
    taps = np.concatenate([golden["lp_30M_1M5_283k"], [0.0]]).astype(np.float32)      # BASELINE config 3
    x = orc.rng_c32(60000, orc.SEED_L)
    blk = blocks.clFilter(*GPU, decim, taps, 1, 0, use_time)
    # ragged calls: 8192-sample scheduler chunks, a multiple of nsamples, tiny and empty ones
    cuts = [0, 8192, 8192 + 257 * 9, 8192 + 257 * 9 + 1, 8192 + 257 * 9 + 1, 30001, 60000]
    got = np.concatenate([blk.work(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    want = _stream_ref(x, taps, decim)
    assert got.size == want.size
    assert rel_err(got, want) < TOL


def test_filter_matches_reference_overlap_add(golden):
    """same blocks the reference's fft_filter_ccf would produce (fftsize 512 / nsamples 257)"""
    taps = golden["ramp_taps_256"]
    x = orc.rng_c32(257 * 40, orc.SEED_L + 1)
    ola = orc.FftFilter(taps, 1).filter(x)
    for use_time in (False, True):
        got = blocks.clFilter(*GPU, 1, taps, 1, 0, use_time).work(x)
        assert rel_err(got, ola) < TOL
    assert blocks.filter_ref_sizes(256) == (512, 257)


@pytest.mark.parametrize("use_time", [False, True])
def test_filter_set_taps2_from_another_thread_while_work_runs(use_time):
    """set_taps2 is called from another thread than work() in a flowgraph (reference: d_setlock,
    lib/clFilter_impl.cc:774-789): every work() call must run with ONE consistent tap set, old or new"""
    import threading
    K, L = 200, 4096
    ta = (orc.rng_f32(K, 31) / K).astype(np.float32)
    tb = (orc.rng_f32(K, 32) / K + 0.01).astype(np.float32)
    sa, sb = float(np.sum(ta.astype(np.float64))), float(np.sum(tb.astype(np.float64)))
    blk = blocks.clFilter(*GPU, 1, ta, 1, 0, use_time)
    x = np.ones(L, c64)
    stop, errors, seen = threading.Event(), [], set()

    def setter():
        i = 0
        while not stop.is_set():
            blk.set_taps2(tb if i % 2 == 0 else ta)
            i += 1

    th = threading.Thread(target=setter)
    th.start()
    try:
        for _ in range(300):
            y = blk.work(x)
            tail = y[K:]                                       # past the history: DC gain of the active tap set
            if np.allclose(tail, sa, atol=1e-4):
                seen.add("a")
            elif np.allclose(tail, sb, atol=1e-4):
                seen.add("b")
            else:
                errors.append((float(tail.real.min()), float(tail.real.max())))
    finally:
        stop.set()
        th.join()
    assert not errors, errors[:3]
    assert seen == {"a", "b"}


@pytest.mark.parametrize("name", ["lp256_d1", "lp256_d4", "ramp256_d1", "short37_d3", "one_tap"])
def test_filter_matches_reference_build_vectors(golden, name):
    """the CUDA clFilter (both modes) against outputs of the reference's OWN fft_filter_ccf / fir_filter_ccf
    objects (tests/golden/ref_filters.npz, produced by oracle/_ref/libref_filters.so = lib/fft_filter.cc +
    lib/fir_filter.cc + lib/fft.cc compiled from the reference tree), fed in the same calls"""
    meta = golden[name + "_meta"]
    D, seed, n, fftsize, nsamples = (int(v) for v in meta[:5])
    calls = [int(v) for v in meta[5:]]
    taps = golden[name + "_taps"]
    assert blocks.filter_ref_sizes(taps.size) == (fftsize, nsamples)
    x = orc.rng_c32(n, seed)
    for use_time in (False, True):
        blk = blocks.clFilter(*GPU, D, taps, 1, 0, use_time)
        unit, pos, ys = nsamples * D, 0, []
        for c in calls:
            ys.append(blk.work(x[pos:pos + c * unit]))
            pos += c * unit
        got = np.concatenate(ys)
        assert got.size == golden[name + "_fft"].size
        assert rel_err(got, golden[name + "_fft"]) < TOL, use_time
        assert rel_err(got, golden[name + "_fir"]) < TOL, use_time


@pytest.mark.parametrize("ntaps", [1, 2, 31, 300, 1000, 3000])
def test_filter_tap_counts(ntaps):
    taps = (orc.rng_f32(ntaps, orc.SEED_L + 2) / ntaps).astype(np.float32)
    x = orc.rng_c32(20011, orc.SEED_L + 3)
    want = _stream_ref(x, taps, 1)
    for use_time in (False, True):
        blk = blocks.clFilter(*GPU, 1, taps, 1, 0, use_time)
        got = np.concatenate([blk.work(x[:7000]), blk.work(x[7000:])])
        assert rel_err(got, want) < TOL, (ntaps, use_time)


def test_filter_impulse_returns_taps_and_tap_swap(golden):
    taps = golden["hp_1M_100k_20k"]
    x = np.zeros(1000, c64)
    x[0] = 1.0
    blk = blocks.clFilter(*GPU, 1, taps)
    y = blk.work(x)
    assert np.allclose(y[:taps.size].real, taps, atol=1e-6) and np.allclose(y[taps.size:], 0, atol=1e-6)
    new = golden["rrc_1M_100k_035_111"]
    blk.set_taps2(new)                                 # history resets (clFilter_impl.cc:774-789)
    assert np.array_equal(blk.taps(), new)
    y = blk.work(x)
    assert np.allclose(y[:new.size].real, new, atol=1e-6)


def test_filter_full_size_linearity(golden):
    """BASELINE config 3 at scale (2^22 samples): linearity + agreement of the two kernels"""
    taps = np.concatenate([golden["lp_30M_1M5_283k"], [0.0]]).astype(np.float32)
    n = 1 << 22
    a, b = orc.rng_c32(n, orc.SEED_L + 4), orc.rng_c32(n, orc.SEED_L + 5)
    f = lambda v, t=False: blocks.clFilter(*GPU, 1, taps, 1, 0, t).work(v)
    ya, yb, yab = f(a), f(b), f((a + 2 * b).astype(c64))
    assert rel_err(yab, ya + 2 * yb) < TOL
    assert rel_err(f(a, True), ya) < TOL
    assert rel_err(ya[:5000], _stream_ref(a[:5000], taps, 1)) < TOL


# -------------------------------------------------------- clPolyphaseChannelizer --
@pytest.mark.parametrize("M,R,T,cmap", [
    (64, 64, 128, None), (64, 32, 128, None), (64, 64, 127, [5, 0, 63, 5]), (8, 8, 24, [3, 0, 5]),
    (2, 2, 7, None), (16, 4, 50, None), (256, 256, 1024, None), (1024, 512, 2048, [0, 1023, 7]),
    (64, 64, 256, None), (16, 16, 64, None), (64, 64, 200, [1, 2, 63]), (32, 32, 97, None),     # 3-4 taps per arm: the run kernel
])
def test_pfb_vs_oracle(M, R, T, cmap):
    taps = (orc.rng_f32(T, orc.SEED_P) * 0.1).astype(np.float32)
    cmap = list(range(M)) if cmap is None else cmap
    niter = 300 if M <= 64 else 21
    x = orc.rng_c32((niter - 1) * R + T, orc.SEED_P + 1)
    blk = blocks.clPolyphaseChannelizer(*GPU, taps, M * 4, M, R, cmap)
    got = blk.work(x, niter)
    assert rel_err(got, orc.pfb(x, taps, M, R, cmap, niter)) < TOL


@pytest.mark.parametrize("M,T", [(64, 128), (64, 256), (8, 16), (256, 512), (1024, 4096)])
def test_pfb_runs_of_time_steps_equal_single_steps(M, T, monkeypatch):
    """critically sampled, <= 4 taps per arm: a thread group walking consecutive time steps with the earlier samples in
    registers (k_pfb_run) gives the bits of the one-step-per-tile kernel, whichever of the two the block would pick"""
    taps = (orc.rng_f32(T, orc.SEED_P + 5) * 0.1).astype(np.float32)
    niter = 1000 if M <= 64 else 77
    x = orc.rng_c32((niter - 1) * M + T, orc.SEED_P + 6)
    outs = []
    for run in ("0", "1"):
        monkeypatch.setenv("CLB200_PFB_RUN", run)
        outs.append(blocks.clPolyphaseChannelizer(*GPU, taps, M * 4, M, M, list(range(M))).work(x, niter))
    assert np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    assert rel_err(outs[1], orc.pfb(x, taps, M, M, list(range(M)), niter)) < TOL


def test_pfb_baseline_config_tone_and_scale(golden):
    """BASELINE config 4: 64 channels, 128-tap prototype, buf_items 65536, tone on channel 5"""
    M = R = 64
    taps = np.concatenate([golden["lp_pfb64"], [0.0]]).astype(np.float32)
    niter = 65536 // R
    n = np.arange((niter - 1) * R + taps.size)
    x = (np.exp(2j * np.pi * 5 / M * n) + 1e-3 * orc.rng_c32(n.size, orc.SEED_P + 2)).astype(c64)
    blk = blocks.clPolyphaseChannelizer(*GPU, taps, 65536, M, R, list(range(M)))
    out = blk.work(x, niter)
    assert rel_err(out, orc.pfb(x, taps, M, R, list(range(M)), niter)) < TOL
    p = np.mean(np.abs(out.reshape(niter, M)[4:]) ** 2, axis=0)
    assert np.argmax(p) == 5 and p[5] > 0.99


# --------------------------------------------------------------------- clXEngine --
def _xe(dtype, npol, A, F, T):
    return blocks.clXEngine(*GPU, False, dtype, npol, A, 1, 0, F, T, [])


@pytest.mark.parametrize("A,F,T,npol", [
    (2, 1, 1, 1), (2, 3, 16, 1), (3, 17, 33, 1), (5, 4, 33, 2), (8, 16, 64, 1), (12, 256, 100, 2),
    (16, 33, 128, 2), (17, 20, 96, 1), (32, 64, 256, 1), (32, 9, 64, 2), (24, 7, 40, 2), (64, 5, 64, 1),
])
def test_xengine_ichar_bit_exact(A, F, T, npol):
    buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X)
    blk = _xe(capi.DTYPE_BYTE, npol, A, F, T)
    assert blk.output_items() == F * A * (A + 1) // 2 * npol * npol       # matrix_flat_length (:204-211)
    got = blk.work_i32(buf)
    assert np.array_equal(got, orc.xengine_exact(buf, A, F, T, npol))
    # float output: exact sum scaled by 1/127^2 vs the reference-order float emulation
    vis = blk.work(buf)
    assert rel_err(vis, orc.xengine_f32(buf, A, F, T, npol)) < TOL


@pytest.mark.parametrize("A,F,T,npol", [(32, 64, 256, 1), (8, 16, 64, 1), (16, 32, 128, 2)])
def test_xengine_ldg_feed_kernel_bit_exact(A, F, T, npol, monkeypatch):
    """The LDG-fed tcgen05 kernel stays the fallback for rows that are not 16 B aligned; force it."""
    monkeypatch.setenv("CLB200_XE_TMA", "0")
    buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X + 9)
    got = _xe(capi.DTYPE_BYTE, npol, A, F, T).work_i32(buf)
    assert np.array_equal(got, orc.xengine_exact(buf, A, F, T, npol))


@pytest.mark.parametrize("A,F,T,npol", [(32, 48, 1000, 1), (30, 40, 77, 1), (16, 24, 200, 2), (13, 8, 31, 2),
                                         (32, 256, 512, 1), (32, 64, 512, 1), (16, 16, 300, 2), (32, 128, 96, 1),
                                         (32, 2400, 64, 1), (5, 8, 2048, 1)])
def test_xengine_tma_feed_ragged_bit_exact(A, F, T, npol):
    """TMA-fed kernel (16 B aligned rows): channel counts that are not a multiple of the channel
    group, station counts below the box, time steps that do not fill the last stage (the out-of-range
    part of every box is zero-filled by the TMA unit); few channel groups -> clusters of 2, 4 and 8
    time-slice CTAs exchanging partial sums through distributed shared memory; many groups -> 16
    channels per CTA, several groups per CTA."""
    buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X + 10)
    got = _xe(capi.DTYPE_BYTE, npol, A, F, T).work_i32(buf)
    assert np.array_equal(got, orc.xengine_exact(buf, A, F, T, npol))


def test_xengine_extreme_values_do_not_overflow():
    A, F, T = 32, 4, 1024
    buf = np.full(T * A * F * 2, -128, np.int8)               # worst case |sum| = 2*128*128*1024 < 2^31
    got = _xe(capi.DTYPE_BYTE, 1, A, F, T).work_i32(buf)
    assert np.all(got[:, 0] == 2 * 128 * 128 * T) and np.all(got[:, 1] == 0)


def test_xengine_identical_and_rotated_stations():
    A, F, T = 8, 32, 512
    one = orc.rng_i8(T * F * 2, orc.SEED_X + 1).reshape(T, 1, F, 2)
    buf = np.repeat(one, A, axis=1).copy()
    buf[:, 1, :, 0], buf[:, 1, :, 1] = -one[:, 0, :, 1], one[:, 0, :, 0]     # station 1 = i * station 0
    got = _xe(capi.DTYPE_BYTE, 1, A, F, T).work_i32(buf).reshape(F, A * (A + 1) // 2, 2)
    auto = got[:, 0, :]
    assert np.all(auto[:, 1] == 0)
    assert np.all(got[:, 1, 0] == 0) and np.all(got[:, 1, 1] == auto[:, 0])    # (1,0): i*|x|^2 -> pure imaginary
    assert np.all(got[:, 5, :] == auto)                                           # (2,2) is an autocorrelation


def test_xengine_accumulate_and_complex_and_packed():
    A, F, T, npol = 6, 10, 64, 2
    buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X + 2)
    blk = _xe(capi.DTYPE_BYTE, npol, A, F, T)
    one = blk.work(buf)
    two = blk.work(buf, accumulate=True)                      # pipeline_integration (:785-808)
    assert rel_err(two, 2 * one) < 1e-6
    xc = orc.rng_c32(T * A * F * npol, orc.SEED_X + 3)
    got = _xe(capi.DTYPE_COMPLEX, npol, A, F, T).work(xc)
    assert rel_err(got, orc.xengine_f32(xc, A, F, T, npol)) < TOL
    packed = orc.rng_i8(T * A * F * npol, orc.SEED_X + 4).view(np.uint8)
    got = _xe(capi.DTYPE_PACKEDXY, npol, A, F, T).work_i32(packed)
    assert np.array_equal(got, orc.xengine_exact(orc.unpack4(packed), A, F, T, npol))


@pytest.mark.parametrize("A,F,T,npol", [(2, 1, 1, 1), (3, 17, 33, 1), (5, 4, 33, 2), (32, 64, 256, 1), (16, 33, 100, 2),
                                         (64, 5, 64, 1), (32, 7, 2048, 1), (9, 40, 50, 1), (32, 3, 16, 2), (32, 1024, 32, 1)])
def test_xengine_complex_float_tiled_kernel(A, F, T, npol):
    """DTYPE_COMPLEX (the reference's default): register-tiled FP32 kernel, time slices summed in a fixed order;
    against the oracle's reference-order float loop (lib/clXEngine_impl.cc:739-810), incl. accumulate and shards"""
    xc = orc.rng_c32(T * A * F * npol, orc.SEED_X + 70)
    blk = _xe(capi.DTYPE_COMPLEX, npol, A, F, T)
    want = orc.xengine_f32(xc, A, F, T, npol)
    got = blk.work(xc)
    assert rel_err(got, want) < TOL
    assert np.array_equal(blk.work(xc), got)                     # deterministic
    assert rel_err(blk.work(xc, accumulate=True, out=got.copy()), 2 * want) < TOL
    if F >= 4:
        sh = _xe(capi.DTYPE_COMPLEX, npol, A, F // 2, T)
        sh.set_shard(F, F - F // 2)
        assert rel_err(sh.work(xc), want.reshape(F, -1)[F - F // 2:].reshape(-1)) < TOL


@pytest.mark.parametrize("A,F,T,npol", [(16, 64, 128, 2), (32, 32, 100, 1), (12, 40, 77, 2), (16, 24, 64, 2), (16, 16, 512, 2),
                                         (16, 1024, 64, 2), (5, 48, 33, 1), (16, 2400, 32, 2)])
def test_xengine_packed_4bit_fused_unpack_bit_exact(A, F, T, npol, monkeypatch):
    """DTYPE_PACKEDXY with 16 B aligned packed rows: the TMA kernel expands the nibbles in its transpose stage (no
    separate unpack pass); every byte pattern, ragged shapes, 1 / 2 / 4 time slices; the separate-pass route
    (CLB200_XE_UNPACK_PASS=1) gives the same integers"""
    packed = orc.rng_i8(T * A * F * npol, orc.SEED_X + 80).view(np.uint8)
    packed[:256] = np.arange(256, dtype=np.uint8)                # every (re, im) nibble pair, incl. the -8 -> 0 entries
    want = orc.xengine_exact(orc.unpack4(packed), A, F, T, npol)
    blk = _xe(capi.DTYPE_PACKEDXY, npol, A, F, T)
    l0 = blk.counters()["launches"]
    got = blk.work_i32(packed)
    assert np.array_equal(got, want)
    fused_launches = blk.counters()["launches"] - l0
    monkeypatch.setenv("CLB200_XE_UNPACK_PASS", "1")
    blk2 = _xe(capi.DTYPE_PACKEDXY, npol, A, F, T)
    l0 = blk2.counters()["launches"]
    assert np.array_equal(blk2.work_i32(packed), want)
    assert blk2.counters()["launches"] - l0 > fused_launches      # the unpack kernel ran only on the second route
    vis = blk.work(packed)
    assert rel_err(vis, (want[:, 0] + 1j * want[:, 1]) / 49.0) < 1e-6


def test_xengine_and_channelizer_match_reference_kernel_outputs(golden):
    """the CUDA path against outputs of the reference's OWN OpenCL kernels (tests/golden/ref_kernels.npz: XCorrelate,
    CharToComplex incl. the packed-XY LUT, filterpfb2 + channel_map; source text emitted by the reference's builder
    functions and run on the CPU by oracle/ref_kernels.py)"""
    for A, F, T, npol, seed in [(5, 4, 33, 2, 6101), (3, 6, 16, 1, 6102), (32, 2, 64, 1, 6103), (16, 3, 40, 2, 6104)]:
        x = orc.rng_c32(T * A * F * npol, seed)
        got = _xe(capi.DTYPE_COMPLEX, npol, A, F, T).work(x)
        assert rel_err(got, golden["xc_c32_%d_%d_%d_%d_fma1" % (A, F, T, npol)]) < TOL
    for A, F, T, npol, seed in [(4, 8, 64, 1, 6201), (3, 4, 32, 2, 6202), (32, 4, 128, 1, 6203)]:
        b = orc.rng_i8(T * A * F * npol * 2, seed)
        assert rel_err(_xe(capi.DTYPE_BYTE, npol, A, F, T).work(b), golden["xc_i8_%d_%d_%d_%d" % (A, F, T, npol)]) < TOL
    for A, F, T, seed in [(4, 16, 32, 6301), (16, 16, 64, 6302)]:
        p = orc.rng_i8(T * A * F * 2, seed).view(np.uint8)
        assert rel_err(_xe(capi.DTYPE_PACKEDXY, 2, A, F, T).work(p), golden["xc_packed_%d_%d_%d" % (A, F, T)]) < TOL
    for M, R, ntaps, niter, cmap, seed in [(8, 8, 24, 19, None, 6401), (8, 4, 19, 21, None, 6402), (64, 64, 128, 16, None, 6403),
                                           (16, 16, 40, 9, [5, 0, 15, 3], 6404)]:
        taps = (orc.rng_f32(ntaps, seed) * 0.1).astype(np.float32)
        x = orc.rng_c32((niter - 1) * R + ntaps + (M - R), seed + 50)
        cm = list(range(M)) if cmap is None else cmap
        blk = blocks.clPolyphaseChannelizer(*GPU, taps, M * 4, M, R, cm)
        want = golden["pfb_%d_%d_%d_%d_%d" % (M, R, ntaps, niter, 0 if cmap is None else len(cmap))]
        assert rel_err(blk.work(x, niter), want) < TOL
    xm = orc.rng_c32(256, orc.SEED_M)
    for op in (1, 2, 3, 4):
        assert np.array_equal(blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 0.7071, op).work(xm), golden["mathconst_op%d" % op])
    a, b = orc.rng_c32(256, 6501), orc.rng_c32(256, 6502)
    for op in (1, 2, 3, 5):
        assert np.array_equal(blocks.clMathOp(capi.DTYPE_COMPLEX, *GPU, op).work(a, b), golden["mathop_op%d" % op])
    for K, seed in ((37, 6601), (256, 6602)):                 # td_FIR_complex: history = the first K-1 samples
        taps = (orc.rng_f32(K, seed) / K).astype(np.float32)
        x = orc.rng_c32(600 + K - 1, seed + 50)
        for use_time in (True, False):
            blk = blocks.clFilter(*GPU, 1, taps, 1, 0, use_time)
            y = blk.work(x)                                   # zero initial state: the first K-1 outputs are the warm-up
            assert rel_err(y[K - 1:], golden["tdfir_%d_fma1" % K]) < TOL


def test_xengine_channel_shard_matches_full():
    A, F, T = 8, 64, 128
    buf = orc.rng_i8(T * A * F * 2, orc.SEED_X + 5)
    full = _xe(capi.DTYPE_BYTE, 1, A, F, T).work_i32(buf).reshape(F, -1, 2)
    parts = []
    for g in range(4):
        blk = _xe(capi.DTYPE_BYTE, 1, A, F // 4, T)
        blk.set_shard(F, g * (F // 4))
        parts.append(blk.work_i32(buf).reshape(F // 4, -1, 2))
    assert np.array_equal(np.concatenate(parts, axis=0), full)


@pytest.mark.parametrize("A,F,T,npol,K", [(32, 64, 256, 1, 5), (32, 1024, 64, 1, 3), (16, 40, 100, 2, 4), (8, 2400, 32, 1, 2),
                                           (5, 7, 33, 1, 3)])
def test_xengine_batched_launch_equals_single_launches(A, F, T, npol, K):
    """clb200_xengine_launch_device_batch: K integrations back to back, one grid (4-D tensor map, the batch index is the
    outermost coordinate, so ragged T / stations / channels are zero-filled per integration); rows that are not 16 B
    aligned fall back to K launches -- either way every matrix equals the single launch and the oracle"""
    import torch
    per = T * A * F * npol * 2
    buf = orc.rng_i8(per * K, orc.SEED_X + 60)
    d_in = torch.from_numpy(buf).cuda()
    nout = F * (A * (A + 1) // 2) * npol * npol
    d_out = torch.zeros(K * nout * 2, dtype=torch.float32, device="cuda")
    d_one = torch.zeros(nout * 2, dtype=torch.float32, device="cuda")
    blk = _xe(capi.DTYPE_BYTE, npol, A, F, T)
    sp = torch.cuda.current_stream().cuda_stream
    blk.launch_device_batch(d_in.data_ptr(), d_out.data_ptr(), K, sp)
    torch.cuda.synchronize()
    got = d_out.cpu().numpy().view(np.complex64).reshape(K, nout)
    for k in range(K):
        blk.launch_device(d_in.data_ptr() + k * per, d_one.data_ptr(), False, sp)
        torch.cuda.synchronize()
        assert np.array_equal(got[k], d_one.cpu().numpy().view(np.complex64)), k
    want = orc.xengine_exact(buf[(K - 1) * per:], A, F, T, npol).astype(np.float64) / (127.0 * 127.0)
    assert np.max(np.abs(got[K - 1].real - want[:, 0])) < 1e-3 * max(1.0, np.max(np.abs(want)))
    assert np.max(np.abs(got[K - 1].imag - want[:, 1])) < 1e-3 * max(1.0, np.max(np.abs(want)))


def _ports_of(buf, T, A, F, npol, sb, planar):
    """the block's input streams for an integration buffer [t][station][chan][pol](sample): one array per station
    (per station and polarisation for unpacked two-polarisation data), each [t][chan]"""
    b = np.ascontiguousarray(buf).view(np.uint8).reshape(T, A, F, npol, sb)
    if planar:
        return [np.ascontiguousarray(b[:, s, :, p, :]).reshape(-1) for p in range(npol) for s in range(A)]
    return [np.ascontiguousarray(b[:, s]).reshape(-1) for s in range(A)]


@pytest.mark.parametrize("dtype,npol,A,F,T", [
    ("byte", 1, 8, 32, 96), ("byte", 2, 5, 16, 64), ("complex", 1, 4, 12, 40), ("complex", 2, 3, 8, 33),
    ("packed", 2, 6, 16, 64), ("byte", 1, 32, 64, 256),
])
def test_xengine_streaming_push_poll_matches_oracle(dtype, npol, A, F, T):
    """clb200_xengine_stream_begin / push_timesteps / poll_result (the general_work shape of the reference block,
    lib/clXEngine_impl.cc:918-1142, 1234-1299): ragged pushes that cross integration boundaries, results picked
    up later, every matrix equal to the whole-buffer call and to the oracle"""
    NI = 5
    dt = {"byte": capi.DTYPE_BYTE, "complex": capi.DTYPE_COMPLEX, "packed": capi.DTYPE_PACKEDXY}[dtype]
    sb = {"byte": 2, "complex": 8, "packed": 1}[dtype]
    planar = npol == 2 and dtype != "packed"
    per = T * A * F * npol
    if dtype == "byte":
        bufs = [orc.rng_i8(per * 2, orc.SEED_X + 20 + i) for i in range(NI)]
    elif dtype == "complex":
        bufs = [orc.rng_c32(per, orc.SEED_X + 20 + i) for i in range(NI)]
    else:
        bufs = [orc.rng_i8(per, orc.SEED_X + 20 + i).view(np.uint8) for i in range(NI)]
    whole = _xe(dt, npol, A, F, T)
    want = [whole.work(b) for b in bufs]
    if dtype == "byte":
        assert rel_err(want[0], orc.xengine_f32(bufs[0], A, F, T, npol)) < TOL
    # the stream: all integrations back to back, per port
    per_int = [_ports_of(b, T, A, F, npol, sb, planar) for b in bufs]
    ports = [np.concatenate([pi[k] for pi in per_int]) for k in range(len(per_int[0]))]
    item = F * sb * (1 if planar else npol)
    blk = _xe(dt, npol, A, F, T)
    blk.stream_begin(0, 8)
    got, pos, cuts = [], 0, [1, 7, T - 3, 2 * T + 5, 3, T, 9999]
    for c in cuts:
        n = min(c, NI * T - pos)
        if n <= 0:
            break
        blk.push([p[pos * item:] for p in ports], n)
        pos += n
        while True:
            m = blk.poll(wait=False)
            if m is None:
                break
            got.append(m)
    while len(got) < NI:
        m = blk.poll(wait=True)
        assert m is not None
        got.append(m)
    assert blk.poll(wait=True) is None
    st = blk.stream_state()
    assert st["integrations"] == NI and st["tracker"] == 0 and st["results_pending"] == 0
    for g, w in zip(got, want):
        assert np.array_equal(g, w) if dtype != "complex" else rel_err(g, w) < 1e-6
    blk.stream_end()


def test_xengine_streaming_pipeline_shard_pinned_ports_and_backpressure():
    A, F, T = 8, 64, 128
    NI = 4
    bufs = [orc.rng_i8(T * A * F * 2, orc.SEED_X + 40 + i) for i in range(NI)]
    whole = _xe(capi.DTYPE_BYTE, 1, A, F, T)
    want = [whole.work(b) for b in bufs]
    ports = [np.concatenate([_ports_of(b, T, A, F, 1, 2, False)[s] for b in bufs]) for s in range(A)]
    # pipeline_integration = 2: one matrix per two integrations, summed on the device (:785-808)
    blk = _xe(capi.DTYPE_BYTE, 1, A, F, T)
    blk.stream_begin(2, 2)
    blk.push(ports, NI * T)
    m0, m1 = blk.poll(wait=True), blk.poll(wait=True)
    assert rel_err(m0, want[0] + want[1]) < 1e-6 and rel_err(m1, want[2] + want[3]) < 1e-6
    # back-pressure: a full result ring refuses the push and changes nothing
    blk.push(ports, NI * T)
    with pytest.raises(capi.Clb200Error) as ei:
        blk.push(ports, T)
    assert ei.value.code == capi.ESTATE
    assert blk.stream_state()["results_pending"] == 2
    assert rel_err(blk.poll(wait=True), want[0] + want[1]) < 1e-6
    # channel shard of a wider stream (set_shard) + page-locked ports (DMA'd in place for pushes >= 1 MiB)
    lib = capi.load()
    A2, F2, T2 = 16, 256, 256
    big = orc.rng_i8(T2 * A2 * F2 * 2, orc.SEED_X + 50)
    full = _xe(capi.DTYPE_BYTE, 1, A2, F2, T2).work(big).reshape(F2, -1)
    pports = _ports_of(big, T2, A2, F2, 1, 2, False)
    for p in pports:
        capi.check(lib.clb200_register_host_buffer(C.c_void_p(p.ctypes.data), p.nbytes))
    try:
        for first in (0, 128):
            sh = _xe(capi.DTYPE_BYTE, 1, A2, 128, T2)
            sh.set_shard(F2, first)
            sh.stream_begin(0, 2)
            sh.push(pports, T2)
            assert np.array_equal(sh.poll(wait=True).reshape(128, -1), full[first:first + 128])
    finally:
        for p in pports:
            lib.clb200_unregister_host_buffer(C.c_void_p(p.ctypes.data))


def test_describe_and_set_debug(capfd):
    blk = blocks.clFFT(8192, capi.FFT_FORWARD, [], capi.DTYPE_COMPLEX, *GPU)
    buf = C.create_string_buffer(512)
    capi.check(capi.load().clb200_describe(blk._h, buf, 512))
    assert b"clFFT 8192-pt forward" in buf.value and b"k_fft<13," in buf.value
    capi.check(capi.load().clb200_set_debug(blk._h, 1))
    blk.work(orc.rng_c32(8192, 1))
    assert "clenabled_b200[debug]" in capfd.readouterr().err


def test_xengine_baseline_config_properties():
    """BASELINE config 5 (32 stations x 1024 channels, integration 1024, IChar): the oracle checks
    a 16-channel slab exactly; the whole result is checked through Hermitian/real-diagonal
    structure and additivity over time halves."""
    A, F, T = 32, 1024, 1024
    buf = orc.rng_i8(T * A * F * 2, orc.SEED_X + 6)
    blk = _xe(capi.DTYPE_BYTE, 1, A, F, T)
    got = blk.work_i32(buf).reshape(F, A * (A + 1) // 2, 2)
    b4 = buf.reshape(T, A, F, 2)
    slab = np.ascontiguousarray(b4[:, :, 500:516, :])
    assert np.array_equal(got[500:516].reshape(-1, 2), orc.xengine_exact(slab, A, 16, T, 1))
    diag = [s * (s + 1) // 2 + s for s in range(A)]
    assert np.all(got[:, diag, 1] == 0) and np.all(got[:, diag, 0] > 0)
    power = np.sum(b4.astype(np.int64) ** 2, axis=(0, 3))                 # [A][F]
    assert np.array_equal(got[:, diag, 0].astype(np.int64), power.T)
    half = _xe(capi.DTYPE_BYTE, 1, A, F, T // 2)
    lo = half.work_i32(np.ascontiguousarray(b4[:T // 2]))
    hi = half.work_i32(np.ascontiguousarray(b4[T // 2:]))
    assert np.array_equal((lo + hi).reshape(got.shape), got)


def test_xengine_fused_gather_writes_every_registered_matrix():
    """clb200_xengine_set_gather / launch_device_gather (the multi-GPU all-gather fused into the epilogue):
    on one GPU two 'ranks' own half of the channels each and store their slabs into BOTH full matrices
    (cudaMalloc'ed through the ABI, as the peers' matrices would be after clb200_ipc_open)."""
    import torch
    lib = capi.load()
    A, F, T = 32, 64, 256
    nbl = A * (A + 1) // 2
    buf = orc.rng_i8(T * A * F * 2, orc.SEED_X + 11)
    want = orc.xengine_exact(buf, A, F, T, 1).astype(np.float64) / (127.0 * 127.0)
    mats = []
    for _ in range(2):
        p = C.c_void_p()
        capi.check(lib.clb200_mem_alloc(0, F * nbl * 8, C.byref(p)))
        mats.append(p)
    # completion flags: one array per 'rank'; a launch releases its epoch into every array, gather_wait acquires them
    flags = []
    for _ in range(2):
        p = C.c_void_p()
        capi.check(lib.clb200_mem_alloc(0, 2 * 32 * 4, C.byref(p)))
        flags.append(p)
    b4 = buf.reshape(T, A, F, 2)
    sp = torch.cuda.current_stream().cuda_stream
    keep = []
    for r in range(2):
        blk = _xe(capi.DTYPE_BYTE, 1, A, F // 2, T)
        blk.set_shard(F, r * (F // 2))
        blk.set_gather([m.value for m in mats])
        blk.set_gather_sync(r, [f.value for f in flags])
        slab = torch.from_numpy(np.ascontiguousarray(b4[:, :, r * (F // 2):(r + 1) * (F // 2), :])).cuda()
        keep.append((blk, slab))
    side = torch.cuda.Stream()
    for epoch in (1, 2, 3):                                   # rank 1 launches on another stream; rank 0's stream waits on the flags
        with torch.cuda.stream(side):
            keep[1][0].launch_device_gather(keep[1][1].data_ptr(), side.cuda_stream)
            keep[1][0].gather_wait(side.cuda_stream)          # every rank signals (and waits) behind its launch
        keep[0][0].launch_device_gather(keep[0][1].data_ptr(), sp)
        keep[0][0].gather_wait(sp)
        torch.cuda.current_stream().synchronize()             # only rank 0's stream: the flags guarantee rank 1's slab
        fl = np.zeros(2 * 32, np.uint32)
        capi.check(lib.clb200_mem_copy_to_host(0, flags[0], fl.ctypes.data_as(C.c_void_p), fl.nbytes))
        assert fl[0] == epoch and fl[32] == epoch
        got = np.zeros(F * nbl, np.complex64)
        capi.check(lib.clb200_mem_copy_to_host(0, mats[0], got.ctypes.data_as(C.c_void_p), got.nbytes))
        assert np.max(np.abs(got.real - want[:, 0])) < 1e-3 and np.max(np.abs(got.imag - want[:, 1])) < 1e-3
    torch.cuda.synchronize()
    for f in flags:
        capi.check(lib.clb200_mem_free(0, f))
    for m in mats:
        got = np.zeros(F * nbl, np.complex64)
        capi.check(lib.clb200_mem_copy_to_host(0, m, got.ctypes.data_as(C.c_void_p), got.nbytes))
        assert np.max(np.abs(got.real - want[:, 0])) < 1e-3 and np.max(np.abs(got.imag - want[:, 1])) < 1e-3
        capi.check(lib.clb200_mem_free(0, m))
    h = C.create_string_buffer(64)
    p = C.c_void_p()
    capi.check(lib.clb200_mem_alloc(0, 4096, C.byref(p)))
    capi.check(lib.clb200_ipc_export(0, p, h))                 # a 64-byte handle another process could open
    assert any(h.raw)
    capi.check(lib.clb200_mem_free(0, p))
