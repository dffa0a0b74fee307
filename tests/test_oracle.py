"""CPU tests: pin the oracle (oracle/oracle_c.c).

The reference ships no golden outputs for the hot path, so the oracle
is pinned three ways:
  1. against the reference's OWN code where it compiles here: lib/window.cc and
     lib/firdes.cc via oracle/_ref (live when present, and always against the
     committed vectors tests/golden/ref_firdes_window.npz generated from it); and
     lib/fft_filter.cc + lib/fir_filter.cc + lib/fft.cc (fft_filter_ccf, fir_filter_ccf,
     fft_complex) compiled against the stand-in VOLK/FFTW3/Boost headers of oracle/shim/
     (oracle/_ref/libref_filters.so; vectors in tests/golden/ref_filters.npz);
  2. against the reference tools' known-answer inputs (tests/golden/kat.npz);
  3. against independent implementations of the published maths: numpy's pocketfft,
     scipy.signal.lfilter/upfirdn, exact int64 numpy einsum.
"""
import numpy as np
import pytest
from scipy import signal

from oracle import oracle as orc

c64 = np.complex64


def rel_err(a, b):
    return float(np.max(np.abs(a.astype(np.complex128) - b.astype(np.complex128))) /
                 max(1e-30, np.max(np.abs(b))))


# ---- 1. reference code: window / firdes ------------------------------------------
def test_oracle_window_matches_reference_vectors(golden):
    assert np.array_equal(orc.window_blackman(8192), golden["win_blackman_8192"])
    assert np.array_equal(orc.window_blackman(2048), golden["win_blackman_2048"])
    assert np.array_equal(orc.window_hamming(255), golden["win_hamming_255"])
    assert np.array_equal(orc.window_hamming(127), golden["win_hamming_127"])


def test_oracle_firdes_matches_reference_vectors(golden):
    t = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
    assert t.size == 255 and t.size % 2 == 1            # firdes.cc:680-683 forces odd lengths
    assert np.array_equal(t, golden["lp_30M_1M5_283k"])
    t = orc.firdes_low_pass_hamming(1.0, 64.0, 0.5, 1.21)
    assert np.array_equal(t, golden["lp_pfb64"])
    assert abs(float(np.sum(golden["lp_30M_1M5_283k"], dtype=np.float64)) - 1.0) < 1e-6   # unit DC gain


def test_golden_vectors_match_live_reference_build(golden):
    R = orc.ref()
    if R is None:
        pytest.skip("oracle/_ref not built on this box (vectors were generated where it was)")
    w = np.zeros(8192, np.float32)
    assert R.ref_window_build(2, 8192, 6.76, w) == 8192
    assert np.array_equal(w, golden["win_blackman_8192"])
    buf = np.zeros(4096, np.float32)
    n = R.ref_firdes_low_pass(1.0, 30e6, 1.5e6, 283000.0, 0, 6.76, buf, buf.size)
    assert np.array_equal(buf[:n], golden["lp_30M_1M5_283k"])


# ---- 2. known answers of the reference tools -------------------------------------
def test_kat_multiply_const(golden):
    out = orc.mathconst(golden["mc_in"], float(golden["mc_k"]), 1)
    assert np.array_equal(out, golden["mc_out"])          # test_clenabled.cc:1342-1356


@pytest.mark.parametrize("N", [2048, 8192])
def test_kat_tone_single_bin(golden, N):
    X = orc.fft(golden["tone_in_%d" % N], N)
    assert rel_err(X, golden["tone_fft_%d" % N]) < 1e-5    # energy only in bin N-1, value i*N
    assert abs(X[N - 1] - 1j * N) < 1e-3 * N


# ---- 3. independent implementations -----------------------------------------------
def test_rng_is_reproducible_and_in_range():
    a = orc.rng_f32(1000, 7)
    b = orc.rng_f32(500, 7, first=500)
    assert np.array_equal(a[500:], b)
    assert a.min() >= -1.0 and a.max() < 1.0
    i8 = orc.rng_i8(100000, orc.SEED_X)
    assert i8.min() == -127 and i8.max() == 127


@pytest.mark.parametrize("op", [1, 2, 3, 4, 254])
def test_mathconst_vs_numpy(op):
    x = orc.rng_c32(4099, orc.SEED_M)
    k = np.float32(0.7071)
    out = orc.mathconst(x, float(k), op)
    xf = x.view(np.float32)
    want = {1: xf * k, 2: xf + k, 3: xf - k, 4: np.conj(x).view(np.float32), 254: xf}[op]
    assert np.array_equal(out.view(np.float32), want.astype(np.float32))


@pytest.mark.parametrize("op", [1, 2, 3, 5])
def test_mathop_vs_numpy(op):
    a = orc.rng_c32(2050, orc.SEED_M)
    b = orc.rng_c32(2050, orc.SEED_M + 1)
    out = orc.mathop(a, b, op)
    ar, ai, br, bi = a.real, a.imag, b.real, b.imag
    if op == 5:
        bi = -bi
    if op in (1, 5):      # each product and sum rounded to float32 separately
        re = (ar * br).astype(np.float32) - (ai * bi).astype(np.float32)
        im = (ar * bi).astype(np.float32) + (ai * br).astype(np.float32)
    elif op == 2:
        re, im = ar + br, ai + b.imag
    else:
        re, im = ar - br, ai - b.imag
    assert np.array_equal(out.real, re.astype(np.float32))
    assert np.array_equal(out.imag, im.astype(np.float32))


def test_secondary_elementwise_vs_numpy():
    x = orc.rng_c32(3000, orc.SEED_M + 2)
    f = np.abs(orc.rng_f32(3000, orc.SEED_M + 3)) + 0.01
    g = np.abs(orc.rng_f32(3000, orc.SEED_M + 4)) + 0.01
    assert np.allclose(orc.log10(f, 10.0, 1.5), 10 * np.log10(f.astype(np.float64)) + 1.5, rtol=1e-6, atol=1e-6)
    assert np.allclose(orc.snr(f, g, 10.0, 0.0), np.abs(10 * np.log10((f / g).astype(np.float64))), rtol=1e-6, atol=1e-6)
    assert np.allclose(orc.complex_to_mag(x), np.abs(x.astype(np.complex128)), rtol=1e-6)
    assert np.allclose(orc.complex_to_arg(x), np.angle(x.astype(np.complex128)), rtol=1e-6, atol=1e-7)
    m, p = orc.complex_to_mag(x), orc.complex_to_arg(x)
    assert np.allclose(orc.magphase_to_complex(m, p), x, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("N", [2, 8, 64, 1024, 8192])
@pytest.mark.parametrize("direction", [-1, 1])
def test_fft_vs_pocketfft(N, direction):
    x = orc.rng_c32(N * 3, orc.SEED_F)
    got = orc.fft(x, N, direction)
    X = x.astype(np.complex128).reshape(3, N)
    want = np.fft.fft(X, axis=1) if direction < 0 else np.fft.ifft(X, axis=1) * N    # scale 1.0 both ways
    assert rel_err(got.reshape(3, N), want) < 2e-6


def test_fft_window_and_shift_rules():
    N = 256
    x = orc.rng_c32(N * 2, orc.SEED_F + 1)
    w = orc.window_blackman(N)
    X = x.astype(np.complex128).reshape(2, N)
    # forward: window then transform, then swap the output halves (clFFT_impl.cc:594-607)
    want = np.fft.fftshift(np.fft.fft(X * w, axis=1), axes=1)
    assert rel_err(orc.fft(x, N, -1, w, True).reshape(2, N), want) < 2e-6
    # backward + shift: input halves swapped on upload, THEN the window (:548-553, :566-580)
    want = np.fft.ifft(np.fft.fftshift(X, axes=1) * w, axis=1) * N
    assert rel_err(orc.fft(x, N, 1, w, True).reshape(2, N), want) < 2e-6


def test_fft_real_input_full_hermitian_spectrum():
    N = 512
    x = orc.rng_f32(N * 2, orc.SEED_F + 2)
    want = np.fft.fft(x.astype(np.float64).reshape(2, N), axis=1)
    assert rel_err(orc.fft_real(x, N).reshape(2, N), want) < 2e-6


@pytest.mark.parametrize("decim", [1, 4])
def test_fir_vs_lfilter(golden, decim):
    taps = golden["lp_30M_1M5_283k"]
    x = orc.rng_c32(5000, orc.SEED_L)
    hist = np.concatenate([np.zeros(taps.size - 1, c64), x])
    got = orc.fir(hist, taps, decim)
    want = signal.lfilter(taps.astype(np.float64), 1.0, x.astype(np.complex128))[::decim]
    assert got.size == want.size and rel_err(got, want) < 1e-5


@pytest.mark.parametrize("decim", [1, 2])
def test_fft_filter_overlap_add_vs_lfilter(golden, decim):
    taps = np.concatenate([golden["lp_30M_1M5_283k"], [0.0]]).astype(np.float32)     # 256 taps
    assert orc.fftfilt_sizes(256) == (512, 257)                                      # fft_filter.cc:77-78
    assert orc.fftfilt_sizes(300) == (1024, 725)
    f = orc.FftFilter(taps, decim)
    x = orc.rng_c32(257 * 2 * 8, orc.SEED_L)
    # two calls: the tail must carry across work() calls (d_tail, fft_filter.cc:172)
    got = np.concatenate([f.filter(x[:257 * 2 * 3]), f.filter(x[257 * 2 * 3:])])
    want = signal.lfilter(taps.astype(np.float64), 1.0, x.astype(np.complex128))
    # dec_ctr (:160-168) carries across the nsamples blocks of a call and both calls
    # are whole multiples of decim: the output is every decim-th sample of the stream
    assert got.size == x.size // decim
    assert rel_err(got, want[::decim]) < 1e-5


def test_fft_filter_ramp_taps(golden):
    taps = golden["ramp_taps_256"]                       # test-clfilter.cc:98-100
    f = orc.FftFilter(taps, 1)
    x = np.full(257 * 4, 1.0 + 0.5j, c64)                # test-clfilter.cc:77
    got = f.filter(x)
    want = signal.lfilter(taps.astype(np.float64), 1.0, x.astype(np.complex128))
    assert rel_err(got, want) < 1e-5
    assert abs(got[-1] - (1.0 + 0.5j) * taps.astype(np.float64).sum()) < 1e-3   # steady state = DC gain


def _pfb_direct(x, taps, M, R, ch_map, niter):
    T = taps.size
    out = np.zeros((niter, len(ch_map)), np.complex128)
    xd, td = x.astype(np.complex128), taps.astype(np.float64)
    n = np.arange(M)
    for i in range(niter):
        filt = np.zeros(M, np.complex128)
        for j in range(M):
            k = np.arange(j, T, M)
            filt[(j + i * (M - R)) % M] = np.sum(xd[i * R - k + T - 1] * td[k])
        spec = np.array([np.sum(filt * np.exp(2j * np.pi * n * c / M)) for c in range(M)])
        out[i] = spec[list(ch_map)]
    return out.reshape(-1)


REF_FILTER_CASES = ["lp256_d1", "lp256_d4", "ramp256_d1", "short37_d3", "one_tap"]


def _ref_case(golden, name):
    meta = golden[name + "_meta"]
    D, seed, n, fftsize, nsamples = (int(v) for v in meta[:5])
    calls = [int(v) for v in meta[5:]]
    return golden[name + "_taps"], D, seed, n, fftsize, nsamples, calls


@pytest.mark.parametrize("name", REF_FILTER_CASES)
def test_oracle_filters_match_reference_build_vectors(golden, name):
    """oracle_c.c's fft_filter_ccf / fir_filter_ccf restatement against outputs of the reference's OWN classes
    (tests/golden/ref_filters.npz, made by tests/golden/make_ref_filters.py from oracle/_ref/libref_filters.so):
    same sizes, same transformed taps, same outputs over several calls (tail and decimation phase carried)."""
    taps, D, seed, n, fftsize, nsamples, calls = _ref_case(golden, name)
    assert orc.fftfilt_sizes(taps.size) == (fftsize, nsamples)
    x = orc.rng_c32(n, seed)
    f = orc.FftFilter(taps, D)
    unit, pos, ys = nsamples * D, 0, []
    for c in calls:
        ys.append(f.filter(x[pos:pos + c * unit]))
        pos += c * unit
    got = np.concatenate(ys)
    want = golden[name + "_fft"]
    assert got.size == want.size
    assert rel_err(got, want) < 1e-6
    fir = orc.fir(np.concatenate([np.zeros(taps.size - 1, np.complex64), x]), taps, D)
    assert fir.size == golden[name + "_fir"].size
    assert rel_err(fir, golden[name + "_fir"]) < 1e-6
    # the transformed taps: H[k] = FFT(taps / fftsize) (fft_filter.cc:52-63)
    H = np.fft.fft(np.concatenate([taps.astype(np.float64) / fftsize, np.zeros(fftsize - taps.size)]))
    assert rel_err(golden[name + "_H"], H) < 1e-6


@pytest.mark.skipif(orc.ref_filters() is None, reason="oracle/_ref/libref_filters.so not built")
@pytest.mark.parametrize("ntaps,D", [(1, 1), (2, 2), (31, 1), (256, 1), (256, 3), (300, 2), (1000, 1)])
def test_oracle_filters_match_live_reference_build(ntaps, D):
    """the same comparison against the reference classes run live (random taps, more shapes, tap swap)"""
    taps = (orc.rng_f32(ntaps, 4242 + ntaps) / ntaps).astype(np.float32)
    rf, of = orc.RefFftFilter(taps, D), orc.FftFilter(taps, D)
    assert (rf.fftsize, rf.nsamples) == (of.fftsize, of.nsamples)
    unit = rf.nsamples * D
    x = orc.rng_c32(unit * 7, 4300 + ntaps)
    a = np.concatenate([rf.filter(x[:unit * 3]), rf.filter(x[unit * 3:])])
    b = np.concatenate([of.filter(x[:unit * 3]), of.filter(x[unit * 3:])])
    assert rel_err(b, a) < 1e-6
    xh = np.concatenate([np.zeros(ntaps - 1, np.complex64), x])
    assert rel_err(orc.fir(xh, taps, D), orc.ref_fir(xh, taps, D)) < 1e-6
    # both reference classes compute the same stream (zero-state convolution)
    assert rel_err(a, orc.ref_fir(xh, taps, D)[:a.size]) < 1e-5
    xv = orc.rng_c32(1024 * 3, 4400)
    for d in (-1, 1):
        assert rel_err(orc.fft(xv, 1024, d), orc.ref_fft(xv, 1024, d)) < 1e-6


@pytest.mark.skipif(orc.ref_filters() is None, reason="oracle/_ref/libref_filters.so not built")
def test_filter_fixture_is_what_the_live_reference_build_gives(golden):
    for name in REF_FILTER_CASES:
        taps, D, seed, n, fftsize, nsamples, calls = _ref_case(golden, name)
        f = orc.RefFftFilter(taps, D)
        x = orc.rng_c32(n, seed)
        unit, pos, ys = nsamples * D, 0, []
        for c in calls:
            ys.append(f.filter(x[pos:pos + c * unit]))
            pos += c * unit
        assert np.array_equal(np.concatenate(ys), golden[name + "_fft"])


# ---- 1b. reference code: the OpenCL kernels themselves ------------------------------
# tests/golden/ref_kernels.npz holds outputs of the reference's OWN kernels -- XCorrelate, CharToComplex (IChar and
# packed-XY LUT), filterpfb2 + channel_map, opconst_complex -- whose source text is emitted by the reference's builder
# functions and run on the CPU (oracle/ref_kernels.py, tests/golden/make_ref_kernels.py).
XC_COMPLEX = [(5, 4, 33, 2, 6101), (3, 6, 16, 1, 6102), (32, 2, 64, 1, 6103), (16, 3, 40, 2, 6104)]
XC_ICHAR = [(4, 8, 64, 1, 6201), (3, 4, 32, 2, 6202), (32, 4, 128, 1, 6203)]
XC_PACKED = [(4, 16, 32, 6301), (16, 16, 64, 6302)]
PFB_CASES = [(8, 8, 24, 19, None, 6401), (8, 4, 19, 21, None, 6402), (64, 64, 128, 16, None, 6403), (16, 16, 40, 9, [5, 0, 15, 3], 6404)]


def test_oracle_xengine_matches_reference_kernel_outputs(golden):
    for A, F, T, npol, seed in XC_COMPLEX:
        x = orc.rng_c32(T * A * F * npol, seed)
        got = orc.xengine_f32(x, A, F, T, npol)
        for fma in (0, 1):                                   # the reference builds either form, by device capability
            assert rel_err(got, golden["xc_c32_%d_%d_%d_%d_fma%d" % (A, F, T, npol, fma)]) < 1e-6
    for A, F, T, npol, seed in XC_ICHAR:
        b = orc.rng_i8(T * A * F * npol * 2, seed)
        want = golden["xc_i8_%d_%d_%d_%d" % (A, F, T, npol)]
        assert rel_err(orc.xengine_f32(b, A, F, T, npol), want) < 1e-6
        ex = orc.xengine_exact(b, A, F, T, npol).astype(np.float64) / (127.0 * 127.0)
        assert rel_err(ex[:, 0] + 1j * ex[:, 1], want) < 1e-5     # the exact integers, scaled like CharToComplex does
    lut = orc.unpack4(np.arange(256, dtype=np.uint8)).astype(np.float32).reshape(-1, 2)
    assert rel_err(lut[:, 0] + 1j * lut[:, 1], golden["packed_lut_all_bytes"] * 7.0) < 1e-6      # every byte through the LUT
    for A, F, T, seed in XC_PACKED:
        p = orc.rng_i8(T * A * F * 2, seed).view(np.uint8)
        ex = orc.xengine_exact(orc.unpack4(p), A, F, T, 2).astype(np.float64) / 49.0
        assert rel_err(ex[:, 0] + 1j * ex[:, 1], golden["xc_packed_%d_%d_%d" % (A, F, T)]) < 1e-5


def test_oracle_channelizer_and_mathconst_match_reference_kernel_outputs(golden):
    for M, R, ntaps, niter, cmap, seed in PFB_CASES:
        taps = (orc.rng_f32(ntaps, seed) * 0.1).astype(np.float32)
        x = orc.rng_c32((niter - 1) * R + ntaps + (M - R), seed + 50)
        cm = list(range(M)) if cmap is None else cmap
        want = golden["pfb_%d_%d_%d_%d_%d" % (M, R, ntaps, niter, 0 if cmap is None else len(cmap))]
        assert rel_err(orc.pfb(x, taps, M, R, cm, niter), want) < 1e-5
    xm = orc.rng_c32(256, orc.SEED_M)
    for op in (1, 2, 3, 4):
        assert np.array_equal(orc.mathconst(xm, 0.7071, op), golden["mathconst_op%d" % op])
    a, b = orc.rng_c32(256, 6501), orc.rng_c32(256, 6502)
    for op in (1, 2, 3, 5):                                   # two-input op_complex (lib/clMathOp_impl.cc:178-236)
        assert np.array_equal(orc.mathop(a, b, op), golden["mathop_op%d" % op])
    for K, seed in ((37, 6601), (256, 6602)):                 # td_FIR_complex (lib/clFilter_impl.cc:162-194)
        taps = (orc.rng_f32(K, seed) / K).astype(np.float32)
        x = orc.rng_c32(600 + K - 1, seed + 50)
        assert np.array_equal(orc.fir(x, taps, 1), golden["tdfir_%d_fma0" % K])
        assert rel_err(orc.fir(x, taps, 1), golden["tdfir_%d_fma1" % K]) < 1e-6
    # MATHOP_EMPTY_W_COPY: the reference's OpenCL string falls through into the multiply (missing break,
    # lib/clMathConst_impl.cc:187-193) while its CPU path copies; this repo copies (deviations)
    assert np.array_equal(golden["mathconst_op254"], golden["mathconst_op1"])
    assert np.array_equal(orc.mathconst(xm, 0.7071, 254), xm)


@pytest.mark.skipif(not __import__("oracle.ref_kernels", fromlist=["x"]).available(), reason="reference tree absent")
def test_reference_kernel_fixture_is_what_the_reference_sources_emit_now(golden):
    from oracle import ref_kernels as rk
    A, F, T, npol, seed = XC_COMPLEX[0]
    x = orc.rng_c32(T * A * F * npol, seed)
    assert np.array_equal(rk.xcorrelate(x, A, F, T, npol, True), golden["xc_c32_%d_%d_%d_%d_fma1" % (A, F, T, npol)])
    M, R, ntaps, niter, cmap, seed = PFB_CASES[1]
    taps = (orc.rng_f32(ntaps, seed) * 0.1).astype(np.float32)
    xp = orc.rng_c32((niter - 1) * R + ntaps + (M - R), seed + 50)
    assert np.array_equal(rk.pfb(xp, taps, M, R, list(range(M)), niter), golden["pfb_%d_%d_%d_%d_0" % (M, R, ntaps, niter)])
    src = rk.kernel_source("xcorr", 3, 4, 8, 2, 1, 0, 1)
    assert "__kernel void XCorrelate" in src and "#define d_num_baselines 6" in src


@pytest.mark.parametrize("M,R,T", [(8, 8, 24), (8, 4, 19), (64, 64, 128)])
def test_pfb_vs_direct_definition(M, R, T):
    taps = (orc.rng_f32(T, orc.SEED_P) * 0.1).astype(np.float32)
    niter = 6
    x = orc.rng_c32((niter - 1) * R + T, orc.SEED_P + 1)
    ch_map = list(range(M)) if M == 64 else [3, 0, 5]
    got = orc.pfb(x, taps, M, R, ch_map, niter)
    assert rel_err(got, _pfb_direct(x, taps, M, R, ch_map, niter)) < 2e-6


def test_pfb_tone_lands_in_one_channel(golden):
    M, R = 64, 64
    taps = np.concatenate([golden["lp_pfb64"], [0.0]]).astype(np.float32)     # 128-tap prototype
    niter, ch = 40, 5
    n = np.arange((niter - 1) * R + taps.size)
    x = np.exp(2j * np.pi * ch / M * n).astype(c64)
    out = orc.pfb(x, taps, M, R, list(range(M)), niter).reshape(niter, M)
    p = np.mean(np.abs(out[4:]) ** 2, axis=0)
    assert np.argmax(p) == ch
    far = np.delete(p, [ch - 1, ch, ch + 1])              # neighbours see the transition band
    assert p[ch] > 0.99 and far.max() < 1e-5 * p[ch]


def _xe_numpy(buf, A, F, T, npol):
    z = buf.reshape(T, A, F, npol, 2).astype(np.int64)
    re, im = z[..., 0], z[..., 1]
    out = []
    for f in range(F):
        for s1 in range(A):
            for s2 in range(s1 + 1):
                for p1 in range(npol):
                    for p2 in range(npol):
                        ar, ai, br, bi = re[:, s1, f, p1], im[:, s1, f, p1], re[:, s2, f, p2], im[:, s2, f, p2]
                        out.append((np.sum(ar * br + ai * bi), np.sum(ai * br - ar * bi)))
    return np.array(out, np.int64)


@pytest.mark.parametrize("A,F,T,npol", [(2, 3, 16, 1), (5, 4, 33, 2), (32, 2, 64, 1)])
def test_xengine_exact_vs_numpy(A, F, T, npol):
    buf = orc.rng_i8(T * A * F * npol * 2, orc.SEED_X)
    got = orc.xengine_exact(buf, A, F, T, npol)
    assert np.array_equal(got.astype(np.int64), _xe_numpy(buf, A, F, T, npol))


def test_xengine_identical_stations_is_autocorrelation():
    A, F, T = 4, 2, 128
    one = orc.rng_i8(T * F * 2, orc.SEED_X).reshape(T, 1, F, 2)
    buf = np.repeat(one, A, axis=1).copy()
    got = orc.xengine_exact(buf, A, F, T, 1).reshape(F, A * (A + 1) // 2, 2)
    assert np.all(got[:, :, 1] == 0)                      # imaginary part exactly 0
    assert np.all(got[:, :, 0] == got[:, :1, 0])          # every baseline = the autocorrelation


def test_xengine_float_emulation_close_to_exact():
    A, F, T = 6, 3, 256
    buf = orc.rng_i8(T * A * F * 2, orc.SEED_X + 1)
    exact = orc.xengine_exact(buf, A, F, T, 1).astype(np.float64) / (127.0 * 127.0)
    f32 = orc.xengine_f32(buf, A, F, T, 1)
    scale = np.max(np.abs(exact))
    assert np.max(np.abs(f32.real - exact[:, 0])) / scale < 1e-5
    assert np.max(np.abs(f32.imag - exact[:, 1])) / scale < 1e-5


def test_unpack4_lut():
    b = np.arange(256, dtype=np.uint8)
    out = orc.unpack4(b).reshape(256, 2)
    lut = np.array([0, 1, 2, 3, 4, 5, 6, 7, 0, -7, -6, -5, -4, -3, -2, -1])   # clXEngine_impl.cc:833
    assert np.array_equal(out[:, 0], lut[b >> 4]) and np.array_equal(out[:, 1], lut[b & 15])


@pytest.mark.parametrize("n", [3, 12, 100, 243, 1001])
def test_fft_oracle_any_length_matches_pocketfft(n):
    """lengths that are not a power of two take the direct DFT (double accumulation); the half swaps use
    vlen_2 = n / 2 (lib/clFFT_impl.cc:81): an odd length leaves its last element in place"""
    x = orc.rng_c32(n * 3, orc.SEED_F + 40)
    xx = x.reshape(3, n).astype(np.complex128)
    h = n // 2
    for direction in (-1, 1):
        for shift in (False, True):
            a = xx.copy()
            if direction > 0 and shift:
                a[:, :h], a[:, h:2 * h] = xx[:, h:2 * h], xx[:, :h]
            ref = np.fft.fft(a, axis=1) if direction < 0 else np.fft.ifft(a, axis=1) * n
            if direction < 0 and shift:
                r = ref.copy()
                ref[:, :h], ref[:, h:2 * h] = r[:, h:2 * h], r[:, :h]
            got = orc.fft(x, n, direction, None, shift).reshape(3, n)
            assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-6
    xr = orc.rng_f32(n * 2, orc.SEED_F + 41)
    ref = np.fft.fft(xr.reshape(2, n).astype(np.float64), axis=1)
    assert np.max(np.abs(orc.fft_real(xr, n).reshape(2, n) - ref)) / np.max(np.abs(ref)) < 1e-6
