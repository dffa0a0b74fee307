"""The persistent kernels take their tiles from a work counter (csrc/common.cuh: tile_fetch / tile_finish) once a
launch has more tiles than resident CTAs; CLB200_STATIC_TILES=1 (read per launch) falls back to static striding.
Both assignments must give the same bits, agree with the oracle, and leave the counter record at zero so that the
next launch on the same stream starts clean (every case below launches several times on the same handle)."""
import os

import numpy as np
import pytest

from gr_clenabled_b200 import blocks, capi
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
GPU = (1, 1, 0, 0)


def both_ways(fn):
    out = {}
    for static in ("1", "0", "0"):          # the second dynamic run reuses the counter record the first one left
        os.environ["CLB200_STATIC_TILES"] = static
        try:
            out.setdefault(static, []).append(fn())
        finally:
            os.environ.pop("CLB200_STATIC_TILES", None)
    assert np.array_equal(out["0"][0].view(np.uint32), out["0"][1].view(np.uint32))
    return out["1"][0], out["0"][0]


def rel_err(a, b):
    a = np.asarray(a).astype(np.complex128)
    b = np.asarray(b).astype(np.complex128)
    return float(np.max(np.abs(a - b)) / max(1e-30, np.max(np.abs(b))))


def test_mathconst_dynamic_tiles_bit_exact():
    n = 6_000_003                           # > 2 x 148 x 8 tiles of 16 KiB, ragged end
    x = orc.rng_c32(n, orc.SEED_M)
    blk = blocks.clMathConst(capi.DTYPE_COMPLEX, *GPU, 0.7071, capi.OP_MULTIPLY)
    st, dy = both_ways(lambda: blk.work(x))
    want = orc.mathconst(x, 0.7071, capi.OP_MULTIPLY)
    assert np.array_equal(st.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(dy.view(np.uint32), want.view(np.uint32))


def test_mathop_dynamic_tiles_bit_exact():
    n = 3_500_001
    a, b = orc.rng_c32(n, orc.SEED_M), orc.rng_c32(n, orc.SEED_M + 7)
    blk = blocks.clMathOp(capi.DTYPE_COMPLEX, *GPU, capi.OP_MULTIPLY)
    st, dy = both_ways(lambda: blk.work(a, b))
    want = orc.mathop(a, b, capi.OP_MULTIPLY)
    assert np.array_equal(st.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(dy.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("n,nvec", [(256, 12000), (512, 5000), (4096, 700), (8192, 400), (16384, 200)])
@pytest.mark.parametrize("direction", [capi.FFT_FORWARD, capi.FFT_BACKWARD])
def test_fft_dynamic_tiles(n, nvec, direction):
    x = orc.rng_c32(n * nvec, orc.SEED_F)
    blk = blocks.clFFT(n, direction, [], capi.DTYPE_COMPLEX, *GPU, 0, 1, True)
    st, dy = both_ways(lambda: blk.work(x))
    assert np.array_equal(st.view(np.uint32), dy.view(np.uint32))
    # oracle on a sample of the vectors (first, last, a stride through the middle)
    pick = sorted(set([0, 1, nvec - 1] + list(range(0, nvec, max(1, nvec // 16)))))
    xs = np.concatenate([x[v * n:(v + 1) * n] for v in pick])
    want = orc.fft(xs, n, direction, None, True)
    got = np.concatenate([dy[v * n:(v + 1) * n] for v in pick])
    assert rel_err(got, want) < 1e-5


@pytest.mark.parametrize("use_time", [False, True])
def test_filter_dynamic_tiles(use_time):
    taps = np.zeros(256, np.float32)
    taps[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)
    n = 1_800_000 if not use_time else 1_700_000      # > 148 x 12 blocks of 769 / > 148 x 5 tiles of 2048
    x = orc.rng_c32(n, orc.SEED_L)

    def run():
        blk = blocks.clFilter(*GPU, 1, taps, 1, 0, use_time)
        return blk.work(x)

    st, dy = both_ways(run)
    assert np.array_equal(st.view(np.uint32), dy.view(np.uint32))
    want = orc.fir(np.concatenate([np.zeros(255, np.complex64), x]), taps, 1)
    assert rel_err(dy, want) < 1e-5
