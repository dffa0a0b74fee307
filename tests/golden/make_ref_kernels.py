"""Regenerates tests/golden/ref_kernels.npz: outputs of the REFERENCE's own OpenCL kernels (XCorrelate,
CharToComplex, filterpfb2 + channel_map, opconst_complex), whose source text is emitted by the reference's own
builder functions and run on the CPU by oracle/ref_kernels.py.  Run where /root/reference exists:

    python tests/golden/make_ref_kernels.py

Inputs come from the oracle's counter-based generator (seeds recorded), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from oracle import ref_kernels as rk  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
XC_COMPLEX = [(5, 4, 33, 2, 6101), (3, 6, 16, 1, 6102), (32, 2, 64, 1, 6103), (16, 3, 40, 2, 6104)]
XC_ICHAR = [(4, 8, 64, 1, 6201), (3, 4, 32, 2, 6202), (32, 4, 128, 1, 6203)]
XC_PACKED = [(4, 16, 32, 6301), (16, 16, 64, 6302)]
PFB = [(8, 8, 24, 19, None, 6401), (8, 4, 19, 21, None, 6402), (64, 64, 128, 16, None, 6403), (16, 16, 40, 9, [5, 0, 15, 3], 6404)]


def main():
    assert rk.available(), "needs /root/reference"
    out = {}
    for A, F, T, npol, seed in XC_COMPLEX:
        x = orc.rng_c32(T * A * F * npol, seed)
        for fma in (0, 1):
            out["xc_c32_%d_%d_%d_%d_fma%d" % (A, F, T, npol, fma)] = rk.xcorrelate(x, A, F, T, npol, bool(fma))
    for A, F, T, npol, seed in XC_ICHAR:
        b = orc.rng_i8(T * A * F * npol * 2, seed)
        xc = rk.char_to_complex(b, A, F, T, npol)
        out["xc_i8_%d_%d_%d_%d" % (A, F, T, npol)] = rk.xcorrelate(xc, A, F, T, npol, True)
    every = np.arange(256, dtype=np.uint8)
    lut = rk.packed_lut_to_complex(np.repeat(every, 1)[:256].reshape(-1), 1, 128, 1)     # 128 channels x (X, Y) = 256 bytes
    out["packed_lut_all_bytes"] = lut                      # [chan][pol] complex: (LUT[hi], LUT[lo]) / 7
    for A, F, T, seed in XC_PACKED:
        p = orc.rng_i8(T * A * F * 2, seed).view(np.uint8)
        xc = rk.packed_lut_to_complex(p, A, F, T)
        out["xc_packed_%d_%d_%d" % (A, F, T)] = rk.xcorrelate(xc, A, F, T, 2, True)
    for M, R, ntaps, niter, cmap, seed in PFB:
        taps = (orc.rng_f32(ntaps, seed) * 0.1).astype(np.float32)
        x = orc.rng_c32((niter - 1) * R + ntaps + (M - R), seed + 50)      # the reference kernel may read M-R past (R < M)
        out["pfb_%d_%d_%d_%d_%d" % (M, R, ntaps, niter, 0 if cmap is None else len(cmap))] = rk.pfb(
            x, taps, M, R, list(range(M)) if cmap is None else cmap, niter)
    xm = orc.rng_c32(256, orc.SEED_M)
    for op in (1, 2, 3, 4, 254):
        out["mathconst_op%d" % op] = rk.mathconst(xm, 0.7071, op)
    a, b = orc.rng_c32(256, 6501), orc.rng_c32(256, 6502)
    for op in (1, 2, 3, 5):
        out["mathop_op%d" % op] = rk.mathop(a, b, op)
    for K, seed in ((37, 6601), (256, 6602)):
        taps = (orc.rng_f32(K, seed) / K).astype(np.float32)
        x = orc.rng_c32(600 + K - 1, seed + 50)               # K-1 history samples + 600 new ones
        for fma in (0, 1):
            out["tdfir_%d_fma%d" % (K, fma)] = rk.td_fir(x, taps, bool(fma))
    np.savez_compressed(os.path.join(HERE, "ref_kernels.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
