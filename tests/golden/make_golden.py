"""Regenerates tests/golden/*.npz.  Run in the build container (needs /root/reference):

    make -C oracle && python tests/golden/make_golden.py

ref_*.npz come from the REFERENCE's own code (lib/window.cc + lib/firdes.cc compiled
by oracle/Makefile into oracle/_ref/libref_firdes.so) -- they pin the oracle's tap and
window restatement and supply the taps the filter / channelizer parity tests use.
kat_*.npz are the reference tools' deterministic inputs with their analytic answers
(test_clenabled.cc:1193,1336,1351-1352; :835-851; test-clfilter.cc:98-100).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECT = 0, 1, 2, 3      # lib/window.h win_type


def main():
    R = orc.ref()
    assert R is not None, "oracle/_ref/libref_firdes.so missing: run make -C oracle where /root/reference exists"
    out = {}
    for name, wt, n in [("blackman_8192", WIN_BLACKMAN, 8192), ("blackman_2048", WIN_BLACKMAN, 2048),
                        ("hamming_255", WIN_HAMMING, 255), ("hamming_127", WIN_HAMMING, 127),
                        ("hann_64", WIN_HANN, 64)]:
        w = np.zeros(n, np.float32)
        assert R.ref_window_build(wt, n, 6.76, w) == n
        out["win_" + name] = w
    buf = np.zeros(65536, np.float32)
    # BASELINE config 3: low-pass, fs 30e6, fc 1.5e6; tw chosen so that ntaps = 255
    n = R.ref_firdes_low_pass(1.0, 30e6, 1.5e6, 283000.0, WIN_HAMMING, 6.76, buf, buf.size)
    out["lp_30M_1M5_283k"] = buf[:n].copy()
    # BASELINE config 4: 64-channel prototype, cutoff fs/128 -> 127 taps
    n = R.ref_firdes_low_pass(1.0, 64.0, 0.5, 1.21, WIN_HAMMING, 6.76, buf, buf.size)
    out["lp_pfb64"] = buf[:n].copy()
    n = R.ref_firdes_high_pass(1.0, 1e6, 1e5, 2e4, WIN_HAMMING, 6.76, buf, buf.size)
    out["hp_1M_100k_20k"] = buf[:n].copy()
    n = R.ref_firdes_band_pass(2.0, 1e6, 1e5, 2e5, 2e4, WIN_BLACKMAN, 6.76, buf, buf.size)
    out["bp_1M_100k_200k_20k_blackman"] = buf[:n].copy()
    n = R.ref_firdes_root_raised_cosine(1.0, 1e6, 1e5, 0.35, 111, buf, buf.size)
    out["rrc_1M_100k_035_111"] = buf[:n].copy()
    np.savez_compressed(os.path.join(HERE, "ref_firdes_window.npz"), **out)
    print("ref_firdes_window.npz:", {k: v.shape for k, v in out.items()})

    kat = {}
    # MultiplyConst: (1.0, 0.5) * 2 = (2.0, 1.0)
    kat["mc_in"] = np.full(8192, 1.0 + 0.5j, np.complex64)
    kat["mc_k"] = np.float32(2.0)
    kat["mc_out"] = np.full(8192, 2.0 + 1.0j, np.complex64)
    # tone: x[n] = i e^{-2 pi i n/N}  ->  forward DFT = i*N at bin N-1
    for N in (2048, 8192):
        x = orc.tone(N)
        X = np.zeros(N, np.complex128)
        X[N - 1] = 1j * N
        kat["tone_in_%d" % N] = x
        kat["tone_fft_%d" % N] = X.astype(np.complex64)
    kat["ramp_taps_256"] = (np.arange(256) / 1000.0).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **kat)
    print("kat.npz:", sorted(kat))


if __name__ == "__main__":
    main()
