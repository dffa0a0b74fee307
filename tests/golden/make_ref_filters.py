"""Regenerates tests/golden/ref_filters.npz: outputs of the REFERENCE's own fft_filter_ccf / fir_filter_ccf
objects (lib/fft_filter.cc, lib/fir_filter.cc, lib/fft.cc compiled from /root/reference by oracle/Makefile into
oracle/_ref/libref_filters.so, against the stand-in VOLK/FFTW3/Boost headers in oracle/shim/).  Run where
/root/reference exists:

    make -C oracle && python tests/golden/make_ref_filters.py

Inputs are the oracle's counter-based generator (seed, length recorded per case), so only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (taps, decimation, calls (in units of nsamples*decimation blocks), seed)
def cases():
    lp = np.zeros(256, np.float32)
    lp[:255] = orc.firdes_low_pass_hamming(1.0, 30e6, 1.5e6, 283000.0)           # BASELINE config 3
    ramp = (np.arange(256) / 1000.0).astype(np.float32)                          # test-clfilter.cc:98-100
    short = orc.firdes_low_pass_hamming(1.0, 64.0, 0.5, 1.21)[:37].copy()
    return {
        "lp256_d1": (lp, 1, [3, 1, 4], 9001),
        "lp256_d4": (lp, 4, [4, 4], 9002),
        "ramp256_d1": (ramp, 1, [5], 9003),
        "short37_d3": (short, 3, [3, 6, 3], 9004),
        "one_tap": (np.array([0.5], np.float32), 1, [7], 9005),
    }


def main():
    assert orc.ref_filters() is not None, "oracle/_ref/libref_filters.so missing"
    out = {}
    for name, (taps, D, calls, seed) in cases().items():
        f = orc.RefFftFilter(taps, D)
        unit = f.nsamples * D
        n = unit * sum(calls)
        x = orc.rng_c32(n, seed)
        ys, pos = [], 0
        for c in calls:                                  # several calls: the tail / decimation phase carry over
            ys.append(f.filter(x[pos:pos + c * unit]))
            pos += c * unit
        y_fft = np.concatenate(ys)
        xh = np.concatenate([np.zeros(taps.size - 1, np.complex64), x])
        y_fir = orc.ref_fir(xh, taps, D)
        out[name + "_taps"] = taps
        out[name + "_meta"] = np.array([D, seed, n, f.fftsize, f.nsamples] + calls, np.int64)
        out[name + "_fft"] = y_fft
        out[name + "_fir"] = y_fir
        out[name + "_H"] = f.xformed_taps()
        print(name, "fftsize", f.fftsize, "nsamples", f.nsamples, "n", n, "fft-vs-fir",
              float(np.max(np.abs(y_fft - y_fir[:y_fft.size])) / np.max(np.abs(y_fir))))
    np.savez_compressed(os.path.join(HERE, "ref_filters.npz"), **out)


if __name__ == "__main__":
    main()
