"""The pybind11 module `clenabled_python` (gr_clenabled_b200/host/python/bindings.cc): the reference's module name,
class names and constructor keyword names (python/bindings/*_python.cc, read into tests/golden/ref_constants.json
by tests/golden/make_constants.py); on the GPU the blocks are driven from Python through it."""
import glob
import importlib.util
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "ref_constants.json")) as f:
    REF = json.load(f)


def _module():
    paths = glob.glob(os.path.join(ROOT, "gr_clenabled_b200", "lib", REF["pybind_module"] + ".*.so"))
    if not paths:
        pytest.skip("clenabled_python not built (make -C gr_clenabled_b200/host)")
    spec = importlib.util.spec_from_file_location(REF["pybind_module"], paths[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_module_has_every_reference_class_with_the_reference_keywords():
    m = _module()
    for cls, args in REF["pybind_init_args"].items():
        assert hasattr(m, cls), cls
        doc = getattr(m, cls).__init__.__doc__
        sig = doc[doc.index("(") + 1:doc.index(") ->")]
        names = [a.split(":")[0].strip() for a in re.split(r",\s*(?![^\[]*\])", sig)][1:]       # drop self
        assert names == args, (cls, names, args)
    assert hasattr(m.clMathConst, "set_k") and hasattr(m.clFilter, "set_taps2") and hasattr(m.clComplexFilter, "taps")


@pytest.mark.gpu
def test_blocks_run_from_python_through_the_pybind_module():
    from oracle import oracle as orc
    m = _module()
    N, nvec = 1024, 6
    x = orc.rng_c32(N * nvec, 501)
    # the reference's (shifted) keyword names: openCLPlatformType is the LAST of the four device arguments
    fft = m.clFFT(fftSize=N, clFFTDir=-1, window=[], idataType=1, devSelector=1, platformId=1, devId=0, openCLPlatformType=0)
    assert fft.name() == "clFFT"
    y = np.zeros_like(x)
    assert fft.general_work(nvec, [x], [y]) == nvec
    want = orc.fft(x, N, -1)
    assert np.max(np.abs(y - want)) / np.max(np.abs(want)) < 1e-5
    mc = m.clMathConst(1, 1, 1, 0, 0, 2.0, 1, setDebug=1)
    assert mc.k() == 2.0 and any("k_map1" in ln or "device" in ln for ln in mc.log_lines())
    out = np.zeros(8192, np.complex64)
    xs = orc.rng_c32(8192, 502)
    assert mc.general_work(8192, [xs], [out]) == 8192
    assert np.array_equal(out, orc.mathconst(xs, 2.0, 1))
    mc.set_k(3.0)
    assert mc.k() == 3.0
    # X-engine: ports in, ("triang_matrix" . c32vector) PDU out after stop()
    A, F, T = 4, 16, 32
    buf = orc.rng_i8(T * A * F * 2, 503)
    ports = [np.ascontiguousarray(buf.reshape(T, A, F * 2)[:, s]).reshape(-1) for s in range(A)]
    xe = m.clXEngine(1, 1, 0, 0, False, 5, 1, A, 1, 0, F, T, [], pipeline_integration=0)
    assert xe.general_work(T, ports, []) == T
    xe.stop()
    msgs = xe.published("xcorr")
    assert len(msgs) == 1 and msgs[0][0] == "triang_matrix"
    want = orc.xengine_f32(buf, A, F, T, 1)
    assert np.max(np.abs(msgs[0][1] - want)) / np.max(np.abs(want)) < 1e-5
    with pytest.raises(Exception):
        m.clXEngine(1, 1, 0, 0, False, 5, 1, 1, 1, 0, F, T, [])          # fewer than 2 inputs (std::out_of_range)
