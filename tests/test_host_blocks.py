"""The C++ block layer (gr_clenabled_b200/host): gr::clenabled::* classes with the reference's
make() signatures over the C ABI, compiled against the GNU Radio stub."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "gr_clenabled_b200", "lib")


def test_block_library_is_built_and_links_the_c_abi():
    so = os.path.join(LIBDIR, "libgnuradio-clenabled-b200.so")
    assert os.path.exists(so), "run __graft_entry__.build()"
    C.CDLL(os.path.join(LIBDIR, "libclenabled_b200.so"), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(so)
    # the factories the reference exports (mangled gr::clenabled::<Block>::make)
    syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    for blk in ("clMathConst", "clMathOp", "clFFT", "clFilter", "clPolyphaseChannelizer", "clXEngine",
                "clLog", "clSNR", "clComplexToMag", "clComplexToArg", "clComplexToMagPhase", "clMagPhaseToComplex",
                "clXCorrelate", "clxcorrelate_fft_vcf", "clComplexFilter", "clQuadratureDemod", "clSignalSource"):
        assert ("9clenabled%d%s4make" % (len(blk), blk)) in syms, blk
    assert lib is not None


def test_driver_fails_loudly_without_a_gpu():
    from gr_clenabled_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([os.path.join(LIBDIR, "test_blocks")], capture_output=True, text=True)
    assert r.returncode != 0          # make() throws: there is no CPU path to fall back to


@pytest.mark.gpu
def test_block_classes_on_gpu():
    r = subprocess.run([os.path.join(LIBDIR, "test_blocks")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_block_layer_thread_safety_under_tsan():
    """host/test/test_threads.cc built with -fsanitize=thread (make -C gr_clenabled_b200/host tsan): setters from a second
    thread while work() runs, the X-engine push / pickup / stop path; no ThreadSanitizer report, no wrong output"""
    exe = os.path.join(LIBDIR, "test_threads_tsan")
    if not os.path.exists(exe):
        pytest.skip("TSAN driver not built")
    env = dict(os.environ, TSAN_OPTIONS="ignore_noninstrumented_modules=1 halt_on_error=0")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "thread test ok" in r.stdout, r.stdout + r.stderr[-2000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr
