"""`import clenabled` -- the reference's python module name (python/__init__.py of gr-clenabled; what GRC's
generated flowgraphs import), served by the B200 build: every block class of gr_clenabled_b200.blocks under
the reference's names, so `clenabled.clFFT(...)`, `clenabled.clXEngine(...)` keep their spelling.
Outside a GNU Radio scheduler the classes take and return numpy arrays through `work()`; inside one, link the
C++ layer in gr_clenabled_b200/host instead."""
from gr_clenabled_b200.blocks import (  # noqa: F401
    clComplexFilter, clComplexToArg, clComplexToMag, clComplexToMagPhase, clFFT, clFilter, clLog,
    clMagPhaseToComplex, clMathConst, clMathOp, clPolyphaseChannelizer, clQuadratureDemod, clSignalSource,
    clSNR, clXCorrelate, clXEngine, clxcorrelate_fft_vcf)
