"""longtr_b200 -- B200-native (sm_100a) implementation of LongTR's read x haplotype hot path.

The package holds only what that path needs: ``csrc/`` (CUDA kernels, the C ABI and the
C++ host mirror of LongTR's HapAligner / Genotyper interfaces) and thin ctypes bindings.
"""
from .abi import DEFAULT_ALN_PARAMS, EXPORTED_SYMBOLS, LIB_PATH  # noqa: F401
from .engine import Engine, Job, LongTRError, Pipeline  # noqa: F401
from .locus_batch import Genotyper, build_locus_batch  # noqa: F401,E402
