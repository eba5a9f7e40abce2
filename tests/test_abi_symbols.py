"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls),
and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from longtr_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(ROOT, "include", fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(ltr_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from longtr_b200 import build
    build.build()
    lib = C.CDLL(abi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 16
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_lists_the_same_symbols():
    assert set(abi.EXPORTED_SYMBOLS) <= set(declared_symbols())


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = abi.load()
    ctx = C.c_void_p()
    rc = lib.ltr_ctx_create(0, C.byref(ctx))
    assert rc == -1 and not ctx.value  # LTR_ERR_NO_DEVICE
    from longtr_b200 import Engine
    from longtr_b200.engine import LongTRError
    with pytest.raises(LongTRError):
        Engine(0)
    # the many-loci entry points need a device as well: no worker thread is started without one
    pipe = C.c_void_p()
    assert lib.ltr_pipeline_create(0, 64, 2, C.byref(pipe)) == -1 and not pipe.value
    from longtr_b200 import Pipeline
    with pytest.raises(LongTRError):
        Pipeline(0)
    assert lib.ltr_process_reads_flat_batch(None, 0, None, None, None) == -3  # LTR_ERR_INVALID without a context


def test_product_does_not_import_oracle():
    """Nothing under longtr_b200/ may reference oracle/ (test infrastructure)."""
    bad = []
    for d, _dirs, files in os.walk(os.path.join(ROOT, "longtr_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(d, fn), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|#include\s+\"[^\"]*oracle|ltr_oracle_|ltr_ref_", txt):
                    bad.append(os.path.join(d, fn))
    assert not bad, bad
