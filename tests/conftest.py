import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine():
    from longtr_b200 import Engine
    from longtr_b200.engine import LongTRError
    try:
        eng = Engine(0)
    except LongTRError as e:
        # a machine without any NVIDIA device: the gpu-marked tests are skipped; on a GPU box a context that cannot be
        # created is a failure (there is no CPU fallback to fall back to)
        if os.path.exists("/dev/nvidiactl"):
            raise
        pytest.skip("no CUDA device on this machine: %s" % e)
    yield eng
    eng.close()
