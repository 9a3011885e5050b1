import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def built_lib():
    from mpcith_kyber_kosk_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def ctxs(built_lib):
    """Context factory: one context per (KYBER_K, chunk) on cuda:0."""
    if not _has_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    from mpcith_kyber_kosk_b200 import KoskContext
    made = {}

    def get(k, chunk=64, lanes=0, tensor=False):
        if (k, chunk, lanes, tensor) not in made:
            made[(k, chunk, lanes, tensor)] = KoskContext(k, 0, chunk, lanes, tensor)
        return made[(k, chunk, lanes, tensor)]
    yield get
    for c in made.values():
        c.close()
