import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "image-feature-extraction_b200"),
          os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session (fails loudly without a GPU)."""
    try:                      # torch first: its bundled NCCL must be the one in the process when
        import torch          # a later test imports torch after the library has dlopen'ed NCCL
    except ImportError:
        pass
    import ife_b200
    c = ife_b200.Context(0)
    yield c
    c.close()
