import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # build what is missing (nvcc cross-compiles without a GPU; gcc for the oracle)
    from clane_b200 import build as _build
    from oracle import oracle as _oracle
    if not _build.LIB.exists():
        _build.build()
    _oracle.build()


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"


@pytest.fixture(scope="session")
def data_root():
    return ROOT / "tests" / "data_root"
