"""pytest configuration: registers the `gpu` marker and puts the product source root
(`robust-multimodal-pd_b200/`) and the repo root (for `oracle/`) on sys.path."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "robust-multimodal-pd_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def _load(name):
        return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    return _load
