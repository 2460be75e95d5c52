"""pytest configuration: the `gpu` marker, and handles to the oracle and to the product package."""
import importlib
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure only)."""
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden():
    return json.loads((ROOT / "tests" / "golden" / "golden.json").read_text())


@pytest.fixture(scope="session")
def pkg():
    """The product package; its shared library must exist (no fallback)."""
    return importlib.import_module("openmp-parallel-computing_b200")


@pytest.fixture(scope="session")
def gpu(pkg):
    """Product package bound to cuda:0; fails loudly when the library or the device is missing."""
    pkg.init(0)
    yield pkg
    pkg.sync()
