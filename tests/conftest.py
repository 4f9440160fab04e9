"""Shared fixtures.  `gpu`-marked tests need a B200; everything else runs on CPU."""
import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 (B200) device")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure only)."""
    return _load("dinosoft_oracle", os.path.join(ROOT, "oracle", "dinosoft_oracle.py"))


@pytest.fixture(scope="session")
def pkg():
    """The product package (its directory name is not a valid identifier, so go through the shim)."""
    import dinosoft_b200

    return dinosoft_b200


def golden_files():
    """Fixtures of the loss path (w<world>_*.npz); other fixtures (pair_stats_*) have their own tests."""
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("w"))
