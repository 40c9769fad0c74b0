import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "speech-enhancement-based-on-a-maximum-likelihood-criterion_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_pkg():
    """The package directory carries the reference's (hyphenated) name, so it is imported by path."""
    name = "se_ml_b200"
    if name in sys.modules:
        return sys.modules[name]
    try:                      # torch must come first: importing it AFTER libggd_b200.so has pulled in its own CUDA / NCCL libraries fails
        import torch  # noqa: F401
    except Exception:
        pass
    spec = importlib.util.spec_from_file_location(name, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def rel_err(a, b):
    """||a-b||_F / ||b||_F  -- the 'relative' of the north-star tolerances."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
