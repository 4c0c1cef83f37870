import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _have_sm100():
    try:
        import torch
        return torch.cuda.is_available() and torch.cuda.get_device_capability(0)[0] == 10
    except Exception:  # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without an sm_100 device skips the gpu tests instead of failing in the library
    (which has no CPU fallback).  An explicit `-m gpu` run is left alone so that a missing device fails loudly there."""
    if "gpu" in (config.getoption("-m") or "") or _have_sm100():
        return
    skip = pytest.mark.skip(reason="needs an sm_100 (B200) device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def ala2():
    return np.load(os.path.join(GOLDEN_DIR, "ala2_xyz.npy"))


@pytest.fixture(scope="session")
def oracle_mod():
    """The CPU oracle (test infrastructure).  Builds liboracle.so on first use."""
    from oracle import oracle as O
    O.port_lib()
    return O


@pytest.fixture(scope="session")
def mdb():
    """The product package, with the CUDA library loaded (fails loudly if it was not built)."""
    import mdtraj_b200
    from mdtraj_b200 import _capi
    _capi.lib()
    return mdtraj_b200
