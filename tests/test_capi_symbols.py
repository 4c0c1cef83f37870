"""CPU: the C-ABI library loads and exports every symbol include/b200rmsd.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200rmsd.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rmsd_\w+)\s*\(", txt)))


def test_header_symbols_exported():
    from mdtraj_b200 import _capi
    assert os.path.exists(_capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200rmsd.h but not exported"
    # the ctypes table covers the header exactly
    assert sorted(_capi.SIGNATURES) == syms


def test_abi_version_and_error_string():
    from mdtraj_b200 import _capi
    L = _capi.lib()
    assert L.b200rmsd_abi_version() == 2
    assert isinstance(L.b200rmsd_last_error(), bytes)
    assert L.b200rmsd_scratch_bytes(1000, 25000) > 1000 * 7 * 64
    assert L.b200rmsd_allpairs_workspace_bytes(100, 22) >= 100 * 3 * 32 * 4


def test_no_cpu_fallback_without_device():
    """Without a GPU the host entry points must fail loudly (ENODEVICE / ECUDA), never compute on the CPU."""
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import mdtraj_b200 as mdb
    from mdtraj_b200 import _capi
    t = mdb.Trajectory(np.zeros((2, 4, 3), np.float32))
    with pytest.raises(_capi.B200RMSDError):
        mdb.rmsd(t, t, 0)
    with pytest.raises(_capi.B200RMSDError):
        t.superpose(t, 0)
    with pytest.raises(RuntimeError):
        mdb.DeviceTrajectory.from_host(t.xyz)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch it --
    not the package, and not the development tools either."""
    for sub in ("mdtraj_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
