"""The reference's own tests, unmodified, against a real mdtraj whose RMSD path has been swapped for the CUDA library by
mdtraj_b200.patch_mdtraj().  Needs baseline/_ref (baseline/build_ref.sh: the unmodified reference,
pip-installed offline; git-ignored, shipped to the GPU box by gpurun).  tests/reference_suite.py does the work in a
subprocess; every reference test that passes on the stock reference there must pass on the patched one."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HAVE_REF = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "mdtraj"))


def _run(*flags):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_suite.py"), *flags], capture_output=True,
                         text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not built (baseline/build_ref.sh)")
def test_stock_reference_passes_its_own_tests_on_the_stand_in_data():
    """CPU: the stand-in for md.load is a fair one -- the stock reference passes (almost) all of its tests on it."""
    got = _run()
    ok = [k for k, v in got["results"].items() if v == "ok"]
    assert len(ok) >= 24, got["results"]


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="baseline/_ref not built (baseline/build_ref.sh)")
def test_reference_tests_pass_through_patch_mdtraj():
    stock, patched = _run(), _run("--patched")
    assert patched["info"]["native_library"], "the CUDA library was never loaded: patch_mdtraj() did not take"
    bad = {k: v for k, v in patched["results"].items() if v != "ok" and stock["results"].get(k) == "ok"}
    assert not bad, bad
    assert sum(v == "ok" for v in patched["results"].values()) >= 24
