"""Run the REFERENCE'S OWN test functions (unmodified files of the installed reference, baseline/_ref/tests/*.py) against a
real mdtraj, optionally after mdtraj_b200.patch_mdtraj() has swapped the CUDA implementation in.

    python tests/reference_suite.py [--patched]     -> one JSON object {test id: "ok" | "error text"}

Not a test module itself (tests/test_reference_integration.py drives it in a subprocess so that the real mdtraj and its
top-level `tests` package never leak into the main pytest process).  The reference's data tarball (tests/data.tar.gz) is
not in /root/reference, and PyTables is not installed, so `md.load` is replaced by a stand-in that returns seeded
MD-like trajectories with a real mdtraj.Topology (22 atoms, 11 residues) for the file names the tests ask for; the test
bodies -- /root/reference/tests/test_rmsd.py:36-334, test_rmsd_memmap.py:11-56, test_alignment.py:39-65,
test_trajectory.py:316-347, test_lprmsd.py:25-79,110-130 -- run as they are.  The reference's conftest.py is not loaded (its pytest_configure untars the
missing data); the two fixtures the tests use, get_fn and the parametrised flags, are passed as plain arguments.
"""
import importlib.util
import json
import os
import sys
import traceback

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "baseline", "_ref")

CASES = [  # (file, function, kwargs)
    ("test_rmsd.py", "test_trajectory_rmsd", {"parallel": True, "superpose": True}),
    ("test_rmsd.py", "test_trajectory_rmsd", {"parallel": False, "superpose": True}),
    ("test_rmsd.py", "test_trajectory_rmsd", {"parallel": True, "superpose": False}),
    ("test_rmsd.py", "test_trajectory_rmsd", {"parallel": False, "superpose": False}),
    ("test_rmsd.py", "test_precentered_1", {}),
    ("test_rmsd.py", "test_precentered_2", {}),
    ("test_rmsd.py", "test_superpose_0", {}),
    ("test_rmsd.py", "test_superpose_1", None),
    ("test_rmsd.py", "test_superpose_2", None),
    ("test_rmsd.py", "test_superpose_refinds", None),
    ("test_rmsd.py", "test_rmsd_atom_indices", {}),
    ("test_rmsd.py", "test_rmsd_ref_ainds_superpose", {"superpose": True}),
    ("test_rmsd.py", "test_rmsd_ref_ainds_superpose", {"superpose": False}),
    ("test_rmsd.py", "test_trajectory_rmsf", {}),
    ("test_rmsd.py", "test_trajectory_rmsf_aligned", {}),
    ("test_rmsd.py", "test_trajectory_rmsf_by_residue", {"parallel": True}),
    ("test_rmsd.py", "test_trajectory_rmsf_by_residue", {"parallel": False}),
    ("test_rmsd.py", "test_rmsd_atom_indices_vs_ref_indices", None),
    ("test_rmsd.py", "test_superpose_with_empty_atom_raises_exception", None),
    ("test_rmsd_memmap.py", "test_1", {}),
    ("test_rmsd_memmap.py", "test_2", {}),
    ("test_alignment.py", "test_rmsd_zero", None),
    ("test_alignment.py", "test_rmsd_nonzero", None),
    ("test_alignment.py", "test_transform", None),
    ("test_alignment.py", "test_transform2", None),
    ("test_lprmsd.py", "test_lprmsd_null", None),
    ("test_lprmsd.py", "test_lprmsd_0", None),
    ("test_lprmsd.py", "test_lprmsd_1", None),
    ("test_lprmsd.py", "test_lprmsd_2", None),
    ("test_lprmsd.py", "test_lprmsd_4", {}),
    ("test_lprmsd.py", "test_lprmsd_5", {}),
    ("test_trajectory.py", "test_center", {}),
    ("test_trajectory.py", "test_center_aind", {}),
]


def make_loader(md):
    """Stand-in for md.load: the same seeded trajectory for every .h5/.dcd name, its base structure for .pdb names."""
    n_res, names = 11, (("N", md.element.nitrogen), ("CA", md.element.carbon))
    top = md.Topology()
    chain = top.add_chain()
    for r in range(n_res):
        res = top.add_residue("ALA", chain)
        for nm, el in names:
            top.add_atom(nm, el, res)
    n_atoms = n_res * len(names)
    rng = np.random.default_rng(1234)
    base = rng.standard_normal((n_atoms, 3)) * 0.35
    q = rng.standard_normal((100, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    R = np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], -1),
                  np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], -1),
                  np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], -1)], 1)
    frames = np.einsum("fni,fij->fnj", base[None] + rng.standard_normal((100, n_atoms, 3)) * 0.05, R) + \
        rng.uniform(-2, 2, size=(100, 1, 3))

    def load(fn, stride=None, frame=None, top=None, **_kw):
        xyz = base[None] if str(fn).endswith(".pdb") else frames
        if frame is not None:
            xyz = xyz[frame:frame + 1]
        if stride:
            xyz = xyz[::stride]
        topo = make_loader.top.copy()
        return md.Trajectory(np.array(xyz, dtype=np.float32), topo, time=0.002 * (1 + np.arange(len(xyz))))
    make_loader.top = top
    return load


def main():
    patched = "--patched" in sys.argv
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import mdtraj as md
    info = {"mdtraj": os.path.dirname(md.__file__), "patched": patched}
    if patched:
        import mdtraj_b200
        mdtraj_b200.patch_mdtraj()
        import mdtraj_b200._rmsd as ours
        assert md.rmsd is ours.rmsd and md.rmsf is ours.rmsf and md._rmsd is ours
        from mdtraj_b200 import lprmsd_impl
        assert md.lprmsd is lprmsd_impl.lprmsd
    md.load = make_loader(md)
    results = {}
    mods = {}
    for fname, func, kwargs in CASES:
        tid = func + ("" if not kwargs else "[" + "-".join(f"{k}={v}" for k, v in kwargs.items()) + "]")
        try:
            if fname not in mods:
                spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(REF, "tests", fname))
                mods[fname] = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mods[fname])
            fn = getattr(mods[fname], func)
            if kwargs is None:
                fn()
            else:
                fn(get_fn=lambda name: name, **kwargs)
            results[tid] = "ok"
        except BaseException as e:  # noqa: BLE001 -- report, never abort the sweep
            results[tid] = "".join(traceback.format_exception_only(type(e), e)).strip()[-400:]
    if patched:
        from mdtraj_b200 import _capi
        info["native_library"] = _capi.LIB_PATH if _capi._lib is not None else None
    print(json.dumps({"info": info, "results": results}))


if __name__ == "__main__":
    main()
