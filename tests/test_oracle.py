"""CPU: pin the oracle.  The C restatement (oracle/oracle.c) is checked against
  * golden vectors produced by the real reference (md.rmsd / superpose / center_coordinates),
    including the two known answers printed in the reference's notebooks,
  * the compiled reference oracle/_ref (when it was built/shipped),
  * float64 Kabsch truth.
Tolerance: the reference's SSE build sums M in four lanes, the restatement sequentially, so agreement is to
float32 rounding noise: 1e-5 nm on iid / ala2 data (noise there is <= 3e-7)."""
import os

import numpy as np
import pytest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SYNTH = [("iid", 64, 100, 11), ("iid", 33, 22, 12), ("iid", 16, 1000, 13), ("md", 40, 303, 14), ("iid", 5, 4100, 15)]
IMPLS = ["port", "reference"]


def _impl_ok(O, impl):
    if impl == "reference" and not O.ref_available():
        pytest.skip("oracle/_ref was not built (needs /root/reference)")


@pytest.mark.parametrize("impl", IMPLS)
def test_ala2_golden(oracle_mod, golden, ala2, impl):
    O = oracle_mod
    _impl_ok(O, impl)
    d = O.rmsd(ala2, ala2, 0, impl=impl)
    assert np.abs(d[1:] - golden["ala2_rmsd_frame0"][1:]).max() < 1e-5
    D = np.stack([O.rmsd(ala2, ala2, i, impl=impl) for i in range(100)])
    off = ~np.eye(100, dtype=bool)
    assert np.abs(D - golden["ala2_allpairs"])[off].max() < 1e-5
    assert "%f" % D[off].max() == "0.188493"  # examples/clustering.ipynb:73
    heavy = golden["ala2_heavy_idx"]
    Dh = np.stack([O.rmsd(ala2, ala2, i, atom_indices=heavy, impl=impl) for i in range(100)])
    assert np.abs(Dh - golden["ala2_allpairs_heavy"])[off].max() < 1e-5
    np.fill_diagonal(Dh, np.diag(golden["ala2_allpairs_heavy"]))
    assert np.exp(-Dh / Dh.std()).sum(axis=1).argmax() == 83  # examples/centroids.ipynb:111


@pytest.mark.parametrize("impl", IMPLS)
def test_ala2_center_superpose_golden(oracle_mod, golden, ala2, impl):
    O = oracle_mod
    _impl_ok(O, impl)
    x = ala2.copy()
    tr = O.center_and_trace(x, impl)
    assert np.abs(x - golden["ala2_centered_xyz"]).max() < 1e-7
    assert np.allclose(tr, golden["ala2_traces"], rtol=1e-6, atol=0)
    d = O.rmsd(x, x, 5, target_traces=tr, ref_traces=tr, impl=impl)
    m = np.arange(100) != 5
    assert np.abs(d - golden["ala2_rmsd_frame5_precentered"])[m].max() < 1e-5
    assert np.abs(O.rmsd(ala2, ala2, 3, superpose=False, impl=impl) - golden["ala2_rmsd_frame3_nosuperpose"]).max() < 1e-6
    # superposed coordinates: the compiled reference must land on the real Trajectory.superpose output to float32
    # rounding (1e-6); the sequential-sum port is allowed the float32 noise of a 10-atom fit (3e-5 nm)
    tol = 1e-6 if impl == "reference" else 3e-5
    assert np.abs(O.superpose(ala2, ala2, 7, impl=impl) - golden["ala2_superposed_frame7"]).max() < tol
    assert np.abs(O.superpose(ala2, ala2, 2, golden["ala2_heavy_idx"], impl=impl)
                  - golden["ala2_superposed_frame2_heavy"]).max() < tol


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind,F,N,seed", SYNTH)
def test_synthetic_golden(oracle_mod, golden, impl, kind, F, N, seed):
    O = oracle_mod
    _impl_ok(O, impl)
    X = (O.synth_iid if kind == "iid" else O.synth_md)(F, N, seed=seed)
    key = f"{kind}_{F}x{N}_s{seed}"
    # MD-like data at a few hundred atoms: float32 noise of either implementation is ~1e-5 (Appendix C)
    tol = 1e-5 if kind == "iid" else 5e-5
    m = np.arange(F) != 1
    assert np.abs(O.rmsd(X, X, 1, impl=impl) - golden[key + "_rmsd_f1"])[m].max() < tol
    idx = np.arange(0, N, 3)
    m2 = np.arange(F) != 2
    assert np.abs(O.rmsd(X, X, 2, atom_indices=idx, impl=impl) - golden[key + "_rmsd_f2_idx3"])[m2].max() < tol
    assert np.abs(O.rmsd(X, X, 0, atom_indices=idx, ref_atom_indices=idx[::-1].copy(), impl=impl)
                  - golden[key + "_rmsd_f0_idx3_refrev"]).max() < tol
    assert np.abs(O.rmsd(X, X, 1, superpose=False, impl=impl) - golden[key + "_rmsd_f1_nosup"]).max() < 1e-5
    x = X.copy()
    tr = O.center_and_trace(x, impl)
    assert np.allclose(tr, golden[key + "_traces"], rtol=2e-6, atol=0)
    # the compiled reference must reproduce the real md.Trajectory.superpose output essentially exactly
    got = O.superpose(X, X, 0, impl=impl)
    truth = O.truth_superpose(X, X, 0)[0]
    e = np.abs(got - golden[key + "_superposed_f0"]).max()
    assert e < 1e-5 or np.abs(got - truth).max() <= 1.5 * np.abs(golden[key + "_superposed_f0"] - truth).max() + 1e-5


def test_reference_lib_reproduces_real_mdtraj_bitwise(oracle_mod, golden, ala2):
    """oracle/_ref is the reference's own arithmetic: outputs equal the real md.rmsd bit for bit."""
    O = oracle_mod
    _impl_ok(O, "reference")
    # md.rmsd(t, t, 0) aliases the reference frame with target frame 0 (views, _rmsd.pyx:197-199), so that frame
    # is centred twice; inplace=True on one shared array reproduces exactly that
    x = ala2.copy()
    d = O.rmsd(x, x, 0, impl="reference", inplace=True)
    assert np.array_equal(d, golden["ala2_rmsd_frame0"])
    X = O.synth_iid(64, 100, seed=11)
    assert np.array_equal(O.rmsd(X, X, 1, impl="reference", inplace=True), golden["iid_64x100_s11_rmsd_f1"])


def test_port_vs_truth_and_largest_root(oracle_mod):
    O = oracle_mod
    X = O.synth_iid(50, 64, seed=2)
    t = O.truth_rmsd(X, X, 3)
    m = np.arange(50) != 3
    assert np.abs(O.rmsd(X, X, 3, impl="port") - t)[m].max() < 1e-5
    # quartic with known roots: (t-1)(t-2)(t+0.5)(t+2.5) has c3 == 0
    r = np.array([1.0, 2.0, -0.5, -2.5])
    c = np.poly(r)
    assert abs(c[1]) < 1e-12
    assert abs(O.port_lib().oracle_qcp_largest_root(c[4], c[3], c[2]) - 2.0) < 1e-12


def test_rotation_convention(oracle_mod):
    """rot maps the first argument onto the second as row-vector x matrix (rotation_generic.h:40-42)."""
    O = oracle_mod
    rng = np.random.default_rng(0)
    b = rng.standard_normal((30, 3)); b -= b.mean(0)
    R = O.random_rotations(1, rng)[0]
    a = (b @ R.T).astype(np.float32); b = b.astype(np.float32)  # a @ R == b
    msd, rot = O.msd_atom_major(a, b, float((a * a).sum()), float((b * b).sum()), want_rot=True)
    assert msd < 1e-6 and np.abs(rot - R).max() < 1e-5 and abs(np.linalg.det(rot.astype(float)) - 1) < 1e-5


# ------------------------------------------------------------------ md.lprmsd (last "next" row)
def lprmsd_case(gold, name):
    """(xyz, ref, atom_indices or None, permute_groups or None) of a case of tests/golden/lprmsd_outputs.npz."""
    idx = gold[name + "_idx"] if bool(gold[name + "_has_idx"]) else None
    groups = None
    if bool(gold[name + "_has_groups"]):
        flat, lens = gold[name + "_groups_flat"], gold[name + "_groups_len"]
        offs = np.concatenate([[0], np.cumsum(lens)])
        groups = [flat[offs[i]: offs[i + 1]] for i in range(len(lens))]
    return gold[name + "_xyz"], gold[name + "_ref"], idx, groups


def test_min_cost_matching_pinned_to_munkres(oracle_mod):
    """The oracle's exact assignment solver returns what Munkres::solve returns: the reference's own known answer
    (tests/test_lprmsd.py:12-22) and six random 12x12 cost matrices run through mdtraj._lprmsd._munkres."""
    O = oracle_mod
    gold = np.load(os.path.join(GOLDEN_DIR, "lprmsd_outputs.npz"))
    known = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]], dtype=np.int32)
    assert np.array_equal(gold["munkres_known"], known)
    m = O.min_cost_matching(np.array([[7, 4, 3], [6, 8, 5], [9, 4, 4]], dtype=np.float64))
    assert np.array_equal(np.eye(3, dtype=np.int32)[m], known)
    for cost, mask in zip(gold["munkres_costs"], gold["munkres_masks"]):
        assert np.array_equal(np.eye(12, dtype=np.int32)[O.min_cost_matching(cost)], mask)


@pytest.mark.parametrize("impl", ["port", "reference"])
def test_lprmsd_oracle_reproduces_real_mdtraj(oracle_mod, impl):
    """oracle.lprmsd (restatement of _lprmsd.pyx:131-227) against the outputs of the real md.lprmsd on every golden
    case, distances and superposed coordinates: bit for bit on the compiled reference's arithmetic, 3e-5 / 1e-5 on the port."""
    O = oracle_mod
    _impl_ok(O, impl)
    gold = np.load(os.path.join(GOLDEN_DIR, "lprmsd_outputs.npz"))
    for name in gold["cases"]:
        X, ref, idx, groups = lprmsd_case(gold, str(name))
        d = O.lprmsd(X, ref, 0, idx, groups, impl=impl)
        d2, xyz = O.lprmsd(X, ref, 0, idx, groups, superpose=True, impl=impl)
        if impl == "reference":
            assert np.array_equal(d, gold[f"{name}_lprmsd"]), name
            assert np.array_equal(d2, gold[f"{name}_lprmsd_superpose"]), name
            assert np.abs(xyz - gold[f"{name}_xyz_superposed"]).max() < 2e-6, name  # rot_atom_major: SSE vs numpy order
        else:
            # float32 msd = (G_a + G_b - 2 lambda) / n in both, summed in a different order: at an RMSD of 0.03 nm under
            # traces of ~50 nm^2 that is ~1e-5 nm of noise (the same class as SURVEY.md Appendix C)
            assert np.abs(d - gold[f"{name}_lprmsd"]).max() < 3e-5, name
            assert np.abs(d2 - gold[f"{name}_lprmsd_superpose"]).max() < 3e-5, name
            assert np.abs(xyz - gold[f"{name}_xyz_superposed"]).max() < 1e-5, name
