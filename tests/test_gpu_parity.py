"""GPU parity tests: the CUDA path (through the C ABI) against
  (1) golden vectors produced by the real reference (tests/golden/make_golden.py),
  (2) the CPU oracle (oracle/, C restatement + compiled reference when shipped),
  (3) float64 Kabsch truth,
on the same seeded inputs.

Tolerances (BASELINE.json north_star): RMSD 1e-5 nm absolute or 1e-4 relative; rotation-matrix
elements 1e-5; superposed coordinates 1e-5 nm (float32 coordinates of magnitude <= ~10 nm carry
~5e-7 nm of rounding each way); atom selection / frame order exact.
Mirrors the reference's tests/test_rmsd.py where noted.
"""
import os
import warnings

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-5, 1e-4
SYNTH = [("iid", 64, 100, 11), ("iid", 33, 22, 12), ("iid", 16, 1000, 13), ("md", 40, 303, 14), ("iid", 5, 4100, 15)]


def close(a, b, atol=ATOL, rtol=RTOL):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return bool(np.all(np.abs(a - b) <= np.maximum(atol, rtol * np.abs(b))))


def assert_close(a, b, atol=ATOL, rtol=RTOL, what=""):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    ok = err <= np.maximum(atol, rtol * np.abs(b))
    assert ok.all(), f"{what}: max abs err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def assert_three_way(got, ref, truth, what="", atol=ATOL):
    """Parity where the reference itself is float32-noisy (ill-conditioned rotations on
    iid data, low-RMSD MD-like data at large N): pass when we match the reference within tolerance, or when
    we are at least as close to the float64 truth as the reference is."""
    got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64); truth = np.asarray(truth, np.float64)
    e_ref = np.abs(got - ref).max(); e_truth = np.abs(got - truth).max(); ref_truth = np.abs(ref - truth).max()
    assert e_ref <= atol or e_truth <= max(atol, 1.5 * ref_truth), \
        f"{what}: |gpu-ref|={e_ref:.3e} |gpu-truth|={e_truth:.3e} |ref-truth|={ref_truth:.3e}"


def assert_close_msd(got, truth, xyz_sel, what=""):
    """For selections whose RMSD can be ~0 between different frames (two atoms: rmsd = |bond_a - bond_b| / 2): the
    quantity float32 arithmetic resolves is N msd = G_a + G_b - 2 lambda against (G_a + G_b) / 2, so the check is on
    rmsd^2 against the mean-square radius of the selection (8 float32 ulp); sqrt() turns that into ~1e-4 nm at rmsd = 0
    (the reference returns 8e-4 ... 6e-3 nm for identical frames in separate memory)."""
    X = np.asarray(xyz_sel, np.float64)
    rg2 = ((X - X.mean(1, keepdims=True)) ** 2).sum((1, 2)).mean() / X.shape[1]
    err = np.abs(np.asarray(got, np.float64) ** 2 - np.asarray(truth, np.float64) ** 2)
    assert err.max() <= 1e-6 * rg2, f"{what}: max |rmsd^2 - truth^2| = {err.max():.3e} vs Rg^2 = {rg2:.3e}"


def gen(O, kind, F, N, seed):
    return (O.synth_iid if kind == "iid" else O.synth_md)(F, N, seed=seed)


@pytest.fixture
def ap_path():
    """Select the all-pairs kernel through the public setting (b200rmsd_allpairs_configure): "tc" = tcgen05 for any
    size, "simt" = exact-fp32 SIMT for any size, None = the default switch at 512 frames."""
    from mdtraj_b200 import allpairs as AP

    def select(name):
        AP.configure(min_tc_frames={"tc": 1, "simt": 1 << 30, None: 512}[name])
    yield select
    AP.configure(min_tc_frames=512)


def unmirrored_matrix(dt, atom_indices=None):
    """The matrix with EVERY entry computed (two row blocks; a full-matrix call computes each unordered pair once and
    mirrors it)."""
    import torch
    from mdtraj_b200 import allpairs as AP
    prep = AP.prepare(dt, atom_indices)
    h = dt.n_frames // 2
    return torch.cat([AP.rows(prep, 0, h), AP.rows(prep, h, dt.n_frames)])


# ------------------------------------------------------------------ C1: ala2 golden
def test_ala2_rmsd_matches_reference(mdb, golden, ala2):
    t = mdb.Trajectory(ala2.copy())
    d = mdb.rmsd(t, t, 0)
    assert d.dtype == np.float32 and d.shape == (100,)
    assert d[0] == 0.0
    assert_close(d, golden["ala2_rmsd_frame0"], what="ala2 rmsd vs reference")
    assert abs(float(d.max()) - 0.16302569) < 1e-5


def test_ala2_device_path_and_precentered(mdb, golden, ala2):
    dt = mdb.DeviceTrajectory.from_host(ala2)
    d = mdb.rmsd(dt, dt, 0)
    assert_close(d, golden["ala2_rmsd_frame0"], what="device path")
    dt.center_coordinates()
    assert_close(dt.xyz, golden["ala2_centered_xyz"], atol=2e-7, rtol=0, what="centred coordinates")
    assert_close(dt._rmsd_traces.cpu().numpy(), golden["ala2_traces"], atol=0, rtol=1e-6, what="traces")
    d5 = mdb.rmsd(dt, dt, 5, precentered=True)
    assert_close(d5, golden["ala2_rmsd_frame5_precentered"], what="precentered")
    # host precentered path
    t = mdb.Trajectory(ala2.copy())
    t.center_coordinates()
    assert_close(t.xyz, golden["ala2_centered_xyz"], atol=2e-7, rtol=0, what="host centred coordinates")
    assert_close(mdb.rmsd(t, t, 5, precentered=True), golden["ala2_rmsd_frame5_precentered"], what="host precentered")


def test_ala2_nosuperpose(mdb, golden, ala2):
    t = mdb.Trajectory(ala2.copy())
    assert_close(mdb.rmsd(t, t, 3, superpose=False), golden["ala2_rmsd_frame3_nosuperpose"], what="superpose=False")


def test_ala2_allpairs_known_answers(mdb, golden, ala2):
    """examples/clustering.ipynb:73 (0.188493) and examples/centroids.ipynb:111 (83)."""
    t = mdb.Trajectory(ala2.copy())
    D = mdb.rmsd_matrix(t)
    assert D.shape == (100, 100) and D.dtype == np.float32
    assert_close(D, golden["ala2_allpairs"], what="all-pairs vs reference loop")
    assert "%f" % D.max() == "0.188493"
    assert np.all(np.diag(D) == 0)
    assert np.abs(D - D.T).max() < 1e-6  # clustering.ipynb cell 8
    heavy = golden["ala2_heavy_idx"]
    Dh = mdb.rmsd_matrix(t, atom_indices=heavy)
    off = ~np.eye(100, dtype=bool)  # with an index list the reference's diagonal is float32 noise, not 0
    assert_close(Dh[off], golden["ala2_allpairs_heavy"][off], what="heavy-atom all-pairs")
    index = np.exp(-1 * Dh / Dh.std()).sum(axis=1).argmax()
    assert index == 83 == int(golden["ala2_centroid_index"])


def test_ala2_rows_equal_one_vs_many(mdb, ala2):
    dt = mdb.DeviceTrajectory.from_host(ala2)
    D = mdb.rmsd_matrix(dt)
    for i in (0, 17, 99):
        assert_close(D[i], mdb.rmsd(dt, dt, i), atol=2e-6, what=f"row {i}")


def test_ala2_superpose(mdb, golden, ala2):
    t = mdb.Trajectory(ala2.copy()); r = mdb.Trajectory(ala2.copy())
    ret = t.superpose(r, 7)
    assert ret is t and t._rmsd_traces is None
    assert np.array_equal(r.xyz, ala2)  # reference never mutated (tests/test_rmsd.py:131-140)
    assert_close(t.xyz, golden["ala2_superposed_frame7"], what="superpose all atoms")
    t = mdb.Trajectory(ala2.copy())
    t.superpose(r, 2, atom_indices=golden["ala2_heavy_idx"])
    assert_close(t.xyz, golden["ala2_superposed_frame2_heavy"], what="superpose heavy atoms")


# ------------------------------------------------------------------ seeded synthetic golden
@pytest.mark.parametrize("kind,F,N,seed", SYNTH)
def test_synthetic_golden(mdb, golden, oracle_mod, kind, F, N, seed):
    O = oracle_mod
    X = gen(O, kind, F, N, seed)
    key = f"{kind}_{F}x{N}_s{seed}"
    idx = np.arange(0, N, 3)
    t = mdb.Trajectory(X.copy())
    d = mdb.rmsd(t, t, 1)
    assert np.array_equal(t.xyz, X)  # host arrays untouched by default (documented deviation)
    assert_close(d, golden[key + "_rmsd_f1"], what="rmsd")
    # three-way report against float64 truth
    truth = O.truth_rmsd(X, X, 1)
    e_gpu, e_ref = np.abs(d - truth).max(), np.abs(golden[key + "_rmsd_f1"] - truth)[np.arange(F) != 1].max()
    assert e_gpu <= max(ATOL, 2 * e_ref), (e_gpu, e_ref)
    # with an index list the reference's frame-vs-itself value is float32 noise (up to ~1e-3 nm, Appendix B #15:
    # the pointer shortcut cannot fire on the fancy-index copy); the truth is 0, so that entry is checked against 0
    not2 = np.arange(F) != 2
    d_idx = mdb.rmsd(t, t, 2, atom_indices=idx)
    assert_close(d_idx[not2], golden[key + "_rmsd_f2_idx3"][not2], what="atom_indices")
    assert d_idx[2] < 2e-3
    assert_close(mdb.rmsd(t, t, 0, atom_indices=idx, ref_atom_indices=idx[::-1].copy()),
                 golden[key + "_rmsd_f0_idx3_refrev"], what="ref_atom_indices")
    assert_close(mdb.rmsd(t, t, 1, superpose=False), golden[key + "_rmsd_f1_nosup"], what="superpose=False")
    # device-resident path gives the same numbers as the host-streamed path
    dt = mdb.DeviceTrajectory.from_host(X)
    assert np.array_equal(mdb.rmsd(dt, dt, 1), d)
    assert np.array_equal(mdb.rmsd(dt, dt, 2, atom_indices=idx), mdb.rmsd(t, t, 2, atom_indices=idx))
    # superpose
    a = mdb.Trajectory(X.copy()); r = mdb.Trajectory(X.copy())
    a.superpose(r, 1, atom_indices=idx)
    assert_three_way(a.xyz, golden[key + "_superposed_f1_idx3"], O.truth_superpose(X, X, 1, idx)[0], "superpose idx")
    b = mdb.Trajectory(X.copy())
    b.superpose(r, 0)
    truth0 = O.truth_superpose(X, X, 0)[0]
    assert_three_way(b.xyz, golden[key + "_superposed_f0"], truth0, "superpose all")
    dt.superpose(dt, 0)
    assert_three_way(dt.xyz, golden[key + "_superposed_f0"], truth0, "device superpose")
    # centring
    c = mdb.Trajectory(X.copy())
    c.center_coordinates()
    assert_close(c._rmsd_traces, golden[key + "_traces"], atol=0, rtol=1e-6, what="traces")
    assert np.abs(c.xyz.mean(1)).max() < 1e-6  # tests/test_trajectory.py:316-344
    assert_close(mdb.rmsd(c, c, 1, precentered=True), golden[key + "_rmsd_f1"], what="precentered == on the fly")


# ------------------------------------------------------------------ oracle on fresh inputs
@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 7, 31, 32, 33, 127, 128, 255, 1000, 4096, 4100, 9000])
def test_rmsd_vs_oracle_ragged_sizes(mdb, oracle_mod, N):
    """Every padding remainder, sub-warp frames, the segment boundary (4096) and multi-segment frames."""
    O = oracle_mod
    F = 37 if N < 2000 else 9
    X = O.synth_iid(F, N, seed=100 + N) + np.float32(3.0)  # off-origin on purpose
    impl = "reference" if O.ref_available() else "port"
    want = O.rmsd(X, X, 2, impl=impl)
    truth = O.truth_rmsd(X, X, 2)
    for target in (mdb.Trajectory(X.copy()), mdb.DeviceTrajectory.from_host(X)):
        got = mdb.rmsd(target, target, 2)
        m = np.arange(F) != 2
        if N >= 3:
            assert_three_way(got[m], want[m], truth[m], f"N={N} vs oracle({impl})")
        # float32 products put a floor of ~eps32 * <r^2> / rmsd under any float32-input QCP
        floor = 3e-7 * float((X.astype(np.float64) - X.mean(1, keepdims=True)).var() * 3) / np.maximum(truth[m], 1e-4)
        err = np.abs(got[m].astype(np.float64) - truth[m])
        assert np.all(err <= np.maximum(np.maximum(ATOL, RTOL * truth[m]), floor)), f"N={N} vs truth: {err.max():.3e}"
        assert got[2] == 0.0


def test_rotation_parity_md(mdb, oracle_mod):
    """Rotation matrices on well-conditioned (MD-like) data: 1e-5 per element vs reference and truth."""
    O = oracle_mod
    X = O.synth_md(48, 1000, seed=5, rg=1.5, sigma=0.1)
    idx = np.arange(0, 1000, 5)
    impl = "reference" if O.ref_available() else "port"
    want_xyz, want_R = O.superpose(X, X, 0, idx, impl=impl, return_rot=True)
    truth_xyz, truth_R = O.truth_superpose(X, X, 0, idx)
    dt = mdb.DeviceTrajectory.from_host(X)
    _, R = dt.superpose(dt, 0, atom_indices=idx, return_rotations=True)
    R = R.cpu().numpy()
    assert np.abs(R - truth_R).max() < 1e-5
    # the reference's own float32 rotation is up to ~7e-5 from the truth on this input (three-way report)
    assert_three_way(R, want_R, truth_R, "rotation elements")
    assert_three_way(dt.xyz, want_xyz, truth_xyz, "superposed coordinates")
    assert_close(dt.xyz, truth_xyz, what="superposed coordinates vs truth")
    assert np.abs(np.linalg.det(R.astype(np.float64)) - 1).max() < 1e-5
    # after superposition the plain RMSD equals the QCP RMSD (tests/test_rmsd.py:98-108)
    ref = mdb.Trajectory(X.copy())
    plain = mdb.rmsd(dt, ref, 0, atom_indices=idx, superpose=False)
    qcp = mdb.rmsd(mdb.Trajectory(X.copy()), ref, 0, atom_indices=idx)
    assert_close(plain[1:], qcp[1:], atol=2e-5, what="superpose then plain rmsd")


def test_superpose_semantics(mdb):
    """tests/test_rmsd.py:111-171, 319-334 restated on the new API."""
    rng = np.random.RandomState(52)
    t1 = mdb.Trajectory(rng.randn(10, 100, 3).astype(np.float32))
    t2 = mdb.Trajectory(rng.randn(10, 100, 3).astype(np.float32) + 100)
    t2_copy = t2.xyz.copy()
    t1.superpose(t2)
    t1.superpose(t2, atom_indices=[1, 2, 3, 4, 5, 6, 7])
    assert np.array_equal(t2.xyz, t2_copy)
    assert 99 < t1.xyz.mean() < 101  # translated onto the reference centroid
    # ref_atom_indices: swapping halves of the reference makes the superposition a no-op
    n = 20
    half = rng.randn(1, n // 2, 3).astype(np.float32)
    other = rng.randn(1, n // 2, 3).astype(np.float32)
    a = mdb.Trajectory(np.concatenate([half, other], axis=1))
    b = mdb.Trajectory(np.concatenate([other * 3 + 1, half], axis=1))
    before = a.xyz.copy()
    a.superpose(b, 0, atom_indices=np.arange(n // 2), ref_atom_indices=np.arange(n // 2, n))
    assert np.abs(a.xyz - before).max() < 2e-6
    with pytest.raises(ValueError):
        a.superpose(b, 0, atom_indices=[])
    one = mdb.Trajectory(rng.randn(3, 10, 3).astype(np.float32))
    one.superpose(one, 0, atom_indices=[4])
    assert np.isfinite(one.xyz).all()


def test_different_atom_counts_with_index_lists(mdb, oracle_mod):
    """tests/test_rmsd.py:291-316."""
    O = oracle_mod
    A = O.synth_iid(12, 40, seed=1); B = O.synth_iid(3, 55, seed=2)
    ia = np.array([3, 1, 9, 30, 30, 7]); ib = np.array([50, 2, 2, 11, 54, 0])  # unsorted + repeated
    got = mdb.rmsd(mdb.Trajectory(A.copy()), mdb.Trajectory(B.copy()), 2, atom_indices=ia, ref_atom_indices=ib)
    want = O.truth_rmsd(A, B, 2, ia, ib)
    assert_close(got, want, what="index lists on different atom counts")
    # atom-count mismatch without index lists: the reference means to raise ValueError (_rmsd.pyx:178-181) but
    # trips over len(slice) first (TypeError); we raise the same type it does.
    with pytest.raises(TypeError):
        mdb.rmsd(mdb.Trajectory(A.copy()), mdb.Trajectory(B.copy()), 0)


def test_warnings_and_readonly(mdb, golden, oracle_mod):
    X = oracle_mod.synth_iid(5, 10, 1)
    t = mdb.Trajectory(X.copy())
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mdb.rmsd(t, t, 0, precentered=True)
    assert f"{w[-1].category.__name__}: {w[-1].message}" == str(golden["msg_warn_precentered_no_traces"])
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mdb.rmsd(t, t, 0, precentered=True, superpose=False)
    assert f"{w[-1].category.__name__}: {w[-1].message}" == str(golden["msg_warn_precentered_nosuperpose"])
    ro = X.copy(); ro.setflags(write=False)

    class Duck:
        xyz = ro
        _rmsd_traces = None
    with pytest.raises(ValueError, match="read-only"):
        mdb.rmsd(Duck(), Duck(), 0)
    assert mdb.rmsd(Duck(), Duck(), 0, atom_indices=[0, 1, 2, 3]).shape == (5,)  # copy path works
    assert mdb.rmsd(t, t, -1).shape == (5,)  # negative frame is python indexing


def test_inplace_centering_compat(mdb, oracle_mod):
    X = oracle_mod.synth_iid(6, 30, 3) + np.float32(2.0)
    t = mdb.Trajectory(X.copy())
    mdb.set_inplace_centering(True)
    try:
        mdb.rmsd(t, t, 0)
    finally:
        mdb.set_inplace_centering(False)
    assert np.abs(t.xyz.mean(1)).max() < 1e-6


def test_empty_and_single_frame(mdb, oracle_mod):
    X = oracle_mod.synth_iid(1, 17, 4)
    t = mdb.Trajectory(X.copy())
    assert mdb.rmsd(t, t, 0).tolist() == [0.0]
    e = mdb.Trajectory(np.zeros((0, 17, 3), np.float32))
    assert mdb.rmsd(e, t, 0).shape == (0,)


# ------------------------------------------------------------------ large sizes: size-independent properties
def test_large_properties(mdb):
    """At BASELINE-like sizes: (a) rigid motions leave RMSD unchanged; (b) RMSD to a noisy copy equals the
    injected noise level; (c) frame order is preserved; (d) sharded == unsharded bit for bit."""
    import torch
    F, N = 20000, 1000
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=7)
    same = mdb.rmsd_device(dt, dt, 0, as_numpy=False)
    assert same[0].item() == 0.0 and torch.isfinite(same).all()
    ref = mdb.DeviceTrajectory(dt.xyz_dev[:1].clone(), N)
    base = mdb.rmsd_device(dt, ref, 0, as_numpy=False)
    assert torch.equal(base[1:], same[1:])
    # (a) rotate + translate every frame by its own rigid motion
    g = torch.Generator(device=dt.device); g.manual_seed(1)
    q = torch.randn((F, 4), generator=g, device=dt.device, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    a, b, c, d = q.unbind(1)
    R = torch.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c), 2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b),
                     2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], dim=1).view(F, 3, 3)
    moved = torch.bmm(dt.xyz_dev.double(), R) + torch.rand((F, 1, 3), generator=g, device=dt.device, dtype=torch.float64) * 10 - 5
    dm = mdb.DeviceTrajectory(moved.float().contiguous(), N)
    moved_r = mdb.rmsd_device(dm, ref, 0, as_numpy=False)
    assert (moved_r[1:] - base[1:]).abs().max().item() < 1e-5
    # (c) order: permute frames, results permute identically
    perm = torch.randperm(F, generator=torch.Generator().manual_seed(3)).to(dt.device)
    dp = mdb.DeviceTrajectory(dt.xyz_dev[perm].contiguous(), N)
    pr = mdb.rmsd_device(dp, ref, 0, as_numpy=False)
    assert torch.equal(pr, base[perm])
    # (d) two halves computed separately == one call
    h1 = mdb.rmsd_device(dt[: F // 2], ref, 0, as_numpy=False); h2 = mdb.rmsd_device(dt[F // 2:], ref, 0, as_numpy=False)
    assert torch.equal(torch.cat([h1, h2]), base)
    # (b) known noise level
    noisy = dt.xyz_dev[:1] + 0.05 * torch.randn((2000, N, 3), generator=g, device=dt.device)
    dn = mdb.DeviceTrajectory(noisy.contiguous(), N)
    r = mdb.rmsd_device(dn, ref, 0, as_numpy=False)
    assert abs(r.mean().item() - 0.05 * (3 ** 0.5)) < 2e-3


def test_allpairs_vs_truth_random(mdb, oracle_mod):
    O = oracle_mod
    X = O.synth_md(70, 300, seed=9, rg=1.0, sigma=0.15)
    D = mdb.rmsd_matrix(mdb.Trajectory(X.copy()))
    for i in (0, 33, 69):
        truth = O.truth_rmsd(X, X, i)
        m = np.arange(70) != i
        assert_close(D[i][m], truth[m], what=f"all-pairs row {i} vs truth")
    assert np.abs(D - D.T).max() < 1e-6


# ------------------------------------------------------------------ tensor-core all-pairs path
@pytest.mark.parametrize("F,N", [(600, 300), (1001, 97), (520, 22), (2100, 30)])
def test_allpairs_tensor_core_vs_simt_and_truth(mdb, oracle_mod, ap_path, F, N):
    """The tcgen05 3xTF32 kernels (F >= 512) against the exact-fp32 SIMT kernel and float64 truth.  F = 2100 spans
    several super-blocks of the tile walk (53 x 44 tiles of 40 x 48 frames) with ragged edges both ways."""
    O = oracle_mod
    X = O.synth_md(F, N, seed=21, rg=1.0, sigma=0.15)
    dt = mdb.DeviceTrajectory.from_host(X)
    ap_path("tc")
    D_tc = mdb.rmsd_matrix(dt)
    ap_path("simt")
    D_simt = mdb.rmsd_matrix(dt)
    ap_path("tc")
    assert D_tc.shape == (F, F) and np.isfinite(D_tc).all()
    assert np.all(np.diag(D_tc) == 0)
    assert_close(D_tc, D_simt, atol=5e-6, what="tcgen05 vs SIMT all-pairs")
    assert np.abs(D_tc - D_tc.T).max() < 2e-6
    for i in (0, F // 2, F - 1):
        truth = O.truth_rmsd(X, X, i)
        m = np.arange(F) != i
        assert_close(D_tc[i][m], truth[m], what=f"tcgen05 row {i} vs truth")
    # all-float32 solve: reference-class precision, still inside the parity tolerance on this data
    D_fast = mdb.rmsd_matrix(dt, precise=False)
    assert_close(D_fast, D_simt, what="tcgen05 float32 solve vs SIMT")
    # a row block, as a rank of a sharded run computes it
    from mdtraj_b200 import allpairs as AP
    prep = AP.prepare(dt)
    blk = AP.rows(prep, 37, 123).cpu().numpy()
    # full-matrix calls compute each unordered pair once and mirror it (exactly symmetric); row blocks compute
    # every entry: identical to the unmirrored full matrix bit for bit, and to the mirrored one within float32 noise
    assert np.array_equal(D_tc, D_tc.T)
    D_full = unmirrored_matrix(dt).cpu().numpy()
    assert np.array_equal(blk, D_full[37:123])
    assert_close(blk, D_tc[37:123], atol=2e-6, what="row block vs mirrored full matrix")


def _sampled_rows_vs_truth(O, X, D, rows, atom_indices=None):
    """max |D[i] - float64 truth| over the sampled rows (a frame against itself left out), and the same for the compiled
    reference / C port run as the notebook does (md.rmsd(traj, traj, i), examples/clustering.ipynb:78-81)."""
    kind = "reference" if O.ref_available() else "port"
    F = len(X)
    e_gpu = e_ref = e_gr = 0.0
    for i in rows:
        truth = O.truth_rmsd(X, X, i, atom_indices=atom_indices)
        ref = O.rmsd(X, X, i, atom_indices=atom_indices, impl=kind)
        m = np.arange(F) != i
        e_gpu = max(e_gpu, float(np.abs(D[i][m] - truth[m]).max()))
        e_ref = max(e_ref, float(np.abs(ref[m] - truth[m]).max()))
        e_gr = max(e_gr, float(np.abs(D[i][m] - ref[m]).max()))
    return e_gpu, e_ref, e_gr


@pytest.mark.parametrize("n_basins,interleave", [(2, False), (3, False), (3, True)])
def test_allpairs_multi_basin(mdb, oracle_mod, n_basins, interleave):
    """Clustering input: MD-like frames in 2-3 basins 1.4 nm apart, 300 atoms, F = 2400 (round 1 aligned everything onto
    frame 0 and was 4.6e-5 nm off inside the other basins).  Every sampled row -- at least two per basin -- within
    1e-5 nm of the float64 truth AND of the reference's md.rmsd(traj, traj, i); one reference structure per basin."""
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    F, N = 2400, 300
    X, which = O.synth_md_basins(F, N, n_basins, seed=31 + n_basins, rg=1.0, sigma=0.1, separation=1.4, interleave=interleave)
    dt = mdb.DeviceTrajectory.from_host(X)
    prep = AP.prepare(dt)
    info = prep.info()
    assert n_basins <= info["n_refs"] <= n_basins + 2 and info["n_far"] == 0 and info["cover_radius"] < 0.3, info
    assert sorted(set(which[info["ref_frames"][:n_basins]])) == list(range(n_basins)), "one reference per basin first"
    D = AP.rows(prep, 0, F).cpu().numpy()
    rows = [int(np.flatnonzero(which == b)[k]) for b in range(n_basins) for k in (0, 7, -1)]
    e_gpu, e_ref, e_gr = _sampled_rows_vs_truth(O, X, D, rows)
    assert e_gpu < 1e-5 and e_gr < 1e-5, (e_gpu, e_ref, e_gr)
    assert e_gpu < 4e-6, f"multi-reference operands should be in the 1e-6 class: {e_gpu:.2e} (reference: {e_ref:.2e})"
    assert np.array_equal(D, D.T) and np.all(np.diag(D) == 0)
    # the same numbers through the public call, with an atom selection, and from a row block of a sharded run
    idx = np.arange(0, N, 2)
    Ds = mdb.rmsd_matrix(dt, atom_indices=idx)
    e_gpu, e_ref, e_gr = _sampled_rows_vs_truth(O, X, Ds, rows[:3], atom_indices=idx)
    assert e_gpu < 1e-5 and e_gr < 1e-5, (e_gpu, e_ref, e_gr)
    blk = AP.rows(prep, 1000, 1100).cpu().numpy()
    assert np.abs(blk - D[1000:1100]).max() < 5e-6  # mirrored vs computed orientation: two results of the 1e-6 class


def test_allpairs_single_reference_would_fail_multi_basin(mdb, oracle_mod):
    """The same input with the traversal capped at ONE reference reproduces round 1's error class inside the second
    basin (> 1e-5 nm): the test above is green because of the references, not because the data is easy."""
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    X, which = O.synth_md_basins(2400, 300, 2, seed=33, rg=1.0, sigma=0.1, separation=1.4)
    dt = mdb.DeviceTrajectory.from_host(X)
    AP.configure(max_refs=1)
    try:
        prep = AP.prepare(dt)
        assert prep.info()["n_refs"] == 1
        D = AP.rows(prep, 0, 2400).cpu().numpy()
    finally:
        AP.configure(max_refs=32)
    second = np.flatnonzero(which == 1)
    i = int(second[3])
    truth = O.truth_rmsd(X, X, i)
    m = second[second != i]
    assert np.abs(D[i][m] - truth[m]).max() > 1e-5


def test_allpairs_drifting_trajectory(mdb, oracle_mod):
    """A trajectory that drifts 0.8 nm away from frame 0 with neighbouring frames 0.02 nm apart: every frame is 'near'
    frame 0, but the error of the difference operands grows like delta^2 / rmsd_ij, so the traversal keeps adding
    references along the path until the covering radius is below 0.25 nm.  Neighbouring-frame RMSDs (the small ones, where
    float32 inner products are weakest) must stay within tolerance of the float64 truth or beat the reference."""
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    F, N = 2000, 300
    X = O.synth_md_drift(F, N, seed=2, step=0.01)
    dt = mdb.DeviceTrajectory.from_host(X)
    prep = AP.prepare(dt)
    info = prep.info()
    assert 3 <= info["n_refs"] <= 32 and info["n_far"] == 0 and info["cover_radius"] <= 0.3, info
    D = AP.rows(prep, 0, F).cpu().numpy()
    e_gpu, e_ref, e_gr = _sampled_rows_vs_truth(O, X, D, [3, 700, 1400, 1998])
    assert e_gpu <= max(1e-5, 1.5 * e_ref), (e_gpu, e_ref, e_gr)


def test_allpairs_iid_frames_stop_the_traversal(mdb, oracle_mod):
    """iid frames are far from everything: references capture nothing, the traversal stops after four, all frames keep
    plain operands and the numbers stay in the 1e-6 class (large RMSDs, no cancellation)."""
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    X = O.synth_iid(1500, 300, seed=5)
    prep = AP.prepare(mdb.DeviceTrajectory.from_host(X))
    info = prep.info()
    assert info["n_refs"] == 4 and info["n_far"] == 1500 - 4, info
    D = AP.rows(prep, 0, 1500).cpu().numpy()
    e_gpu, e_ref, e_gr = _sampled_rows_vs_truth(O, X, D, [0, 777, 1499])
    assert e_gpu < 3e-6 and e_gr < 1e-5, (e_gpu, e_ref, e_gr)


def test_allpairs_tensor_core_window_sweep(mdb, oracle_mod, ap_path):
    """Row/column windows that start and end anywhere relative to the 40- and 48-frame tile grids, incl. diagonal
    squares off the origin (symmetric mode with different i- and j-grid origins): every entry equals the full matrix."""
    import torch
    from mdtraj_b200 import allpairs as AP
    ap_path("tc")
    X = oracle_mod.synth_md(1300, 40, seed=44, rg=0.8, sigma=0.1)
    dt = mdb.DeviceTrajectory.from_host(X)
    full = unmirrored_matrix(dt)
    prep = AP.prepare(dt)
    for (r0, r1, c0, c1) in [(0, 1300, 0, 1300), (50, 1251, 50, 1251), (47, 1001, 47, 1001), (960, 1300, 960, 1300),
                             (1, 2, 0, 1300), (0, 1300, 1299, 1300), (39, 41, 47, 49), (1000, 1300, 0, 500),
                             (119, 121, 143, 145)]:
        out = torch.full((r1 - r0, 1300), -1.0, dtype=torch.float32, device=dt.device)
        AP.block(prep, r0, r1, c0, c1, out)
        got, want = out[:, c0:c1], full[r0:r1, c0:c1]
        if (r0, r1) == (c0, c1):  # mirrored halves: exactly symmetric, equal to the unmirrored values within float32 noise
            assert torch.equal(got, got.t())
            assert torch.equal(torch.triu(got), torch.triu(want))
            assert (got - want).abs().max().item() < 2e-6
        else:
            assert torch.equal(got, want), (r0, r1, c0, c1)
        assert (out[:, :c0] == -1).all() and (out[:, c1:] == -1).all(), "wrote outside the column window"


def test_allpairs_tensor_core_degenerate_geometries(mdb, oracle_mod, ap_path):
    """Double largest root of the QCP quartic -- atoms on a line, two-atom selections -- where Newton alone lands on
    the wrong root (tests/test_qcp_host.py); the epilogue's certificate sends these pairs to the closed form."""
    O = oracle_mod
    ap_path("tc")
    rng = np.random.default_rng(7)
    F = 640
    line = rng.standard_normal((30, 1)) * np.array([[1.0, 0.0, 0.0]]) + 0.05 * rng.standard_normal((F, 30, 3))
    line = np.einsum("fnk,fkl->fnl", line, O.random_rotations(F, np.random.default_rng(8))).astype(np.float32)
    for what, X, idx in (("line", line, None), ("two atoms", O.synth_md(F, 20, seed=5, rg=0.5, sigma=0.1), [3, 11])):
        D = mdb.rmsd_matrix(mdb.Trajectory(X.copy()), atom_indices=idx)
        assert np.isfinite(D).all()
        for i in (0, 17, F - 1):
            truth = O.truth_rmsd(X, X, i, atom_indices=idx)
            m = np.arange(F) != i
            # float32 inner products carry ~1e-7 * (G_a+G_b)/2, amplified by Rg^2/rmsd in the final cancellation
            # (SURVEY.md Appendix C): 1e-4 nm here.  Before the certificate these rows were off by up to 2 nm.
            if idx is None:
                assert_close(D[i][m], truth[m], atol=1e-4, what=f"degenerate all-pairs ({what}) row {i}")
            else:
                assert_close_msd(D[i][m], truth[m], X[:, idx], what=f"degenerate all-pairs ({what}) row {i}")


def test_one_vs_many_degenerate_geometries(mdb, oracle_mod):
    """md.rmsd / superpose on inputs whose QCP quartic has a double largest root (atoms on a line, two-atom
    selections): qcp_solve's certificate + closed form.  The reference's closed-form root from float32 coefficients
    is off by up to 2.5e-2 nm here, so the yardstick is float64 truth (three-way check)."""
    O = oracle_mod
    rng = np.random.default_rng(17)
    F = 3000
    line = rng.standard_normal((30, 1)) * np.array([[1.0, 0.0, 0.0]]) + 0.05 * rng.standard_normal((F, 30, 3))
    line = np.einsum("fnk,fkl->fnl", line, O.random_rotations(F, rng)).astype(np.float32)
    for what, X, idx in (("line", line, None), ("two of 20 atoms", O.synth_md(F, 20, seed=6, rg=0.5, sigma=0.1), [3, 11]),
                         ("two atoms", O.synth_iid(F, 2, seed=7), None)):
        t = mdb.Trajectory(X.copy())
        got = mdb.rmsd(t, t, 5, atom_indices=idx)
        truth = O.truth_rmsd(X, X, 5, atom_indices=idx)
        m = np.arange(F) != 5
        # float32 partial sums: ~1e-7 * (G_a+G_b)/2, amplified by Rg^2/rmsd; the reference
        # itself is off by up to 2.5e-2 nm on these inputs
        if what == "line":
            assert_close(got[m], truth[m], atol=1e-4, what=f"degenerate one-vs-many ({what})")
        else:
            assert_close_msd(got[m], truth[m], X if idx is None else X[:, idx], what=f"degenerate one-vs-many ({what})")
        ref = O.rmsd(X, X, 5, atom_indices=idx, impl="reference" if O.ref_available() else "port")
        assert np.abs(got[m] - truth[m]).max() <= max(1e-4, 1.5 * np.abs(ref[m] - truth[m]).max()), \
            f"degenerate one-vs-many ({what}): farther from float64 truth than the reference is"
        dt = mdb.DeviceTrajectory.from_host(X)
        dt.superpose(dt, 5, atom_indices=idx)   # rotations are undetermined here; the result must stay finite and rigid
        Y = dt.xyz
        assert np.isfinite(Y).all()
        d0 = np.linalg.norm(X[:, 0] - X[:, 1], axis=1)
        d1 = np.linalg.norm(Y[:, 0] - Y[:, 1], axis=1)
        assert np.abs(d0 - d1).max() < 1e-5


# ------------------------------------------------------------------ "next" rows: rmsf, align/displace
@pytest.mark.parametrize("kind,F,N,seed", [("md", 40, 303, 14), ("iid", 64, 100, 11)])
def test_rmsf_and_align_displace_golden(mdb, golden, oracle_mod, kind, F, N, seed):
    """md.rmsf (_rmsd.pyx:247-484) and getMultipleAlignDisplaceRMSDs_atom_major (:679-759) vs the real reference.
    The reference's own tests use decimal=3 for rmsf (tests/test_rmsd.py:251); we hold 2e-5."""
    O = oracle_mod
    X = gen(O, kind, F, N, seed)
    key = f"{kind}_{F}x{N}_s{seed}"
    idx = np.arange(0, N, 3)
    t = mdb.Trajectory(X.copy())
    assert_close(mdb.rmsf(t, t, 1), golden[key + "_rmsf_f1"], atol=2e-5, what="rmsf")
    # index list + reference: upstream passes a non-contiguous copy to rot_atom_major (_rmsd.pyx:408,415) and
    # returns frame-mixed values (golden key *_rmsf_f2_idx3 documents them); we return the all-atom result of the
    # sliced trajectory instead (see mdtraj_b200/rmsf_impl.py)
    sliced = mdb.Trajectory(X[:, idx].copy())
    assert_close(mdb.rmsf(t, t, 2, atom_indices=idx), mdb.rmsf(sliced, sliced, 2), atol=2e-6, what="rmsf idx == sliced")
    assert np.abs(golden[key + "_rmsf_f2_idx3"] - mdb.rmsf(sliced, sliced, 2)).max() > 1e-2  # upstream bug still there
    raw = X[:, idx].astype(np.float64)
    assert_close(mdb.rmsf(t, None, atom_indices=idx), np.sqrt(3 * np.mean((raw - raw.mean(0)) ** 2, axis=(0, 2))),
                 atol=2e-5, what="rmsf prealigned idx (tests/test_rmsd.py:255-267)")
    assert_close(mdb.rmsf(t, None), golden[key + "_rmsf_prealigned"], atol=2e-5, what="rmsf prealigned")
    c = mdb.Trajectory(X.copy()); c.center_coordinates()
    assert_close(mdb.rmsf(c, c, 0, precentered=True), golden[key + "_rmsf_f0_precentered"], atol=2e-5,
                 what="rmsf precentered")
    dt = mdb.DeviceTrajectory.from_host(X)
    assert_close(mdb.rmsf(dt, dt, 1), golden[key + "_rmsf_f1"], atol=2e-5, what="rmsf device path")
    # independent check of the definition (tests/test_rmsd.py:244-252)
    # (rmsf removes each frame's own centroid on the all-atom path, then rotates onto the reference frame)
    Y = np.einsum("fni,fij->fnj", X.astype(np.float64) - X.astype(np.float64).mean(1, keepdims=True), O.truth_superpose(X, X, 1)[1])
    want = np.sqrt(3 * np.mean((Y - Y.mean(0)) ** 2, axis=(0, 2)))
    assert_close(mdb.rmsf(t, t, 1), want, atol=2e-5, what="rmsf vs float64 definition")
    with pytest.raises(ValueError, match="Mode must be one of"):
        mdb.rmsf(t, t, 0, mode="bogus")
    with pytest.raises(ValueError, match="Cannot calculate RMSF of frame"):
        mdb.rmsf(t, t, F)
    # align on the first 60 atoms, measure on the last 40
    A = np.zeros((F, 60, 3), np.float32); D = np.zeros((F, 40, 3), np.float32)
    A[:] = X[:, :60]; D[:] = X[:, N - 40:]
    A -= A.mean(1, keepdims=True)
    gA = np.einsum("ijk,ijk->i", A, A).astype(np.float32)
    r, rot = mdb.getMultipleAlignDisplaceRMSDs_atom_major(A, A, gA, gA, D, D, 60, 40, 3)
    m = np.arange(F) != 3
    assert_close(r[m], golden[key + "_aligndispl_rmsd"][m], atol=2e-5, what="align/displace rmsd")
    if kind == "md":
        assert np.abs(rot[m] - golden[key + "_aligndispl_rot"][m]).max() < 1e-4
    assert r.shape == (F,) and rot.shape == (F, 3, 3)


def test_rmsf_residue_mode(mdb, oracle_mod):
    class _El:  # minimal topology duck type (mdtraj.Topology API used at _rmsd.pyx:467-471)
        def __init__(self, m): self.mass = m
    class _Res:
        def __init__(self, i): self.index = i
    class _Atom:
        def __init__(self, r, m): self.residue = _Res(r); self.element = _El(m)
    class _Top:
        def __init__(self, n): self.atoms_ = [_Atom(i // 4, 12.0 if i % 2 else 1.0) for i in range(n)]; self.n_residues = (n + 3) // 4
        def atom(self, i): return self.atoms_[i]
    X = oracle_mod.synth_md(30, 22, seed=8, rg=0.4, sigma=0.05)
    t = mdb.Trajectory(X.copy(), topology=_Top(22))
    per_atom = mdb.rmsf(t, t, 0)
    per_res = mdb.rmsf(t, t, 0, mode="residue")
    masses = np.array([12.0 if i % 2 else 1.0 for i in range(22)])
    for r in range(6):
        sel = np.arange(22)[np.arange(22) // 4 == r]
        want = np.sqrt(np.sum(per_atom[sel] ** 2 * masses[sel]) / masses[sel].sum())
        assert abs(per_res[r] - want) < 1e-6
    sub = mdb.rmsf(t, t, 0, atom_indices=[0, 1, 2, 3], mode="residue")
    assert sub[0] > 0 and np.all(sub[1:] == -1.0)


# ------------------------------------------------------------------ frame-resident kernels at scale
@pytest.mark.parametrize("F,N,stride", [(6000, 1000, 4), (1500, 5000, 5), (20000, 300, 1), (40000, 22, 1), (700, 8000, 7),
                                        (30001, 100, 1), (25000, 50, 3), (4001, 1001, 4), (3000, 2000, 1), (2500, 1400, 4),
                                        (2000, 1600, 1)])
def test_superpose_and_center_properties_at_scale(mdb, F, N, stride):
    """Every geometry of frame_resident_kernel (8 single-slot groups with 1-64 frames per slot, 2-16 lanes per frame for
    small frames, 3 single-buffer groups for 60 KB frames, partial last slots, the two-pass fallback for frames that do
    not fit) and many frames per CTA, checked through
    size-independent properties: superpose-then-plain-RMSD == QCP RMSD (tests/test_rmsd.py:98-108), a second superpose
    is the identity, centring leaves zero means and traces == sum |x|^2, repeated launches are bit-identical."""
    import torch
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=11)
    # MD-like: one base structure + 0.1 nm noise + a per-frame offset (well-conditioned rotations)
    g = torch.Generator(device=dt.device); g.manual_seed(5)
    base = torch.randn((N, 3), generator=g, device=dt.device)
    dt.xyz_dev[:, :N] = base[None] + 0.1 * dt.xyz_dev[:, :N] + (torch.rand((F, 1, 3), generator=g, device=dt.device) * 6 - 3)
    idx = None if stride == 1 else np.arange(0, N, stride)
    ref = mdb.DeviceTrajectory(dt.xyz_dev[:1].clone(), N)
    qcp = mdb.rmsd_device(dt, ref, 0, atom_indices=idx, as_numpy=False)
    a = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    b = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    a.superpose(ref, 0, atom_indices=idx)
    b.superpose(ref, 0, atom_indices=idx)
    assert torch.equal(a.xyz_dev, b.xyz_dev)
    assert torch.equal(a.xyz_dev[:, N:], torch.zeros_like(a.xyz_dev[:, N:]))  # padding atoms stay zero
    # frame 0 IS the reference (true RMSD 0, where float32 sums leave ~1e-4 nm of sqrt-amplified noise): skip it
    assert (a.last_superpose_rmsd - qcp)[1:].abs().max().item() < 1e-5
    plain = mdb.rmsd_device(a, ref, 0, atom_indices=idx, superpose=False, as_numpy=False)
    assert (plain - qcp)[1:].abs().max().item() < 2e-5
    before = a.xyz_dev.clone()
    a.superpose(ref, 0, atom_indices=idx)
    assert (a.xyz_dev - before).abs().max().item() < 2e-5
    # centring
    c = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    c.center_coordinates()
    x = c.xyz_dev[:, :N].double()
    assert x.mean(1).abs().max().item() < 2e-6
    assert ((x * x).sum((1, 2)) - c._rmsd_traces.double()).abs().max().item() <= 2e-6 * float((x * x).sum((1, 2)).max())
    pre = mdb.rmsd_device(c, c, 0, precentered=True, as_numpy=False)
    fly = mdb.rmsd_device(dt, dt, 0, as_numpy=False)
    assert (pre - fly)[1:].abs().max().item() < 1e-5


def test_allpairs_block_api(mdb, oracle_mod, ap_path):
    """b200rmsd_allpairs_block_dev: rectangular blocks, transposed copies and diagonal squares on both kernels."""
    import torch
    from mdtraj_b200 import allpairs as AP
    X = oracle_mod.synth_md(700, 64, seed=33, rg=0.8, sigma=0.1)
    dt = mdb.DeviceTrajectory.from_host(X)
    for path in ("tc", "simt"):
        ap_path(path)
        full = unmirrored_matrix(dt)
        prep = AP.prepare(dt)
        out = torch.zeros((190, 700), dtype=torch.float32, device=dt.device)
        out_t = torch.zeros((251, 190), dtype=torch.float32, device=dt.device)
        AP.block(prep, 123, 313, 407, 658, out, out_t)
        assert torch.equal(out[:, 407:658], full[123:313, 407:658])
        assert torch.equal(out_t, full[123:313, 407:658].t())
        assert torch.count_nonzero(out[:, :407]) == 0 and torch.count_nonzero(out[:, 658:]) == 0
        sq = torch.zeros((190, 700), dtype=torch.float32, device=dt.device)
        AP.block(prep, 123, 313, 123, 313, sq)  # diagonal square: each pair once, mirrored
        blk = sq[:, 123:313]
        assert torch.equal(blk, blk.t()) and torch.count_nonzero(torch.diagonal(blk)) == 0
        assert (blk - full[123:313, 123:313]).abs().max().item() < 2e-6


@pytest.mark.parametrize("path,F,N,basins", [("tc", 700, 300, 1), ("tc", 900, 130, 3), ("simt", 300, 64, 1)])
def test_allpairs_block_rotations(mdb, oracle_mod, ap_path, path, F, N, basins):
    """north_star: the fused epilogue turns each 3x3 block "into an RMSD, plus a rotation matrix for superpose"
    (b200rmsd_allpairs_block_rot_dev).  Row i of the rotations == what the one-vs-many path (and the reference's
    md.rmsd / superpose, theobald_rmsd.cpp:280-334) finds for every frame against reference frame i: 1e-5 per element
    against the float64 truth on MD-like frames, in one and in three basins (frames aligned onto different references
    in the prepare step), on both kernels; the RMSDs of the same call are the matrix's own."""
    import torch
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    if basins == 1:
        X = O.synth_md(F, N, seed=91, rg=1.0, sigma=0.1)
    else:
        X = O.synth_md_basins(F, N, basins, seed=92, rg=1.0, sigma=0.1, separation=1.4, interleave=True)[0]
    X = (X + np.float32(1.0)).astype(np.float32)           # off-centre frames: the rotations must not care
    ap_path(path)
    dt = mdb.DeviceTrajectory.from_host(X)
    prep = AP.prepare(dt)
    r0, r1, c0, c1 = 37, 37 + 83, 5, F - 11                # windows that start inside tiles
    D, U = AP.block_rotations(prep, r0, r1, c0, c1)
    full = unmirrored_matrix(dt)
    assert torch.equal(D, full[r0:r1, c0:c1])
    U = U.cpu().numpy()
    assert U.shape == (r1 - r0, c1 - c0, 3, 3)
    assert np.abs(np.linalg.det(U.astype(np.float64)) - 1).max() < 1e-5
    impl = "reference" if O.ref_available() else "port"
    for i in (r0, r0 + 41, r1 - 1):
        _, truth_R = O.truth_superpose(X, X, i, None)      # frame j onto frame i, float64 Kabsch
        got = U[i - r0]
        err = np.abs(got - truth_R[c0:c1]).max()
        assert err < 1e-5, f"row {i}: rotation elements {err:.2e} from the float64 truth"
        _, want_R = O.superpose(X, X, i, None, impl=impl, return_rot=True)
        assert_three_way(got, want_R[c0:c1], truth_R[c0:c1], f"all-pairs rotations, row {i}")
        # the same rotations from the one-vs-many kernel of this library
        d2 = mdb.DeviceTrajectory.from_host(X)
        _, R1 = d2.superpose(d2, i, return_rotations=True)
        assert np.abs(got - R1.cpu().numpy()[c0:c1]).max() < 1e-5
        if c0 <= i < c1:
            assert np.array_equal(got[i - c0], np.eye(3, dtype=np.float32))
    # applying row i's rotations superposes the frames: plain RMSD after == QCP RMSD
    i = r0 + 41
    Xc = X.astype(np.float64) - X.astype(np.float64).mean(1, keepdims=True)
    moved = np.einsum("jna,jab->jnb", Xc[c0:c1], U[i - r0].astype(np.float64))
    plain = np.sqrt(((moved - Xc[i][None]) ** 2).sum((1, 2)) / N)
    assert np.abs(plain - D[i - r0].cpu().numpy().astype(np.float64)).max() < 1e-5


# ------------------------------------------------------------------ md.lprmsd (last "next" row)
def _lprmsd_case(gold, name):
    from test_oracle import lprmsd_case
    return lprmsd_case(gold, name)


def test_lprmsd_matches_real_reference_goldens(mdb, oracle_mod):
    """Every golden case of the real md.lprmsd (tests/golden/make_golden_lprmsd.py): distances within 1e-5 nm (or, for the
    float32 reference's own noise, at least as close to the float64 Kabsch RMSD of the same matching), the matching equal
    to the oracle's, superposed coordinates within 1e-5 of what the reference leaves in target.xyz."""
    O = oracle_mod
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "lprmsd_outputs.npz"))
    for name in map(str, gold["cases"]):
        X, ref, idx, groups = _lprmsd_case(gold, name)
        t, r = mdb.Trajectory(X.copy()), mdb.Trajectory(ref.copy())
        d, mapping = mdb.lprmsd(t, r, 0, atom_indices=idx, permute_groups=groups, return_mapping=True)
        assert np.array_equal(t.xyz, X), "lprmsd without superpose must not touch the target"
        _, want_map = O.lprmsd(X, ref, 0, idx, groups, return_mapping=True)
        assert np.array_equal(mapping, want_map), name
        sel = np.arange(X.shape[1]) if idx is None else np.unique(idx)
        truth = O.truth_lprmsd_given_mapping(X, ref[0], sel, want_map)
        assert_three_way(d, gold[name + "_lprmsd"], truth, f"lprmsd {name}")
        d2 = mdb.lprmsd(t, r, 0, atom_indices=idx, permute_groups=groups, superpose=True)
        assert_three_way(d2, gold[name + "_lprmsd_superpose"], truth, f"lprmsd superpose {name}")
        assert np.abs(t.xyz - gold[name + "_xyz_superposed"]).max() < 1e-5, name
        # a DeviceTrajectory goes the same way and is modified in place
        dt = mdb.DeviceTrajectory.from_host(X)
        d3 = mdb.lprmsd(dt, r, 0, atom_indices=idx, permute_groups=groups, superpose=True)
        assert np.array_equal(d3, d2) and np.array_equal(dt.xyz, t.xyz)


def test_lprmsd_reference_test_semantics(mdb):
    """/root/reference/tests/test_lprmsd.py:25-79 restated on seeded data: identical frames, a pure relabelling, a pure
    rotation, both; and permute_groups=[[]] == md.rmsd (test_lprmsd_4, :110-120)."""
    rng = np.random.RandomState(0)
    ref = rng.randn(1, 10, 3).astype(np.float32)
    T = mdb.Trajectory
    assert mdb.lprmsd(T(ref.copy()), T(ref.copy()))[0] < 1e-3
    assert mdb.lprmsd(T(ref[:, rng.permutation(10)].copy()), T(ref.copy()))[0] < 1e-3
    ref = rng.randn(1, 50, 3).astype(np.float32)
    from oracle import oracle as O
    rot = O.random_rotations(1, np.random.default_rng(3))[0]
    new = ref.dot(rot).astype(np.float32)
    assert mdb.lprmsd(T(new.copy()), T(ref.copy()), permute_groups=[[]])[0] < 1e-2
    mapping = np.concatenate((rng.permutation(10), 10 + np.arange(40)))
    new = ref[:, mapping].dot(rot).astype(np.float32)
    assert mdb.lprmsd(T(new.copy()), T(ref.copy()), permute_groups=[np.arange(10)])[0] < 1e-2
    X = (ref + 0.05 * rng.randn(20, 50, 3)).astype(np.float32)
    idx = rng.permutation(50)[:45]
    got = mdb.lprmsd(T(X.copy()), T(ref.copy()), atom_indices=idx, permute_groups=[[]])
    want = mdb.rmsd(T(X.copy()), T(ref.copy()), atom_indices=np.unique(idx))
    assert np.abs(got - want).max() < 1e-5


def test_lprmsd_edge_cases(mdb, oracle_mod):
    """Unsorted / repeated atom_indices and group members (the reference runs np.unique over both), one-atom groups,
    a selection that is one atom, zero frames, a DeviceTrajectory as the reference, frame != 0."""
    O = oracle_mod
    rng = np.random.default_rng(11)
    N = 37
    ref = rng.standard_normal((4, N, 3)).astype(np.float32)
    X = (ref[2][None] + 0.05 * rng.standard_normal((9, N, 3))).astype(np.float32)
    idx = [30, 2, 2, 5, 7, 9, 11, 12, 30, 20, 21, 22, 3]
    groups = [[9, 5, 5, 7], [22], [20, 21]]
    X[:, [5, 7, 9]] = X[:, [9, 5, 7]]
    X[:, [20, 21]] = X[:, [21, 20]]
    got, gmap = mdb.lprmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(ref.copy()), 2, atom_indices=idx, permute_groups=groups,
                           return_mapping=True)
    want, wmap = O.lprmsd(X, ref, 2, idx, groups, impl="reference" if O.ref_available() else "port", return_mapping=True)
    assert np.array_equal(gmap, wmap)
    sel = np.unique(idx)
    assert_three_way(got, want, O.truth_lprmsd_given_mapping(X, ref[2], sel, wmap), "lprmsd edge cases")
    assert got.max() < 0.15, "the relabelling must have been undone (noise 0.05 nm per coordinate: ~0.09 nm)"
    # the reference held on the device; one selected atom (RMSD of a point onto a point is 0); zero frames
    again = mdb.lprmsd(mdb.Trajectory(X.copy()), mdb.DeviceTrajectory.from_host(ref), 2, atom_indices=idx, permute_groups=groups)
    assert np.array_equal(again, got)
    one = mdb.lprmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(ref.copy()), 1, atom_indices=[4])
    assert one.shape == (9,) and np.all(one < 1e-6)
    empty = mdb.lprmsd(mdb.Trajectory(X[:0].copy()), mdb.Trajectory(ref.copy()), 0)
    assert empty.shape == (0,) and empty.dtype == np.float32


def test_lprmsd_water_box_vs_oracle(mdb, oracle_mod):
    """2000 frames of 60 distinguishable atoms + two groups of 120 exchangeable ones, every frame relabelled at random and
    moved rigidly: the matching undoes the relabelling, distances agree with the oracle three ways on a sample of
    frames; a long group (600) runs through the same kernel."""
    O = oracle_mod
    rng = np.random.default_rng(7)
    F, N = 2000, 300
    groups = [np.arange(60, 180), np.arange(180, 300)]
    ref = (rng.standard_normal((1, N, 3)) * 1.2).astype(np.float32)
    X = np.repeat(ref, F, 0) + 0.02 * rng.standard_normal((F, N, 3))
    perms = np.tile(np.arange(N), (F, 1))
    for f in range(F):
        for g in groups:
            perms[f, g] = rng.permutation(g)
    X = np.take_along_axis(X, perms[:, :, None], 1)            # target atom k is the reference's atom perms[f, k]
    X = (np.einsum("fni,fij->fnj", X, O.random_rotations(F, rng)) + rng.uniform(-3, 3, (F, 1, 3))).astype(np.float32)
    d, mapping = mdb.lprmsd(mdb.Trajectory(X), mdb.Trajectory(ref), permute_groups=groups, return_mapping=True)
    inv = np.empty_like(perms)
    np.put_along_axis(inv, perms, np.tile(np.arange(N), (F, 1)), 1)   # reference atom i sits at target position inv[f, i]
    # the optimum undoes the relabelling except where two atoms of a group happen to sit within the noise of each other
    # (a random cloud has a few such pairs); the exact comparison is the one with the oracle's solver below
    assert (mapping == inv).mean() > 0.999, (mapping == inv).mean()
    assert all(np.array_equal(np.sort(m), np.arange(N)) for m in mapping[::50]), "every matching is a permutation"
    assert 0.025 < d.min() and d.max() < 0.045                         # ~ sigma * sqrt(3)
    sample = np.arange(0, F, 97)
    want, want_map = O.lprmsd(X[sample], ref, 0, None, groups, impl="reference" if O.ref_available() else "port",
                              return_mapping=True)
    assert np.array_equal(mapping[sample], want_map)
    truth = O.truth_lprmsd_given_mapping(X[sample], ref[0], np.arange(N), want_map)
    assert_three_way(d[sample], want, truth, "lprmsd water box")
    # one long group
    N2 = 640
    ref2 = (rng.standard_normal((1, N2, 3)) * 1.5).astype(np.float32)
    X2 = np.repeat(ref2, 6, 0) + 0.01 * rng.standard_normal((6, N2, 3))
    g2 = np.arange(40, N2)
    for f in range(6):
        X2[f, g2] = X2[f, rng.permutation(g2)]
    X2 = X2.astype(np.float32)
    d2, m2 = mdb.lprmsd(mdb.Trajectory(X2), mdb.Trajectory(ref2), permute_groups=[g2], return_mapping=True)
    w2, wm2 = O.lprmsd(X2, ref2, 0, None, [g2], return_mapping=True)
    assert np.array_equal(m2, wm2)
    assert_three_way(d2, w2, O.truth_lprmsd_given_mapping(X2, ref2[0], np.arange(N2), wm2), "lprmsd long group")


# ------------------------------------------------------------------ consumers of the matrix
def test_clustering_consumers_ala2_known_answers(mdb, golden, ala2):
    """examples/centroids.ipynb:117 -> centroid index 83 on the heavy atoms; examples/clustering.ipynb:101 -> squareform."""
    heavy = golden["ala2_heavy_idx"]
    t = mdb.Trajectory(ala2.copy())
    assert mdb.centroid_index(t, atom_indices=heavy) == int(golden["ala2_centroid_index"]) == 83
    Dh = golden["ala2_allpairs_heavy"].astype(np.float64)
    want = np.exp(-1.0 * Dh / Dh.std()).sum(axis=1)
    scores, std = mdb.similarity_scores(t, atom_indices=heavy)
    assert abs(std - Dh.std()) < 1e-6
    assert_close(scores, want, atol=0, rtol=2e-4, what="similarity scores")  # 1e-5 nm on d ~ 1e-4 relative on exp(-d/std)
    D = golden["ala2_allpairs"]
    iu = np.triu_indices(100, k=1)
    cond = mdb.rmsd_condensed(t)
    assert cond.dtype == np.float64 and cond.shape == (100 * 99 // 2,)
    assert_close(cond, D[iu], what="condensed distances == squareform(D)")
    assert abs(cond.max() - 0.188493) < 5e-6  # clustering.ipynb:73


@pytest.mark.parametrize("F,N", [(700, 50), (300, 22)])
def test_clustering_consumers_vs_numpy(mdb, F, N):
    """Every reduction against numpy on the same device-computed matrix, on both all-pairs kernels (F >= 512: tcgen05),
    with and without row blocking."""
    import torch
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=21)
    g = torch.Generator(device=dt.device); g.manual_seed(3)
    base = torch.randn((N, 3), generator=g, device=dt.device)
    dt.xyz_dev[:, :N] = base[None] + 0.3 * dt.xyz_dev[:, :N]
    D = mdb.rmsd_matrix(dt).astype(np.float64)
    cond = mdb.rmsd_condensed(dt, dtype=np.float32)
    assert np.array_equal(cond, D[np.triu_indices(F, k=1)].astype(np.float32))
    want = np.exp(-2.0 * D / D.std()).sum(axis=1)
    for max_bytes in (1 << 40, 4 * F * 97):  # one block; many ragged row blocks
        scores, std = mdb.similarity_scores(dt, beta=2.0, max_matrix_bytes=max_bytes)
        assert abs(std - D.std()) < 1e-9 * max(1.0, D.std()) + 1e-7
        assert_close(scores, want, atol=0, rtol=1e-5, what="similarity scores")
        assert mdb.centroid_index(dt, beta=2.0, max_matrix_bytes=max_bytes) == int(want.argmax())
    # nearest leader: leaders = frames 5, 77, 200 (+ a duplicate to pin first-minimum semantics)
    lead_idx = [5, 77, 200, 77]
    leaders = mdb.DeviceTrajectory(dt.xyz_dev[lead_idx].clone(), N)
    for max_bytes in (1 << 40, 4 * len(lead_idx) * 33):
        labels, dist = mdb.assign_to_leaders(dt, leaders, max_matrix_bytes=max_bytes)
        block = D[:, lead_idx]
        assert labels.dtype == np.int32 and labels.shape == (F,)
        # a frame against a COPY of itself has true RMSD 0, where the float32 sums leave sqrt-amplified noise (~1e-3 nm on
        # the 3xTF32 path) while D's diagonal is exactly 0 by the same-frame shortcut: leave those three rows out
        m = ~np.isin(np.arange(F), lead_idx)
        assert_close(dist[m], block.min(axis=1)[m], atol=2e-6, what="distance to the nearest leader")
        assert np.all(dist[~m] < 5e-3)
        # ties between near-equal leaders may resolve differently at the 1e-6 level: compare through the distances
        assert np.all(np.abs(block[np.arange(F), labels] - block.min(axis=1)) < 2e-6)
        assert labels[77] == 1 and labels[5] == 0 and labels[200] == 2  # the duplicate never wins over the first minimum
    one = mdb.Trajectory(dt.xyz[:50].copy())
    want_lab = np.array([np.argmin(mdb.rmsd(leaders, mdb.DeviceTrajectory.from_trajectory(one), i)) for i in range(50)])
    got = mdb.assign_to_leaders(one, leaders)[0]
    agree = got == want_lab
    assert agree.mean() > 0.95 and np.all(np.abs(D[:50][:, lead_idx][np.arange(50), got] - D[:50][:, lead_idx].min(1)) < 2e-6)


def test_rmsd_matrix_into_host_buffer(mdb):
    """rmsd_matrix(out=...) produces the matrix in row blocks whose device->host copies overlap the next block's
    compute: same matrix as the single-call path (entries of the mirrored half agree to rounding), page-locked or not."""
    import torch
    F, N = 900, 64
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=8)
    want = mdb.rmsd_matrix(dt)
    for pinned in (True, False):
        host = torch.empty((F, F), dtype=torch.float32, pin_memory=pinned)
        host.fill_(-1.0)
        got = mdb.rmsd_matrix(dt, out=host.numpy(), row_block=128)
        assert got is not None and got.shape == (F, F) and np.shares_memory(got, host.numpy())
        assert_close(got, want, atol=2e-6, what="blocked matrix vs single call")
        assert np.all(np.diag(got) == 0.0)
    with pytest.raises(ValueError):
        mdb.rmsd_matrix(dt, out=np.empty((F, F), dtype=np.float64))


# ------------------------------------------------------------------ host pipeline: pageable staging, lanes, several devices
@pytest.fixture
def small_chunks(mdb):
    """1 MB chunks: a few thousand frames already exercise the three-lane pipeline, the finalizer thread and the memcpy
    pool (default 64 MB)."""
    mdb.set_host_pipeline(chunk_mb=1, staged_chunk_mb=1)
    yield
    mdb.set_host_pipeline(chunk_mb=64, staged_chunk_mb=16)


@pytest.mark.parametrize("N", [100, 22, 301])
def test_host_pipeline_pageable_pinned_and_chunking_agree(mdb, oracle_mod, small_chunks, N):
    """md.rmsd / superpose / centring on host arrays: pageable input (staged through page-locked lanes by the memcpy
    pool), page-locked input (DMA'd directly) and a single-chunk run give bit-identical results, for frame sizes with
    and without padding atoms (N % 4 != 0 takes the row-wise staging copy)."""
    import torch
    O = oracle_mod
    F = 9000 if N > 30 else 40000
    X = O.synth_md(F, N, seed=70 + N, rg=0.4 if N < 30 else 1.5)   # a 22-atom molecule is 0.3-0.4 nm across, not 1.5
    ref = mdb.Trajectory(X[:3].copy())
    idx = np.arange(0, N, 3)
    pinned = torch.empty((F, N, 3), dtype=torch.float32).pin_memory()
    pinned.copy_(torch.from_numpy(X))
    got = {}
    for name, arr in (("pageable", X.copy()), ("pinned", pinned.numpy())):
        t = mdb.Trajectory.__new__(mdb.Trajectory)
        t.topology, t._xyz, t._rmsd_traces = None, arr, None
        got[name] = (mdb.rmsd(t, ref, 1), mdb.rmsd(t, ref, 2, atom_indices=idx))
    mdb.set_host_pipeline(chunk_mb=1024, staged_chunk_mb=1024)
    one = mdb.rmsd(mdb.Trajectory(X.copy()), ref, 1)
    mdb.set_host_pipeline(chunk_mb=1, staged_chunk_mb=1)
    assert np.array_equal(got["pageable"][0], got["pinned"][0]) and np.array_equal(got["pageable"][0], one)
    # pageable memory goes up by streamed staging (pieces through the pool threads' slots) by default: whole-chunk
    # staging and other piece sizes move the same bytes
    try:
        for kb in (-1, 16, 2048):
            mdb.set_host_pipeline(stage_piece_kb=kb)
            t = mdb.Trajectory(X.copy())
            assert np.array_equal(mdb.rmsd(t, ref, 1), one), f"stage_piece_kb={kb}"
            t.superpose(ref, 0, atom_indices=idx)
            if kb == -1:
                whole = t.xyz.copy()
            else:
                assert np.array_equal(t.xyz, whole), f"stage_piece_kb={kb}"
    finally:
        mdb.set_host_pipeline(stage_piece_kb=0)
    assert np.array_equal(got["pageable"][1], got["pinned"][1])
    want = O.rmsd(X, X[:3], 1, impl="reference" if O.ref_available() else "port")
    m = np.arange(F) != 1   # the reference frame against itself: noise floor here, exactly 0 in the reference (same memory)
    assert_three_way(one[m], want[m], O.truth_rmsd_batch(X, X[1])[m], "host pipeline vs oracle")
    # in-place operations: superposed coordinates and traces come back through the finalizer
    a, b = mdb.Trajectory(X.copy()), mdb.Trajectory(pinned.numpy().copy())
    tp = mdb.Trajectory.__new__(mdb.Trajectory)
    tp.topology, tp._xyz, tp._rmsd_traces = None, pinned.numpy(), None
    a.superpose(ref, 0, atom_indices=idx)
    tp.superpose(ref, 0, atom_indices=idx)
    mdb.set_host_pipeline(chunk_mb=1024, staged_chunk_mb=1024)
    b.superpose(ref, 0, atom_indices=idx)
    mdb.set_host_pipeline(chunk_mb=1, staged_chunk_mb=1)
    assert np.array_equal(a.xyz, b.xyz) and np.array_equal(a.xyz, tp.xyz)
    c = mdb.Trajectory(X.copy())
    c.center_coordinates()
    Xc = X.copy()
    tr = O.center_and_trace(Xc, "reference" if O.ref_available() else "port")
    assert np.abs(c.xyz - Xc).max() <= 2e-6 and np.allclose(c._rmsd_traces, tr, rtol=1e-6)


def test_host_pipeline_several_devices(mdb, oracle_mod, small_chunks):
    """b200rmsd_*_host_multi: chunks handed out dynamically to every visible device from one process; the result does not
    depend on which device took which chunk.  Needs >= 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    O = oracle_mod
    X = O.synth_md(30000, 100, seed=91)
    ref = mdb.Trajectory(X[:1].copy())
    try:
        mdb.set_devices([0])
        one = mdb.rmsd(mdb.Trajectory(X.copy()), ref, 0)
        s1 = mdb.Trajectory(X.copy()); s1.superpose(ref, 0)
        mdb.set_devices(list(range(torch.cuda.device_count())))
        many = mdb.rmsd(mdb.Trajectory(X.copy()), ref, 0)
        s2 = mdb.Trajectory(X.copy()); s2.superpose(ref, 0)
        # streamed staging with several devices: pool threads (and the devices' own host threads, which help) send pieces
        # for whichever device a piece belongs to
        mdb.set_host_pipeline(stage_piece_kb=64)
        streamed = [mdb.rmsd(mdb.Trajectory(X.copy()), ref, 0) for _ in range(3)]
    finally:
        mdb.set_host_pipeline(stage_piece_kb=0)
        mdb.set_devices(None)
    assert np.array_equal(one, many) and np.array_equal(s1.xyz, s2.xyz)
    assert all(np.array_equal(one, s) for s in streamed)


def test_host_pipeline_error_paths(mdb):
    """Bad device lists are rejected before anything is launched; a failed call leaves no transfer in flight."""
    from mdtraj_b200 import _capi
    L = _capi.lib()
    X = np.zeros((10, 8, 3), np.float32); out = np.zeros(10, np.float32)
    for devs in ([99], [0, 0], []):
        d = np.asarray(devs, np.int32)
        rc = L.b200rmsd_rmsd_host_multi(X.ctypes.data, 10, 8, X[0].ctypes.data, 8, None, None, 0, 1, 0, None, 0.0,
                                        out.ctypes.data, d.ctypes.data if len(devs) else None, len(devs))
        assert rc == _capi.EINVAL or rc == _capi.ENODEVICE, (devs, rc)
        assert L.b200rmsd_last_error()


# ---------------------------------------------------------------------------------------------------------------------
# legacy raw-array entry points of mdtraj._rmsd (/root/reference/mdtraj/rmsd/_rmsd.pyx:502-674) against the compiled
# reference loops; the only callers of b200rmsd_rotate_dev
# ---------------------------------------------------------------------------------------------------------------------
def _ref_superpose_atom_major(O, align_target_frame, g_target, align_mobile, g_mobile, displace):
    """The reference's superpose_atom_major loop (oracle/ref_loops.cpp, or the C port), in place on `displace`."""
    F, n_align, n_disp = align_mobile.shape[0], align_mobile.shape[1], displace.shape[1]
    rot = np.zeros((F, 9), dtype=np.float32)
    tgt = np.ascontiguousarray(align_target_frame, dtype=np.float32)
    if O.ref_available():
        O.ref_lib().refloops_superpose_atom_major(O._ptr(tgt), float(g_target), O._ptr(align_mobile), O._ptr(g_mobile), F,
                                                  n_align, O._ptr(displace), n_disp, 1, O._ptr(rot))
    else:
        O.port_lib().oracle_superpose_atom_major(O._ptr(tgt), float(g_target), O._ptr(align_mobile), O._ptr(g_mobile), F,
                                                 n_align, O._ptr(displace), n_disp, O._ptr(rot))
    return rot.reshape(F, 3, 3)


@pytest.mark.parametrize("F,N,Nd", [(300, 57, 57), (120, 1000, 1203), (40, 4100, 300)])
def test_legacy_raw_array_entry_points(mdb, oracle_mod, F, N, Nd):
    O = oracle_mod
    kind = "reference" if O.ref_available() else "port"
    X = O.synth_md(F, N, seed=900 + N)
    Y = O.synth_md(7, N, seed=901 + N)
    Xc, Yc = X.copy(), Y.copy()
    g, gy = O.center_and_trace(Xc, kind), O.center_and_trace(Yc, kind)
    m = np.arange(F) != 3

    # getMultipleRMSDs_atom_major: same array twice (frame against itself is exactly 0), and two different arrays
    got = mdb.getMultipleRMSDs_atom_major(Xc, Xc, g, g, 3)
    want = O.one_vs_many_centered(Xc, g, Xc[3], g[3], impl=kind)
    assert got.dtype == np.float32 and got.shape == (F,) and got[3] == 0.0
    assert_three_way(got[m], want[m], O.truth_rmsd_batch(X, X[3])[m], "getMultipleRMSDs_atom_major (same array)")
    got2 = mdb.getMultipleRMSDs_atom_major(Yc, Xc, gy, g, 5)
    want2 = O.one_vs_many_centered(Xc, g, Yc[5], gy[5], impl=kind)
    assert_three_way(got2, want2, O.truth_rmsd_batch(X, Y[5]), "getMultipleRMSDs_atom_major (two arrays)")
    with pytest.raises(ValueError):
        mdb.getMultipleRMSDs_atom_major(Yc, Xc, gy, g, 7)           # frame out of range, _rmsd.pyx:586-588
    with pytest.raises(ValueError):
        mdb.getMultipleRMSDs_atom_major(Yc[:, :-1].copy(), Xc, gy, g, 0)  # atom counts differ, :583-585

    # getMultipleRMSDs_axis_major: (F, 3, N) layout, the same numbers
    Xa, Ya = np.ascontiguousarray(Xc.transpose(0, 2, 1)), np.ascontiguousarray(Yc.transpose(0, 2, 1))
    got_a = mdb.getMultipleRMSDs_axis_major(Xa, Xa, g, g, 3)
    assert got_a[3] == 0.0 and np.abs(got_a - got).max() <= 1e-6
    got_a2 = mdb.getMultipleRMSDs_axis_major(Ya, Xa, gy, g, 5)
    assert np.abs(got_a2 - got2).max() <= 1e-6

    # superpose_atom_major: rotation from the centred align arrays, applied in place to a displace array of another
    # atom count (align != displace, trajectory.py:1152-1160)
    rng = np.random.default_rng(5 + N)
    D = (rng.standard_normal((F, Nd, 3)) * 2.0).astype(np.float32)
    D_gpu, D_ref = D.copy(), D.copy()
    assert mdb.superpose_atom_major(Xc, Xc, g, g, D_gpu, 3) is None
    R_ref = _ref_superpose_atom_major(O, Xc[3], g[3], Xc, g, D_ref)
    D_truth = np.empty_like(D, dtype=np.float64)
    for i in range(F):
        R, _ = O.truth_kabsch(X[i], X[3])
        D_truth[i] = D[i].astype(np.float64) @ R
    assert_three_way(D_gpu[m], D_ref[m], D_truth[m], "superpose_atom_major displaced coordinates", atol=2e-5)
    assert np.abs(D_gpu[m] - D_truth[m]).max() < 2e-5
    assert np.abs(R_ref[m] - np.stack([O.truth_kabsch(X[i], X[3])[0] for i in range(F)])[m]).max() < 1e-3  # oracle sanity
    with pytest.raises(ValueError):
        mdb.superpose_atom_major(Xc, Xc[:-1], g, g[:-1], D_gpu, 0)   # frame counts differ, :650-651


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json shapes against the oracle inside pytest (sizes the CPU side finishes in seconds)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("F,N,step,gen", [(10000, 1000, 1, "md"), (2000, 5000, 5, "md"), (800, 25000, 1, "md"),
                                          (800, 25000, 1, "chain"), (3000, 1000, 1, "chain")])
def test_baseline_shapes_vs_oracle_three_way(mdb, oracle_mod, F, N, step, gen):
    """C2 (N = 1000), C3's selection (every 5th of 5000 atoms) and C5 (N = 25000: atom segments + ovm_finish_kernel)
    against the compiled reference and the float64 truth; and the accuracy class this round's accumulation buys:
    at least 3x closer to the truth than the reference's float32 accumulation wherever that is measurable."""
    O = oracle_mod
    kind = "reference" if O.ref_available() else "port"
    X = O.synth_md(F, N, seed=40 + N // 1000) if gen == "md" else O.synth_md_chain(F, N, seed=41 + N // 1000)
    idx = None if step == 1 else np.arange(0, N, step)
    got = mdb.rmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(X.copy()), 0, atom_indices=idx)
    want = O.rmsd(X, X, 0, atom_indices=idx, impl=kind)
    truth = O.truth_rmsd_batch(X if idx is None else X[:, idx], X[0] if idx is None else X[0, idx])
    assert_three_way(got[1:], want[1:], truth[1:], f"N={N} step={step} {gen}")
    e_gpu, e_ref = np.abs(got[1:] - truth[1:]).max(), np.abs(want[1:] - truth[1:]).max()
    assert e_gpu < 1e-5, (e_gpu, e_ref)
    if e_ref > 6e-6:
        assert e_gpu < e_ref / 3, (e_gpu, e_ref)
    if idx is None:  # superposed coordinates of the same shape
        a = mdb.Trajectory(X[:200].copy()); a.superpose(mdb.Trajectory(X[:1].copy()), 0)
        sup_truth, _ = O.truth_superpose(X[:200], X[:1], 0)
        assert np.abs(a.xyz - sup_truth).max() < 1e-5


@pytest.mark.parametrize("N", [4800, 5000, 6200])
def test_superpose_all_atoms_large_frames(mdb, oracle_mod, N):
    """All atoms selected on ~5,000-atom frames: three frame buffers and a resident reference do not fit shared memory
    together, so superpose_pipe_kernel reads the reference from global memory (round 1 ran two passes over HBM here).
    Superposed coordinates and rotations against the float64 truth and the reference, through both entry points."""
    O = oracle_mod
    F = 600
    X = O.synth_md(F, N, seed=50 + N // 100)
    kind = "reference" if O.ref_available() else "port"
    want_xyz = O.superpose(X, X, 3, None, impl=kind)
    truth_xyz, truth_R = O.truth_superpose(X, X, 3, None)
    dt = mdb.DeviceTrajectory.from_host(X)
    _, R = dt.superpose(dt, 3, return_rotations=True)
    assert_three_way(dt.xyz, want_xyz, truth_xyz, f"superposed coordinates, all atoms, N={N}")
    assert np.abs(dt.xyz - truth_xyz).max() < 1e-5
    assert np.abs(R.cpu().numpy().reshape(F, 3, 3) - truth_R).max() < 1e-5
    t = mdb.Trajectory(X.copy())
    t.superpose(mdb.Trajectory(X.copy()), 3)
    assert np.array_equal(t.xyz, dt.xyz), "host entry point and device entry point run the same kernel"


def test_nccl_two_ranks_match_single_gpu():
    """The NCCL paths (mdtraj_b200.distributed: frame-sharded md.rmsd / superpose with all_gather, all-pairs with the
    broadcast + symmetric block exchange) against the single-GPU results, two ranks under torchrun
    (tools/multi_gpu_check.py holds the assertions).  Skipped on a one-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "multi_gpu_check.py")], capture_output=True, text=True, timeout=900,
                         cwd=root)
    assert res.returncode == 0, (res.stdout[-1500:], res.stderr[-3000:])
    assert "multi-GPU check ok" in res.stdout


def test_allpairs_cta_pair_geometry_is_bit_identical(mdb, oracle_mod):
    """The cta_group::2 geometry of the tcgen05 kernel (2-CTA clusters, one M = 256 MMA per pair of tiles, each CTA loading
    half of the B rows) against the single-CTA kernel: the same matrix bit for bit -- full symmetric matrix, an
    unsymmetric row block at an odd tile offset, and a block with a transposed copy (odd and even tile counts)."""
    import torch
    from mdtraj_b200 import allpairs as AP
    O = oracle_mod
    for F in (2090, 1640):
        X, _ = O.synth_md_basins(F, 150, 2, seed=3 + F, rg=1.0, sigma=0.1, separation=1.2)
        dt = mdb.DeviceTrajectory.from_host(X)
        prep = AP.prepare(dt)
        got = {}
        try:
            for pair in (False, True):
                AP.configure(cta_pair=pair)
                full = AP.rows(prep, 0, F).clone()
                blk = AP.rows(prep, 41, 41 + 333).clone()
                got[pair] = (full, blk)
        finally:
            AP.configure(cta_pair=True)
        assert torch.equal(got[False][0], got[True][0]) and torch.equal(got[False][1], got[True][1])
        assert torch.isfinite(got[True][0]).all()


# ---------------------------------------------------------------------------------------------------------------------
# bench.py: every record builds (a typo in a record's dict once cost the driver's run its C3 numbers)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("only", ["ovm", "superpose", "allpairs_20k", "ovm25k"])
def test_bench_records_build(only):
    import json
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--only", only, "--steps", "1", "--warmup", "3",
                          "--no-e2e", "--no-subs"], capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stderr[-2000:]
    rec = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "roofline"):
        assert key in rec, key
    assert rec["value"] > 0 and rec["roofline"]["frac"] > 0.3 and "workload" in rec["config"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rec["roofline"], key
