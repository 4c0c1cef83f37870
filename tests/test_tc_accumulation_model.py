"""CPU: the numpy model of the tensor-core all-pairs arithmetic (tests/tc_model.py) -- why the operands are stored as
differences from a common reference.  The model's one hardware assumption (fp32 accumulator truncated once per K=8
tcgen05.mma) is pinned to the bias measured on a B200 (profiles/r01_tc_accumulation_probe.json)."""
import json
import os

import numpy as np
import pytest

import tc_model as T

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def truth_rows(X, rows):
    from oracle import oracle as O
    return np.stack([O.truth_rmsd_batch(X, X[i]) for i in rows])


@pytest.fixture(scope="module")
def md_frames():
    from oracle import oracle as O
    return O.synth_md(160, 300, seed=3, rg=1.0, sigma=0.1)


def test_model_reproduces_measured_accumulator_bias(md_frames):
    probe = [json.loads(l) for l in open(os.path.join(ROOT, "profiles", "r01_tc_accumulation_probe.json")) if l.strip()]
    measured = next(p for p in probe if p["data"] == "md")
    ops = T.prepare_operands(md_frames, aligned=False)
    a = slice(0, 30)
    acc = T.tc_gemm(ops["a_hi"][a], ops["a_lo"][a], ops["b_hi"], ops["b_lo"])
    exact = (ops["a_hi"][a].astype(np.float64) + ops["a_lo"][a]) @ (ops["b_hi"].astype(np.float64) + ops["b_lo"]).T
    err = acc - exact
    big = np.abs(exact) > 0.5 * np.abs(exact).max()
    bias = float((err[big] * np.sign(exact[big])).mean())
    # measured on hardware: -5.3e-4 on values of magnitude ~324 (17 ulp); the model must land within 25 % of it
    assert measured["same_on_large_values"] < -3e-4
    assert bias < 0 and abs(bias - measured["same_on_large_values"]) < 0.25 * abs(measured["same_on_large_values"]), bias
    assert (err[big] * np.sign(exact[big])).max() <= 0.0   # truncation never overshoots


def test_difference_operands_remove_the_bias_from_the_rmsd(md_frames):
    rows = np.array([0, 57, 159])
    truth = truth_rows(md_frames, rows)
    m = np.ones_like(truth, bool); m[np.arange(len(rows)), rows] = False
    plain = T.prepare_operands(md_frames, aligned=False)
    diff = T.prepare_operands(md_frames, aligned=True)
    e_plain = np.abs(T.rmsd_rows(plain, rows) - truth)[m].max()
    e_diff = np.abs(T.rmsd_rows(diff, rows) - truth)[m].max()
    T.ROUND_BIAS = False   # the half-ulp nudge of G is meant for the truncating accumulator, not for an exact one
    try:
        e_diff_exact = np.abs(T.rmsd_rows(T.prepare_operands(md_frames, aligned=True), rows, exact=True) - truth)[m].max()
    finally:
        T.ROUND_BIAS = True
    # plain operands: the truncation bias costs more than the parity tolerance (GPU, before the change: 3.8e-5 nm)
    assert e_plain > 1e-5
    # difference operands: an order of magnitude inside it (GPU: 1.1-1.4e-6 nm); operand construction alone: 1e-7 class
    assert e_diff < 3e-6 and e_diff < e_plain / 8
    assert e_diff_exact < 5e-7


def test_dissimilar_frames_keep_plain_operands():
    """iid frames are farther from the reference than half their own size: B stays x', no unit vectors, same numbers as
    the plain layout (the per-frame switch of allpairs_tc144_prepare_kernel)."""
    from oracle import oracle as O
    X = O.synth_iid(60, 100, seed=9)
    ops = T.prepare_operands(X, aligned=True)
    k0 = (100 + 31) // 32 * 32
    assert len(ops["references"]) == 4               # three references in a row captured nothing: the traversal stops
    assert ops["near"].sum() == 4 and ops["b_hi"][:, k0:].sum() == 2 * 3 * 4   # only the references themselves are "near"
    rows = np.array([1, 30])
    truth = truth_rows(X, rows)
    m = np.ones_like(truth, bool); m[np.arange(2), rows] = False
    assert np.abs(T.rmsd_rows(ops, rows) - truth)[m].max() < 2e-6


def test_multi_reference_plan_on_three_basins():
    """The reference traversal of csrc/allpairs_refs.cu run through the model: with one reference (round 1's kernel) pairs
    inside a basin far from frame 0 keep the plain-operand error; greedy farthest-point references bring every basin to the
    1e-6 class."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    N, per = 300, 40
    b1 = rng.standard_normal((N, 3))
    bases = [b1, b1 + 0.8 * rng.standard_normal((N, 3)), b1 + 0.8 * rng.standard_normal((N, 3))]
    X = np.concatenate([b + 0.1 * rng.standard_normal((per, N, 3)) for b in bases])
    X = np.einsum("fni,fij->fnj", X, O.random_rotations(len(X), rng)).astype(np.float32)
    rows = np.array([1, per + 1, 2 * per + 1])
    truth = truth_rows(X, rows)
    m = np.ones_like(truth, bool); m[np.arange(3), rows] = False
    one = T.prepare_operands(X, aligned=True, max_refs=1)
    many = T.prepare_operands(X, aligned=True)
    assert one["references"] == [0]
    refs = np.asarray(many["references"])
    assert sorted(refs[:3] // per) == [0, 1, 2]               # one per basin first ...
    assert 3 <= len(refs) <= 5                                # ... then at most two refinement steps (thermal noise floor)
    assert (refs[many["owner"]] // per == np.arange(3 * per) // per).all() and many["near"].all()   # owned inside the basin
    e_one = np.abs(T.rmsd_rows(one, rows) - truth)
    e_many = np.abs(T.rmsd_rows(many, rows) - truth)
    inside2 = e_one[1, per:2 * per][np.arange(per) != 1].max()
    assert inside2 > 1e-5                                     # one reference: inside the second basin
    assert e_one[0, :per][np.arange(per) != 1].max() < 3e-6   # one reference: inside the reference's basin
    assert e_many[m].max() < 3e-6                             # traversal: everywhere
    assert many["a_hi"].shape[1] == 320 + 32 * ((len(refs) + 3) // 4)   # atom columns + one augmentation K block per 4 references


def test_drifting_trajectory_gets_references_along_the_path():
    """A trajectory that drifts away from frame 0 (rmsd 0.5 nm at the end) with neighbouring frames 0.04 nm apart: every
    frame is 'near' frame 0, yet with that single reference late neighbours are 2e-5 nm off (error ~ delta^2 / rmsd_ij);
    the traversal keeps adding references until the covering radius is under 0.25 nm."""
    from oracle import oracle as O
    rng = np.random.default_rng(1)
    N, F = 300, 240
    base = rng.standard_normal((N, 3))
    X = base[None] + np.cumsum(0.02 * rng.standard_normal((F, N, 3)), 0) + 0.01 * rng.standard_normal((F, N, 3))
    X = np.einsum("fni,fij->fnj", X, O.random_rotations(F, rng)).astype(np.float32)
    rows = np.array([5, 120, 230])
    truth = truth_rows(X, rows)
    m = np.ones_like(truth, bool); m[np.arange(3), rows] = False
    one = T.prepare_operands(X, aligned=True, max_refs=1)
    many = T.prepare_operands(X, aligned=True)
    assert one["near"].all() and np.abs(T.rmsd_rows(one, rows) - truth)[m].max() > 1.5e-5
    assert 3 <= len(many["references"]) <= 8
    assert np.abs(T.rmsd_rows(many, rows) - truth)[m].max() < 5e-6
