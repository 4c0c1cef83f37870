"""Golden vectors for md.lprmsd (tests/golden/lprmsd_outputs.npz), produced by the REAL reference.

Run in the build container only, with the reference installed by baseline/build_ref.sh:

    PYTHONPATH=baseline/_ref python tests/golden/make_golden_lprmsd.py

Inputs are small seeded synthetic cases (stored next to the outputs: they are a few KB); outputs are whatever
mdtraj.lprmsd / mdtraj._lprmsd._munkres return (mdtraj/rmsd/_lprmsd.pyx:71-221, :276-307).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import mdtraj as md  # the real reference  # noqa: E402
from mdtraj._lprmsd import _munkres  # noqa: E402

from oracle import oracle as O  # noqa: E402


def traj(xyz):
    return md.Trajectory(np.array(xyz, dtype=np.float32, copy=True), None)


def scrambled(rng, F, N, groups, sigma, rotate):
    """A reference conformation and F copies of it: noise sigma, the atoms of every group relabelled at random,
    optionally a random rigid motion."""
    ref = (rng.standard_normal((1, N, 3)) * 0.8).astype(np.float32)
    X = np.repeat(ref, F, 0) + sigma * rng.standard_normal((F, N, 3))
    for f in range(F):
        for g in groups:
            g = np.asarray(g, dtype=np.int64)
            if len(g) == 0:
                continue
            X[f, g] = X[f, rng.permutation(g)]
    if rotate:
        R = O.random_rotations(F, rng)
        X = np.einsum("fni,fij->fnj", X, R) + rng.uniform(-2, 2, size=(F, 1, 3))
    return X.astype(np.float32), ref


def main():
    out = {}
    rng = np.random.default_rng(20261017)
    # the reference's own known answer (tests/test_lprmsd.py:12-22) and a few random cost matrices
    out["munkres_known"] = _munkres(np.array([[7, 4, 3], [6, 8, 5], [9, 4, 4]], dtype=np.double))
    costs = rng.random((6, 12, 12))
    out["munkres_costs"] = costs
    out["munkres_masks"] = np.stack([_munkres(np.ascontiguousarray(c)) for c in costs])

    cases = {
        # name: (F, N, atom_indices, permute_groups, sigma, rotate)
        "all_one_group": (6, 24, None, None, 0.02, False),
        "group10_of_50": (8, 50, None, [np.arange(10)], 0.05, True),
        "two_groups_subset": (8, 60, np.r_[2:14, 20:44, 50:58], [np.arange(2, 10), np.arange(24, 36)], 0.05, True),
        "no_groups": (8, 40, np.arange(0, 40, 2), [[]], 0.05, True),
        "waters_100_of_120": (3, 120, None, [np.arange(20, 120)], 0.03, True),
    }
    names = []
    for name, (F, N, idx, groups, sigma, rotate) in cases.items():
        sel_groups = [np.arange(N) if idx is None else idx] if groups is None else groups
        X, ref = scrambled(rng, F, N, sel_groups, sigma, rotate)
        out[name + "_xyz"] = X
        out[name + "_ref"] = ref
        out[name + "_idx"] = np.zeros(0, dtype=np.int64) if idx is None else np.asarray(idx, dtype=np.int64)
        out[name + "_has_idx"] = np.array(idx is not None)
        out[name + "_has_groups"] = np.array(groups is not None)
        flat = [np.asarray(g, dtype=np.int64) for g in (groups or [])]
        out[name + "_groups_flat"] = np.concatenate(flat) if flat else np.zeros(0, dtype=np.int64)
        out[name + "_groups_len"] = np.array([len(g) for g in flat], dtype=np.int64)
        out[name + "_lprmsd"] = md.lprmsd(traj(X), traj(ref), 0, atom_indices=idx, permute_groups=groups)
        t = traj(X)
        out[name + "_lprmsd_superpose"] = md.lprmsd(t, traj(ref), 0, atom_indices=idx, permute_groups=groups, superpose=True)
        out[name + "_xyz_superposed"] = np.array(t.xyz, dtype=np.float32)
        names.append(name)
        print(name, out[name + "_lprmsd"][:4])
    out["cases"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "lprmsd_outputs.npz"), **out)


if __name__ == "__main__":
    main()
