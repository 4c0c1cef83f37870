"""One-vs-many accuracy against float64 truth on MD-like and chain-like frames (development aid).
    python tools/ovm_accuracy.py  ->  one JSON line per (generator, N)"""
import json, os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import mdtraj_b200 as mdb
from _truth import truth_rmsd_batch


def rotations(n, rng):
    q = rng.standard_normal((n, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    return np.stack([np.stack([a*a+b*b-c*c-d*d, 2*(b*c-a*d), 2*(b*d+a*c)], -1),
                     np.stack([2*(b*c+a*d), a*a-b*b+c*c-d*d, 2*(c*d-a*b)], -1),
                     np.stack([2*(b*d-a*c), 2*(c*d+a*b), a*a-b*b-c*c+d*d], -1)], 1)


def frames(kind, F, N, seed, rg):
    rng = np.random.default_rng(seed)
    if kind == "cloud":
        base = rng.standard_normal((N, 3)) * rg
    else:
        st = rng.standard_normal((N, 3)); st /= np.linalg.norm(st, axis=1, keepdims=True)
        base = np.cumsum(0.15 * st, 0); base -= base.mean(0)
    X = base[None] + rng.standard_normal((F, N, 3)) * 0.1
    X = np.einsum("fni,fij->fnj", X, rotations(F, rng)) + rng.uniform(-5, 5, size=(F, 1, 3))
    return X.astype(np.float32)


def main():
    for kind, N, F, rg in (("cloud", 300, 4000, 1.0), ("cloud", 700, 4000, 1.5), ("cloud", 1000, 4000, 1.5), ("cloud", 1000, 4000, 3.0),
                           ("chain", 300, 4000, 0), ("chain", 700, 4000, 0), ("chain", 1000, 4000, 0), ("chain", 4000, 1000, 0),
                           ("cloud", 5000, 800, 3.0), ("chain", 5000, 800, 0), ("cloud", 25000, 160, 5.0), ("chain", 25000, 160, 0)):
        X = frames(kind, F, N, 11, rg)
        truth = truth_rmsd_batch(X, X[0])
        got = mdb.rmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(X.copy()), 0)
        idx = np.arange(0, N, 3)
        got_i = mdb.rmsd(mdb.Trajectory(X.copy()), mdb.Trajectory(X.copy()), 0, atom_indices=idx)
        truth_i = truth_rmsd_batch(X[:, idx], X[0, idx])
        t = mdb.Trajectory(X.copy()); t.center_coordinates()
        got_p = mdb.rmsd(t, t, 0, precentered=True)
        print(json.dumps({"kind": kind, "N": N, "F": F, "rg": rg, "max_abs_vs_truth": float(np.abs(got[1:] - truth[1:]).max()),
                          "every3rd_vs_truth": float(np.abs(got_i[1:] - truth_i[1:]).max()),
                          "precentered_vs_truth": float(np.abs(got_p[1:] - truth[1:]).max())}), flush=True)


if __name__ == "__main__":
    main()
