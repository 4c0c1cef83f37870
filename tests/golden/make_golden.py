"""Generate the golden vectors under tests/golden/ by running the REAL reference.

Run in the build container only (it needs /root/reference built and importable):

    # one-off build of the reference into /tmp
    cp -r /root/reference/. /tmp/mdtraj_build && cd /tmp/mdtraj_build
    #   strip `versioneer` from setup.py, then:
    CC=/usr/bin/gcc CXX=/usr/bin/g++ LDSHARED="/usr/bin/g++ -shared" python setup.py build_ext --inplace
    mkdir -p /tmp/shim && ln -s /usr/lib/python3/dist-packages/pip/_vendor/pyparsing /tmp/shim/pyparsing
    PYTHONPATH=/tmp/mdtraj_build:/tmp/shim:/root/repo python tests/golden/make_golden.py

Inputs are either the reference's own fixture (examples/ala2.h5, coordinates committed as
ala2_xyz.npy because /root/reference does not exist on the GPU box) or seeded synthetic
data regenerated at test time by oracle.synth_iid / oracle.synth_md.  Outputs are whatever
the reference's md.rmsd / Trajectory.superpose / center_coordinates return.
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

import mdtraj as md  # the real reference  # noqa: E402

from mdtraj_b200.h5min import load_coordinates  # noqa: E402
from oracle import oracle as O  # noqa: E402


def traj(xyz):
    return md.Trajectory(np.array(xyz, dtype=np.float32, copy=True), None)


def main():
    out = {}
    # ---- C1: examples/ala2.h5 --------------------------------------------------------------
    ala2 = load_coordinates("/root/reference/examples/ala2.h5")
    np.save(os.path.join(HERE, "ala2_xyz.npy"), ala2)
    t = traj(ala2)
    out["ala2_rmsd_frame0"] = md.rmsd(t, t, 0)
    # clustering.ipynb:78-81 -- all-pairs, all atoms
    t = traj(ala2)
    D = np.empty((t.n_frames, t.n_frames), dtype=np.float32)
    for i in range(t.n_frames):
        D[i] = md.rmsd(t, t, i)
    out["ala2_allpairs"] = D
    print("Max pairwise rmsd: %f nm" % np.max(D))  # notebook prints 0.188493
    assert abs(float(np.max(D)) - 0.188493) < 5e-7
    # centroids.ipynb:80-82,117 -- heavy atoms
    heavy = np.array([1, 4, 5, 6, 8, 10, 14, 15, 16, 18])
    t = traj(ala2)
    Dh = np.empty((t.n_frames, t.n_frames), dtype=np.float32)
    for i in range(t.n_frames):
        Dh[i] = md.rmsd(t, t, i, atom_indices=heavy)
    out["ala2_heavy_idx"] = heavy
    out["ala2_allpairs_heavy"] = Dh
    beta = 1
    index = np.exp(-beta * Dh / Dh.std()).sum(axis=1).argmax()
    print("centroid index", index)  # notebook: 83
    assert index == 83
    out["ala2_centroid_index"] = np.int64(index)
    # precentered path on ala2
    t = traj(ala2)
    t.center_coordinates()
    out["ala2_centered_xyz"] = t.xyz.copy()
    out["ala2_traces"] = np.asarray(t._rmsd_traces).copy()
    out["ala2_rmsd_frame5_precentered"] = md.rmsd(t, t, 5, precentered=True)
    # superpose=False
    t = traj(ala2)
    out["ala2_rmsd_frame3_nosuperpose"] = md.rmsd(t, t, 3, superpose=False)
    # superpose with and without indices
    t = traj(ala2); r = traj(ala2)
    t.superpose(r, 7)
    out["ala2_superposed_frame7"] = t.xyz.copy()
    t = traj(ala2); r = traj(ala2)
    t.superpose(r, 2, atom_indices=heavy)
    out["ala2_superposed_frame2_heavy"] = t.xyz.copy()

    # ---- seeded synthetic cases (inputs regenerated from the seed at test time) -------------
    cases = [("iid", 64, 100, 11), ("iid", 33, 22, 12), ("iid", 16, 1000, 13), ("md", 40, 303, 14), ("iid", 5, 4100, 15)]
    for kind, F, N, seed in cases:
        gen = O.synth_iid if kind == "iid" else O.synth_md
        X = gen(F, N, seed=seed)
        key = f"{kind}_{F}x{N}_s{seed}"
        t = traj(X)
        out[key + "_rmsd_f1"] = md.rmsd(t, t, 1)
        idx = np.arange(0, N, 3)
        t = traj(X)
        out[key + "_rmsd_f2_idx3"] = md.rmsd(t, t, 2, atom_indices=idx)
        ridx = idx[::-1].copy()
        t = traj(X)
        out[key + "_rmsd_f0_idx3_refrev"] = md.rmsd(t, t, 0, atom_indices=idx, ref_atom_indices=ridx)
        t = traj(X)
        out[key + "_rmsd_f1_nosup"] = md.rmsd(t, t, 1, superpose=False)
        t = traj(X); r = traj(X)
        t.superpose(r, 1, atom_indices=idx)
        out[key + "_superposed_f1_idx3"] = t.xyz.copy()
        t = traj(X); r = traj(X)
        t.superpose(r, 0)
        out[key + "_superposed_f0"] = t.xyz.copy()
        t = traj(X)
        t.center_coordinates()
        out[key + "_traces"] = np.asarray(t._rmsd_traces).copy()
        out[key + "_centered_max_abs_mean"] = np.float64(np.abs(t.xyz.mean(1)).max())

    # ---- "next" rows: rmsf and the align/displace entry point ---------------------------------------
    from mdtraj import _rmsd as ref_rmsd
    for kind, F, N, seed in (("md", 40, 303, 14), ("iid", 64, 100, 11)):
        gen = O.synth_iid if kind == "iid" else O.synth_md
        X = gen(F, N, seed=seed)
        key = f"{kind}_{F}x{N}_s{seed}"
        out[key + "_rmsf_f1"] = md.rmsf(traj(X), traj(X), 1)
        idx = np.arange(0, N, 3)
        out[key + "_rmsf_f2_idx3"] = md.rmsf(traj(X), traj(X), 2, atom_indices=idx)
        out[key + "_rmsf_prealigned"] = md.rmsf(traj(X), None)
        t = traj(X); t.center_coordinates()
        out[key + "_rmsf_f0_precentered"] = md.rmsf(t, t, 0, precentered=True)
        # getMultipleAlignDisplaceRMSDs_atom_major: align on the first 60 atoms, measure on the last 40
        n_al, n_di = 60, 40
        A = np.zeros((F, 60, 3), np.float32); D = np.zeros((F, 40, 3), np.float32)
        A[:] = X[:, :n_al]; D[:] = X[:, N - n_di:]
        A -= A.mean(1, keepdims=True)
        gA = np.einsum("ijk,ijk->i", A, A).astype(np.float32)
        r, rot = ref_rmsd.getMultipleAlignDisplaceRMSDs_atom_major(A, A, gA, gA, D, D, n_al, n_di, 3)
        out[key + "_aligndispl_rmsd"] = np.asarray(r).copy(); out[key + "_aligndispl_rot"] = np.asarray(rot).copy()

    # ---- error/warning behaviour captured verbatim (Appendix B #12) ---------------------------
    msgs = {}
    t = traj(O.synth_iid(5, 10, 1))
    for name, fn in (
        ("bad_index", lambda: md.rmsd(t, t, 0, atom_indices=[0, 10])),
        ("len_mismatch", lambda: md.rmsd(t, t, 0, atom_indices=[0, 1], ref_atom_indices=[0])),
        ("bad_frame", lambda: md.rmsd(t, t, 5)),
        ("ref_only_indices", lambda: md.rmsd(t, t, 0, ref_atom_indices=[0, 1])),
        ("empty_superpose", lambda: t.superpose(t, 0, atom_indices=[])),
    ):
        try:
            fn()
            msgs[name] = "NO ERROR"
        except Exception as e:  # noqa: BLE001
            msgs[name] = f"{type(e).__name__}: {e}"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        md.rmsd(t, t, 0, precentered=True)
        msgs["warn_precentered_no_traces"] = f"{w[-1].category.__name__}: {w[-1].message}"
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        md.rmsd(t, t, 0, precentered=True, superpose=False)
        msgs["warn_precentered_nosuperpose"] = f"{w[-1].category.__name__}: {w[-1].message}"
    for k, v in msgs.items():
        print(k, "->", v)
        out["msg_" + k] = np.array(v)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
