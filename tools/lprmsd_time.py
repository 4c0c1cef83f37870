"""Timing of md.lprmsd on the GPU (device-resident frames, CUDA events) and, with --cpu, of the oracle's restatement
of the reference loop on a few frames (scipy's exact assignment in place of Munkres: a LOWER bound on the reference's
time -- the real Munkres took 0.73 s per frame at 300 atoms in the build container).  One JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402

CASES = [  # frames, atoms, distinguishable, groups
    (20000, 300, 60, 2), (20000, 300, 0, 1), (4000, 1000, 100, 3), (2000, 2000, 200, 1)]
rng = np.random.default_rng(0)
for F, N, nd, ng in CASES:
    ref = (rng.standard_normal((1, N, 3)) * 1.5).astype(np.float32)
    X = (np.repeat(ref, F, 0) + 0.02 * rng.standard_normal((F, N, 3))).astype(np.float32)
    bounds = np.linspace(nd, N, ng + 1).astype(int)
    groups = [np.arange(bounds[i], bounds[i + 1]) for i in range(ng)]
    for g in groups:
        X[:, g] = X[:, rng.permutation(g)]
    dt = mdb.DeviceTrajectory.from_host(X)
    r = mdb.Trajectory(ref)
    mdb.lprmsd(dt, r, permute_groups=groups)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d = mdb.lprmsd(dt, r, permute_groups=groups)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rec = {"what": "lprmsd", "frames": F, "atoms": N, "distinguishable": nd, "groups": [len(g) for g in groups],
           "ms": ms, "frames_per_s": F / (ms * 1e-3), "rmsd_mean": float(d.mean())}
    if "--cpu" in sys.argv:
        from oracle import oracle as O
        t0 = time.perf_counter()
        O.lprmsd(X[:8], ref, 0, None, groups, impl="reference" if O.ref_available() else "port")
        rec["cpu_restatement_frames_per_s_1core"] = 8 / (time.perf_counter() - t0)
    print(json.dumps(rec), flush=True)
