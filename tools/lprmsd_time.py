"""Timing of md.lprmsd on the GPU (device-resident frames, CUDA events).  One JSON line per case.  For scale: the real
reference (md.lprmsd of baseline/_ref, Munkres on the dense cost matrix) took 145.6 s for 200 frames x 300 atoms, one
group, in the build container = 1.4 frames/s on one core."""
import json
import os
import sys

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
    print(json.dumps(rec), flush=True)
