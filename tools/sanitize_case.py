"""Small superpose / centring / one-vs-many launches over every kernel geometry, for compute-sanitizer runs
(development aid):  compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import mdtraj_b200 as mdb

for N, F in ((22, 3000), (50, 2000), (100, 1500), (300, 700), (516, 500), (1000, 400), (1400, 300), (2000, 300), (5000, 160), (6200, 60), (8000, 40)):
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=3)
    ref = mdb.DeviceTrajectory(dt.xyz_dev[:1].clone(), N)
    r0 = mdb.rmsd_device(dt, ref, 0, as_numpy=False)
    idx = np.arange(0, N, 5)
    a = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    a.superpose(ref, 0, atom_indices=idx)
    b = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    b.superpose(ref, 0)
    r1 = mdb.rmsd_device(b, ref, 0, superpose=False, as_numpy=False)
    c = mdb.DeviceTrajectory(dt.xyz_dev.clone(), N)
    c.center_coordinates()
    torch.cuda.synchronize()
    print(N, F, "max |superposed plain - qcp| =", float((r1 - r0)[1:].abs().max()), flush=True)
# round 2: all-pairs block with rotations (ROT instantiation of the tcgen05 kernel and the SIMT kernel), md.lprmsd
from mdtraj_b200 import allpairs as AP
for min_tc, F, N in ((1, 600, 64), (1 << 30, 200, 30)):
    AP.configure(min_tc_frames=min_tc)
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=4)
    D, U = AP.block_rotations(AP.prepare(dt), 7, 130, 3, F - 5)
    torch.cuda.synchronize()
    print("rotations", min_tc, float(D.max()), float(torch.linalg.det(U.double()).sub(1).abs().max()), flush=True)
AP.configure(min_tc_frames=512)
rng = np.random.default_rng(5)
for N, groups in ((60, [np.arange(10, 30), np.arange(30, 60)]), (300, None), (700, [np.arange(100, 700)])):
    ref = rng.standard_normal((1, N, 3)).astype(np.float32)
    X = (np.repeat(ref, 50, 0) + 0.05 * rng.standard_normal((50, N, 3))).astype(np.float32)
    d = mdb.lprmsd(mdb.DeviceTrajectory.from_host(X), mdb.Trajectory(ref), permute_groups=groups, superpose=True)
    print("lprmsd", N, float(d.mean()), flush=True)
print("done")
