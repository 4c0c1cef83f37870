"""Small superpose / centring / one-vs-many launches over every kernel geometry, for compute-sanitizer runs
(development aid):  compute-sanitizer --tool racecheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import mdtraj_b200 as mdb

for N, F in ((22, 3000), (50, 2000), (100, 1500), (300, 700), (516, 500), (1000, 400), (1400, 300), (2000, 300), (5000, 160), (8000, 40)):
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
print("done")
