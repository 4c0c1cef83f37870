"""Small all-pairs launches on the tensor-core path (full matrix, windows, index lists, ragged K) for compute-sanitizer
runs (development aid):  compute-sanitizer --tool memcheck python tools/sanitize_allpairs.py"""
import os, sys
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import mdtraj_b200 as mdb
from mdtraj_b200 import allpairs as AP

os.environ["B200RMSD_ALLPAIRS"] = "tc"
for N, F, idx in ((22, 130, None), (300, 700, None), (97, 1001, None), (250, 64, np.arange(0, 250, 5)), (33, 520, [3, 11])):
    dt = mdb.DeviceTrajectory.synthetic_iid(F, N, seed=3)
    prep = AP.prepare(dt, idx)
    D = AP.rows(prep, 0, F)
    blk = torch.zeros((F // 3, F), dtype=torch.float32, device=dt.device)
    t = torch.zeros((F // 2, F // 3), dtype=torch.float32, device=dt.device)
    AP.block(prep, 7, 7 + F // 3, F // 4, F // 4 + F // 2, blk, t)
    torch.cuda.synchronize()
    print(N, F, "asym", float((D - D.t()).abs().max()), "window ok", bool(torch.equal(blk[:, F // 4: F // 4 + F // 2].t(), t)), flush=True)
print("done")
