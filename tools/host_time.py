"""Host-array pipeline timing: md.rmsd / superpose on pageable vs page-locked host memory (development aid).
    python tools/host_time.py [frames] [atoms] [copy_threads]"""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch
import mdtraj_b200 as mdb

F = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
if len(sys.argv) > 3:
    mdb.set_host_pipeline(copy_threads=int(sys.argv[3]))
CHUNKS = [int(v) for v in sys.argv[4:]] or [16]
rng = np.random.default_rng(0)
X = rng.standard_normal((F, N, 3), dtype=np.float32)
pinned = torch.empty((F, N, 3), dtype=torch.float32).pin_memory()
pinned.copy_(torch.from_numpy(X))
ref = mdb.Trajectory(X[:1].copy())
res = {"F": F, "N": N, "args": sys.argv[3:], "host_cpus": os.cpu_count(), "gpus": torch.cuda.device_count()}


def traj(arr):
    t = mdb.Trajectory.__new__(mdb.Trajectory)
    t.topology, t._xyz, t._rmsd_traces = None, arr, None
    return t


def timeit(fn, n=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


for devs in ([0], None) if torch.cuda.device_count() > 1 else ([0],):
    mdb.set_devices(devs)
    tag = "1gpu" if devs else "allgpu"
    for mb in CHUNKS:
        mdb.set_host_pipeline(staged_chunk_mb=mb)
        t = traj(X)
        dt = timeit(lambda: mdb.rmsd(t, ref, 0))
        res[f"rmsd_pageable_{mb}MB_{tag}"] = {"ms": round(dt * 1e3, 2), "rmsd_per_s": round(F / dt), "h2d_GBs": round(F * N * 12 / dt / 1e9, 2)}
        t = traj(X.copy())
        dt = timeit(lambda: t.superpose(ref, 0, atom_indices=np.arange(0, N, 5)), n=2)
        res[f"superpose_pageable_{mb}MB_{tag}"] = {"ms": round(dt * 1e3, 2), "frames_per_s": round(F / dt), "each_way_GBs": round(F * N * 12 / dt / 1e9, 2)}
    t = traj(pinned.numpy())
    dt = timeit(lambda: mdb.rmsd(t, ref, 0))
    res[f"rmsd_pinned_{tag}"] = {"ms": round(dt * 1e3, 2), "rmsd_per_s": round(F / dt), "h2d_GBs": round(F * N * 12 / dt / 1e9, 2)}
    dt = timeit(lambda: t.superpose(ref, 0, atom_indices=np.arange(0, N, 5)), n=2)
    res[f"superpose_pinned_{tag}"] = {"ms": round(dt * 1e3, 2), "frames_per_s": round(F / dt), "each_way_GBs": round(F * N * 12 / dt / 1e9, 2)}
# plain memcpy bandwidth of this host, one thread, for scale
Y = np.empty_like(X[: F // 4])
t0 = time.perf_counter(); np.copyto(Y, X[: F // 4]); res["numpy_copy_GBs_1thread"] = Y.nbytes / (time.perf_counter() - t0) / 1e9
print(json.dumps(res))
