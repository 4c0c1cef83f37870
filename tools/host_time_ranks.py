"""End-to-end md.rmsd on PAGEABLE host arrays when several ranks share one host (torchrun, one rank per GPU): the
staging-chunk sweep behind b200rmsd_host_configure's defaults.  Every rank pushes its own F x N frames through its GPU;
time = barrier -> call -> barrier, max over ranks; aggregate rmsd/s and per-GPU H2D GB/s per setting, one JSON line each.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
        tools/host_time_ranks.py [frames per rank] [atoms] [copy threads per rank] [staged chunk MB ...]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
F = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
threads = int(sys.argv[3]) if len(sys.argv) > 3 else max(2, len(os.sched_getaffinity(0)) // world - 1)
chunks = [int(v) for v in sys.argv[4:]] or [16, 8, 4, 2, 1]
torch.cuda.set_device(local)
mdb.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
mdb.set_host_pipeline(copy_threads=threads)
X = np.random.default_rng(rank).standard_normal((F, N, 3), dtype=np.float32)
pinned = torch.empty((F, N, 3), dtype=torch.float32).pin_memory()
pinned.copy_(torch.from_numpy(X))
ref = mdb.Trajectory(X[:1].copy())


def traj(arr):
    t = mdb.Trajectory.__new__(mdb.Trajectory)
    t.topology, t._xyz, t._rmsd_traces = None, arr, None
    return t


def timed(fn, n=3):
    fn()
    best = []
    for _ in range(n):
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        fn()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best.append(dt.item())
    return sorted(best)[len(best) // 2]


t = traj(X)
pieces = [int(v) for v in os.environ.get("STAGE_PIECES_KB", "0,-1").split(",")]   # 0 = automatic, -1 = whole chunks
for kb in pieces:
    for mb in chunks:
        mdb.set_host_pipeline(staged_chunk_mb=mb, stage_piece_kb=kb)
        dt = timed(lambda: mdb.rmsd(t, ref, 0))
        if rank == 0:
            print(json.dumps({"what": "md.rmsd pageable", "ranks": world, "frames_per_rank": F, "atoms": N,
                              "copy_threads": threads, "staged_chunk_mb": mb, "stage_piece_kb": kb, "ms": dt * 1e3,
                              "rmsd_per_s": world * F / dt, "h2d_GBs_per_gpu": F * N * 12 / dt / 1e9}), flush=True)
mdb.set_host_pipeline(staged_chunk_mb=16, stage_piece_kb=0)
ts = traj(X.copy())
dt = timed(lambda: ts.superpose(ref, 0, atom_indices=np.arange(0, N, 5)), n=2)
if rank == 0:
    print(json.dumps({"what": "superpose pageable", "ranks": world, "frames_per_rank": F, "atoms": N, "ms": dt * 1e3,
                      "frames_per_s": world * F / dt, "each_way_GBs_per_gpu": F * N * 12 / dt / 1e9}), flush=True)
tp = traj(pinned.numpy())
dt = timed(lambda: mdb.rmsd(tp, ref, 0))
if rank == 0:
    print(json.dumps({"what": "md.rmsd page-locked", "ranks": world, "frames_per_rank": F, "atoms": N, "ms": dt * 1e3,
                      "rmsd_per_s": world * F / dt, "h2d_GBs_per_gpu": F * N * 12 / dt / 1e9}), flush=True)
    os.system("lscpu | grep -E 'Model name|Socket|NUMA node\\(s\\)|L3|^CPU\\(s\\)' 1>&2")
if world > 1:
    dist.destroy_process_group()
