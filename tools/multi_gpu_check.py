"""Multi-GPU correctness check (run under torchrun on a box with >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

  * one-vs-many: frames sharded across ranks + all_gather == single-GPU result, bit for bit;
  * all-pairs: frames broadcast over NCCL, each rank's row block == the same rows of the single-GPU matrix;
  * superpose: frames sharded across ranks + all_gather == single-GPU result, bit for bit.
"""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mdtraj_b200 as mdb  # noqa: E402
from mdtraj_b200 import distributed as D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    mdb.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(0)
    F, N = 4000, 300
    X = rng.standard_normal((F, N, 3), dtype=np.float32)
    t = mdb.Trajectory(X.copy())
    # ---- one-vs-many, sharded over ranks
    full = D.rmsd_sharded(t, t, 5)
    single = mdb.rmsd(t, t, 5)
    assert np.array_equal(full, single), "sharded one-vs-many differs from single GPU"
    # ---- all-pairs: rank 0's frames are broadcast, every rank computes its row block
    if rank == 0:
        dt = mdb.DeviceTrajectory.from_host(X, dev)
    else:
        dt = mdb.DeviceTrajectory(torch.zeros((F, N, 3), dtype=torch.float32, device=dev), N)
    r0, r1, blk = D.rmsd_matrix_sharded(dt, symmetric=False)
    ref = mdb.rmsd_matrix_device(mdb.DeviceTrajectory.from_host(X, dev), row_block=(r0, r1))
    assert torch.equal(blk, ref), "sharded all-pairs block differs from single GPU"
    # symmetric plan: each unordered pair computed once, transposed blocks exchanged over NCCL send/recv
    r0s, r1s, blk_s = D.rmsd_matrix_sharded(dt, symmetric=True)
    assert (r0s, r1s) == (r0, r1)
    assert (blk_s - ref).abs().max().item() < 2e-6, "symmetric sharded all-pairs differs from the row-block result"
    full = [torch.empty((b - a, F), dtype=torch.float32, device=dev) for a, b in D.all_shard_bounds(F, world)]
    dist.all_gather(full, blk_s) if len({t.shape for t in full}) == 1 else None
    if len({t.shape for t in full}) == 1:
        M = torch.cat(full)
        assert torch.equal(M, M.t()), "symmetric sharded matrix is not exactly symmetric"
    assert (r0, r1) == D.shard_bounds(F, rank, world)
    truth_row = single if r0 <= 5 < r1 else None
    if truth_row is not None:
        got = blk[5 - r0].cpu().numpy()
        assert np.abs(got - truth_row).max() < 5e-6
    # ---- superpose, frames sharded over ranks, blocks all-gathered over NCCL
    Xs = rng.standard_normal((1, N, 3), dtype=np.float32) + 0.1 * rng.standard_normal((F, N, 3), dtype=np.float32) + 2.0
    ts, t1 = mdb.Trajectory(Xs.copy()), mdb.Trajectory(Xs.copy())
    idx = np.arange(0, N, 3)
    D.superpose_sharded(ts, ts, 7, atom_indices=idx)
    t1.superpose(t1, 7, atom_indices=idx)
    assert np.array_equal(ts.xyz, t1.xyz), "sharded superpose differs from single GPU"
    dist.barrier()
    if rank == 0:
        print(f"multi-GPU check ok on {world} GPUs: one-vs-many shards and all-pairs row blocks match single GPU")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
