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
    # symmetric plan: each unordered pair computed once; the transposed blocks are written into the owner's row block by
    # the computing kernel (peer memory over NVLink) or shipped with NCCL send/recv -- the same matrix bit for bit
    blocks = {}
    for exchange in (os.environ.get("MGC_ORDER", "peer,nccl").split(",")):
        r0s, r1s, blk_s = D.rmsd_matrix_sharded(dt, symmetric=True, exchange=exchange)
        assert (r0s, r1s) == (r0, r1)
        # D[i][j] here is the value computed for the pair (j, i) on another rank: the float32 epilogue solve accepts an
        # estimated error of 4e-6 nm + 5e-5 relative (qcp_msd_shift; ~40 pairs per million of iid frames are off by
        # 2e-6 .. 7e-6 nm at an RMSD of 2.4 nm), so the two evaluations agree to the parity tolerance, not to the bit
        bad = (blk_s - ref).abs() >= 1e-5
        if bad.any():   # where: rows / columns (absolute) of the first and last wrong entry
            idx = bad.nonzero()
            print(f"[rank {rank}] {exchange}: {int(bad.sum())} wrong entries of {bad.numel()}, rows {int(idx[:, 0].min()) + r0}.."
                  f"{int(idx[:, 0].max()) + r0}, cols {int(idx[:, 1].min())}..{int(idx[:, 1].max())}, "
                  f"max err {float((blk_s - ref).abs().max())}", flush=True)
        assert not bad.any(), f"symmetric sharded all-pairs ({exchange}) differs from the row-block result"
        full = [torch.empty((b - a, F), dtype=torch.float32, device=dev) for a, b in D.all_shard_bounds(F, world)]
        if len({t.shape for t in full}) == 1:
            dist.all_gather(full, blk_s)
            M = torch.cat(full)
            assert torch.equal(M, M.t()), f"symmetric sharded matrix ({exchange}) is not exactly symmetric"
        blocks[exchange] = blk_s.clone()   # a "peer" block lives in the exchange's buffer until the next call
        if exchange == "peer":             # second matrix through the same (cached) exchange: the handshake path
            blk_s.zero_()
            _, _, again = D.rmsd_matrix_sharded(dt, symmetric=True, exchange="peer")
            assert again.data_ptr() == blk_s.data_ptr() and torch.equal(again, blocks["peer"]), \
                "a reused peer exchange gives a different matrix"
    if len(blocks) == 2:
        assert torch.equal(blocks["peer"], blocks["nccl"]), "peer-memory and NCCL exchanges disagree"
    # timing of the two exchanges on a larger matrix (device time, max over ranks)
    if os.environ.get("MGC_TIME"):
        Fb = int(os.environ["MGC_TIME"])
        big = mdb.DeviceTrajectory.synthetic_iid(Fb, N, seed=2000, device=dev)
        for exchange in ("peer", "nccl"):    # untimed: builds the peer exchange, fills the allocator's cache
            D.rmsd_matrix_sharded(big, symmetric=True, exchange=exchange, broadcast=False)
        for exchange in ("peer", "nccl", "peer", "nccl"):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            D.rmsd_matrix_sharded(big, symmetric=True, exchange=exchange, broadcast=False)
            e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if rank == 0:
                print(f'{{"what": "rmsd_matrix_sharded", "frames": {Fb}, "atoms": {N}, "gpus": {world}, "exchange": "{exchange}", "ms": {ms.item():.3f}}}', flush=True)
        del big
    # more matrix sizes than the exchange cache holds: the oldest exchange is closed (a collective) and a new one built
    for Fx in (1100, 1237, 1400, 1100):
        Xx = X[:Fx]
        dx = mdb.DeviceTrajectory.from_host(Xx, dev) if rank == 0 else \
            mdb.DeviceTrajectory(torch.zeros((Fx, N, 3), dtype=torch.float32, device=dev), N)
        a0, a1, bx = D.rmsd_matrix_sharded(dx, symmetric=True, exchange="peer")
        want = mdb.rmsd_matrix_device(mdb.DeviceTrajectory.from_host(Xx, dev), row_block=(a0, a1))
        assert (bx - want).abs().max().item() < 1e-5, f"peer exchange after cache turnover, F = {Fx}"
    assert len(D._EXCHANGES) <= 2
    # the clustering notebooks' reduction over the sharded matrix == the single-GPU one
    Xc = X[:1500]
    dc = mdb.DeviceTrajectory.from_host(Xc, dev) if rank == 0 else \
        mdb.DeviceTrajectory(torch.zeros((1500, N, 3), dtype=torch.float32, device=dev), N)
    sc, sd = D.similarity_scores_sharded(dc, beta=1.0)
    from mdtraj_b200 import clustering
    sc1, sd1 = clustering.similarity_scores(mdb.DeviceTrajectory.from_host(Xc, dev), beta=1.0)
    assert abs(sd - sd1) < 1e-6 * sd1 and np.allclose(sc, sc1, rtol=1e-6, atol=0), "sharded similarity scores"
    assert D.centroid_index_sharded(dc) == int(np.argmax(sc1))
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
    D.release_peer_exchanges()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
