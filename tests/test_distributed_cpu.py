"""CPU, world_size = 2 over gloo: the N>1 host path -- frame partition, ragged all-gather, shortcut
placement.  The per-shard kernel is replaced by the CPU oracle (test infrastructure standing in for the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, F, N, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mdtraj_b200 as mdb
        from mdtraj_b200 import distributed as D
        from oracle import oracle as O
        X = O.synth_iid(F, N, seed=5)
        t = mdb.Trajectory(X.copy())

        def shard_fn(sub, reference, frame, ai, rai, parallel, precentered, superpose):
            # stand-in for mdtraj_b200.rmsd on this rank's GPU
            return O.rmsd(sub.xyz, reference.xyz, frame, ai, rai, superpose=superpose, impl="port")
        full = D.rmsd_sharded(t, t, 3, shard_fn=shard_fn)
        a, b = D.shard_bounds(F, rank, world)
        blk = D.gather_frames(np.full(b - a, float(rank), np.float32), F)
        q.put((rank, full, blk))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("F", [7, 64])
def test_rmsd_sharded_world2(F):
    from oracle import oracle as O
    N, world = 30, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, F, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = O.synth_iid(F, N, seed=5)
    want = O.rmsd(X, X, 3, impl="port")
    want[3] = 0.0
    for rank, full, blk in res:
        assert full.shape == (F,) and np.array_equal(full, want)  # identical to the unsharded result, bit for bit
        assert np.array_equal(blk, np.concatenate([np.zeros(F // 2, np.float32), np.ones(F - F // 2, np.float32)]))


def _worker_superpose(rank, world, port, F, N, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import mdtraj_b200 as mdb
        from mdtraj_b200 import distributed as D
        from oracle import oracle as O
        X = O.synth_md(F, N, seed=9)
        t = mdb.Trajectory(X.copy())
        idx = np.arange(0, N, 2)

        def shard_fn(sub, ref, frame, ai, rai):  # stand-in for this rank's GPU
            sub.xyz = O.superpose(sub.xyz, ref.xyz, frame, ai, rai)
        bounds = D.superpose_sharded(t, t, 4, atom_indices=idx, shard_fn=shard_fn)
        q.put((rank, bounds, np.asarray(t.xyz).copy()))
    finally:
        dist.destroy_process_group()


def test_superpose_sharded_world2():
    """Frames split over two ranks, reference frame replicated, blocks all-gathered: identical to the unsharded call."""
    from oracle import oracle as O
    F, N, world = 11, 24, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_superpose, args=(r, world, port, F, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    X = O.synth_md(F, N, seed=9)
    want = O.superpose(X, X, 4, np.arange(0, N, 2))
    assert [r[1] for r in res] == [(0, 5), (5, 11)]
    for rank, bounds, xyz in res:
        assert xyz.shape == (F, N, 3) and np.array_equal(xyz, want)


def _worker_exchange(rank, world, port, F, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mdtraj_b200 import distributed as D
        # a known symmetric "distance" matrix stands in for the all-pairs kernel
        i = np.arange(F, dtype=np.float64)
        M = torch.from_numpy((np.abs(i[:, None] - i[None, :]) + 0.001 * (i[:, None] * i[None, :] % 97)).astype(np.float32))
        calls = []

        def compute(a0, a1, c0, c1, rows_view, out_t):   # what allpairs.block does on this rank's GPU
            calls.append((a0, a1, c0, c1))
            rows_view[:, c0:c1] = M[a0:a1, c0:c1]
            if out_t is not None:
                out_t.copy_(M[a0:a1, c0:c1].t())
        r0, r1 = D.shard_bounds(F, rank, world)
        out = torch.full((r1 - r0, F), float("nan"), dtype=torch.float32)
        D.symmetric_exchange(D.symmetric_plan(F, world)[rank], rank, world, r0, out, compute)
        q.put((rank, (r0, r1), out.numpy(), torch.equal(out, M[r0:r1]), sum((a1 - a0) * (c1 - c0) for a0, a1, c0, c1 in calls)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F", [(2, 37), (3, 50), (4, 64), (4, 3)])
def test_symmetric_exchange_over_gloo(world, F):
    """The all-pairs exchange step (symmetric block plan, rounds, paired send/recv groups) on CPU tensors over gloo with a
    stand-in for the kernel: every rank ends up with exactly its rows of the matrix, every entry produced once, and the
    computed entries split evenly (each unordered block pair on one rank)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, F, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    computed = 0
    for rank, (r0, r1), out, same, n in res:
        assert same and not np.isnan(out).any(), f"rank {rank}: row block differs from the matrix"
        computed += n
    # diagonal blocks are computed in full by the stand-in (the kernel mirrors them), off-diagonal pairs once
    bounds = [(F * r // world, F * (r + 1) // world) for r in range(world)]
    diag = sum((b - a) ** 2 for a, b in bounds)
    assert computed == diag + (F * F - diag) // 2


def test_exchange_rounds_pair_up():
    """Host logic of the exchange: in every round what rank r sends to s is what s expects from r in the SAME round."""
    from mdtraj_b200 import distributed as D
    for world in range(2, 10):
        for F in (world, 10 * world + 3, 257):
            plan = D.symmetric_plan(F, world)
            rounds = [dict((d, (s, r)) for d, s, r in D.exchange_rounds(plan[k], k, world)) for k in range(world)]
            for k in range(world):
                for d, (sends, _) in rounds[k].items():
                    for (a0, a1, c0, c1, dst) in sends:
                        assert d in rounds[dst], (world, F, k, d)
                        assert (k, c0, c1, a0, a1) in rounds[dst][d][1], (world, F, k, dst, d)
                for d, (_, recvs) in rounds[k].items():
                    for (src, a0, a1, c0, c1) in recvs:
                        assert (c0, c1, a0, a1, k) in rounds[src][d][0]
