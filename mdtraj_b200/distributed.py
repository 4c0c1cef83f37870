"""Multi-GPU sharding of the RMSD path: one process per GPU, ``torch.distributed`` for plumbing.

* one-vs-many (``rmsd``, ``superpose``, ``center_coordinates``): frames are independent, so every
  rank takes a contiguous frame block ``[F*r/W, F*(r+1)/W)``; the single reference frame is
  replicated; results are gathered (``all_gather`` of F floats).  No exchange step on the data path.
* all-pairs: the frames are broadcast once (NCCL over NVLink for CUDA tensors), every rank prepares
  the same operands and computes its own row block of the matrix, which stays on that rank.

The functions take any initialised process group: NCCL on GPUs, gloo on CPU (used by the
world_size=2 tests of the partition / gather logic, which inject a stand-in for the per-shard kernel).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced partition; identical on every rank; covers [0, n_items) exactly once."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (n_items * rank) // world, (n_items * (rank + 1)) // world


def all_shard_bounds(n_items: int, world: int):
    return [shard_bounds(n_items, r, world) for r in range(world)]


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def gather_frames(local: "np.ndarray | object", n_total: int, group=None, device=None):
    """All-gather ragged per-rank result blocks (1-D float32) into the full (n_total,) array on every rank."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = all_shard_bounds(n_total, world)
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float32))
    if device is None and not t.is_cuda and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())  # NCCL moves device memory only
    if device is not None:
        t = t.to(device)
    assert t.numel() == bounds[rank][1] - bounds[rank][0], "local block does not match this rank's shard"
    width = max(b - a for a, b in bounds)
    padded = torch.zeros(width, dtype=torch.float32, device=t.device)
    padded[: t.numel()] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    full = torch.cat([p[: b - a] for p, (a, b) in zip(parts, bounds)])
    return full if is_tensor else full.cpu().numpy()


def rmsd_sharded(target, reference, frame=0, atom_indices=None, ref_atom_indices=None, precentered=False,
                 superpose=True, group=None, shard_fn=None):
    """``md.rmsd`` with the target frames split across the ranks of ``group``; every rank returns the full
    (F,) float32 result.  ``target`` is the *whole* host trajectory on every rank (each rank only touches
    its block).  ``shard_fn(sub_target, reference, ...)`` defaults to ``mdtraj_b200.rmsd``.
    """
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    F = target.xyz.shape[0]
    a, b = shard_bounds(F, rank, world)
    if shard_fn is None:
        from ._rmsd import rmsd as shard_fn
    from .trajectory import Trajectory
    sub = Trajectory.__new__(Trajectory)
    sub.topology = getattr(target, "topology", None)
    sub._xyz = target.xyz[a:b]                      # a view: no copy of the shard
    tr = getattr(target, "_rmsd_traces", None)
    sub._rmsd_traces = None if tr is None else np.asarray(tr)[a:b]
    local = shard_fn(sub, reference, frame, atom_indices, ref_atom_indices, True, precentered, superpose)
    full = gather_frames(np.asarray(local, dtype=np.float32), F, group)
    if superpose and atom_indices is None and target is reference:
        full[frame % F] = 0.0  # same-memory shortcut holds for the whole trajectory, not per shard
    return full


def gather_rows(local, n_total: int, group=None):
    """All-gather ragged per-rank row blocks ``(n_local, ...)`` float32 into the full ``(n_total, ...)`` array."""
    import torch
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bounds = all_shard_bounds(n_total, world)
    is_tensor = isinstance(local, torch.Tensor)
    t = local if is_tensor else torch.from_numpy(np.ascontiguousarray(local, dtype=np.float32))
    if not t.is_cuda and dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", torch.cuda.current_device()))
    assert t.shape[0] == bounds[rank][1] - bounds[rank][0], "local block does not match this rank's shard"
    width = max(b - a for a, b in bounds)
    padded = torch.zeros((width,) + tuple(t.shape[1:]), dtype=torch.float32, device=t.device)
    padded[: t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    full = torch.cat([p[: b - a] for p, (a, b) in zip(parts, bounds)])
    return full if is_tensor else full.cpu().numpy()


def superpose_sharded(traj, reference, frame=0, atom_indices=None, ref_atom_indices=None, group=None, gather=True,
                      shard_fn=None):
    """``Trajectory.superpose`` with the frames of ``traj`` split across the ranks of ``group`` (SURVEY.md section 8(e):
    frames are independent, the reference frame is replicated, there is no exchange step).  ``traj`` is the whole host
    trajectory on every rank; each rank superposes its contiguous block in place.  With ``gather=True`` the blocks are
    all-gathered so that ``traj.xyz`` is the fully superposed trajectory on every rank; otherwise only
    ``traj.xyz[a:b]`` (the returned bounds) is valid.  ``shard_fn(sub, reference, frame, atom_indices,
    ref_atom_indices)`` defaults to ``Trajectory.superpose`` of this package."""
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    F = traj.xyz.shape[0]
    a, b = shard_bounds(F, rank, world)
    from .trajectory import Trajectory
    ref_frame = np.array(reference.xyz[frame], dtype=np.float32, copy=True)  # before this rank's block is modified
    ref = Trajectory(ref_frame[None])
    sub = Trajectory(np.ascontiguousarray(traj.xyz[a:b], dtype=np.float32))
    if b > a:
        if shard_fn is None:
            sub.superpose(ref, 0, atom_indices, ref_atom_indices)
        else:
            shard_fn(sub, ref, 0, atom_indices, ref_atom_indices)
    if gather:
        traj.xyz = gather_rows(sub.xyz, F, group)
    else:
        traj.xyz[a:b] = sub.xyz
    return a, b


def broadcast_frames(xyz_dev, src=0, group=None):
    """Broadcast a staged trajectory tensor from ``src`` to every rank (NCCL over NVLink on GPUs)."""
    dist = _dist()
    dist.broadcast(xyz_dev, src=src, group=group)
    return xyz_dev


def symmetric_plan(n_frames: int, world: int):
    """Work split of the symmetric all-pairs matrix over ``world`` ranks that own contiguous row blocks.

    D[i][j] == D[j][i], so each off-diagonal block pair {(r,s),(s,r)} is computed once and its transpose shipped.
    Returns ``plan[rank] = {"compute": [(r0,r1,c0,c1,send_to)], "recv": [(src, r0,r1,c0,c1)]}`` in absolute frame
    indices; ``send_to`` is None for blocks that stay local.  Rank r computes its diagonal block, the blocks (r,s)
    with 0 < (s-r) mod W < W/2, and -- for even W -- half of the antipodal block (r, r+W/2), so the work is balanced.
    Every entry of every rank's row block is produced exactly once (tests/test_host_logic.py checks this).
    """
    bounds = all_shard_bounds(n_frames, world)
    plan = [{"compute": [], "recv": []} for _ in range(world)]
    for r in range(world):
        r0, r1 = bounds[r]
        if r1 > r0:
            plan[r]["compute"].append((r0, r1, r0, r1, None))
        for s in range(world):
            if s == r:
                continue
            s0, s1 = bounds[s]
            if s1 == s0 or r1 == r0:
                continue
            d = (s - r) % world
            if 2 * d < world:                      # r computes the whole block (r,s) and ships its transpose to s
                plan[r]["compute"].append((r0, r1, s0, s1, s))
                plan[s]["recv"].append((r, s0, s1, r0, r1))
            elif 2 * d == world and r < s:         # antipodal pair: split the rows of the lower rank in two halves
                h = r0 + (r1 - r0) // 2
                if h > r0:                         # r computes rows [r0,h) x cols of s, ships the transpose to s
                    plan[r]["compute"].append((r0, h, s0, s1, s))
                    plan[s]["recv"].append((r, s0, s1, r0, h))
                if r1 > h:                         # s computes its rows x cols [h,r1) of r, ships the transpose to r
                    plan[s]["compute"].append((s0, s1, h, r1, r))
                    plan[r]["recv"].append((s, h, r1, s0, s1))
    return plan


def rmsd_matrix_sharded(traj_dev, atom_indices=None, group=None, broadcast=True, diag_zero=True, precise=True,
                        symmetric=True):
    """Row-block sharded all-pairs matrix.  ``traj_dev``: a DeviceTrajectory with identical shape on every rank
    (rank 0's coordinates are broadcast over NCCL unless ``broadcast=False``).  Returns ``(row0, row1, block)`` where
    ``block`` is this rank's ``(row1-row0, F)`` CUDA tensor; the matrix is never gathered.

    ``symmetric=True`` (default): every unordered pair of frames is computed once (``symmetric_plan``); the ranks
    exchange transposed blocks with NCCL send/recv over NVLink -- half the tensor work for F^2/(2W) floats of traffic
    per rank.  ``symmetric=False``: every rank computes its whole row block, no exchange.
    """
    import torch
    from . import allpairs
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if broadcast and world > 1:
        broadcast_frames(traj_dev.xyz_dev, 0, group)
    prep = allpairs.prepare(traj_dev, atom_indices)
    F = traj_dev.n_frames
    r0, r1 = shard_bounds(F, rank, world)
    if not symmetric or world == 1:
        return r0, r1, allpairs.rows(prep, r0, r1, diag_zero=diag_zero, precise=precise)
    dev = traj_dev.device
    out = torch.empty((r1 - r0, F), dtype=torch.float32, device=dev)
    plan = symmetric_plan(F, world)[rank]
    sends = []
    for (a0, a1, c0, c1, dst) in plan["compute"]:
        rows_view = out[a0 - r0: a1 - r0]
        if dst is None:
            allpairs.block(prep, a0, a1, c0, c1, rows_view, None, diag_zero, precise)
        else:
            t = torch.empty((c1 - c0, a1 - a0), dtype=torch.float32, device=dev)
            allpairs.block(prep, a0, a1, c0, c1, rows_view, t, diag_zero, precise)
            sends.append((dst, t))
    recvs = [(src, a0, a1, c0, c1, torch.empty((a1 - a0, c1 - c0), dtype=torch.float32, device=dev))
             for (src, a0, a1, c0, c1) in plan["recv"]]
    ops = [dist.P2POp(dist.isend, t, dst, group) for dst, t in sends] + \
          [dist.P2POp(dist.irecv, buf, src, group) for (src, _, _, _, _, buf) in recvs]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for (_, a0, a1, c0, c1, buf) in recvs:
        out[a0 - r0: a1 - r0, c0:c1] = buf
    return r0, r1, out
