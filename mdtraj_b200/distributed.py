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


def broadcast_frames(xyz_dev, src=0, group=None):
    """Broadcast a staged trajectory tensor from ``src`` to every rank (NCCL over NVLink on GPUs)."""
    dist = _dist()
    dist.broadcast(xyz_dev, src=src, group=group)
    return xyz_dev


def rmsd_matrix_sharded(traj_dev, atom_indices=None, group=None, broadcast=True, diag_zero=True):
    """Row-block sharded all-pairs matrix.  ``traj_dev``: a DeviceTrajectory with identical shape on every rank
    (rank 0's coordinates are broadcast unless ``broadcast=False``).  Returns ``(row0, row1, block)`` where
    ``block`` is this rank's ``(row1-row0, F)`` CUDA tensor; the matrix is never gathered."""
    from . import allpairs
    dist = _dist()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if broadcast and world > 1:
        broadcast_frames(traj_dev.xyz_dev, 0, group)
    prep = allpairs.prepare(traj_dev, atom_indices)
    r0, r1 = shard_bounds(traj_dev.n_frames, rank, world)
    return r0, r1, allpairs.rows(prep, r0, r1, diag_zero=diag_zero)
