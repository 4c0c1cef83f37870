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


class _PeerBuffer:
    """A float32 (rows, cols) device buffer other processes can map (``b200rmsd_peer_alloc``); ``torch.as_tensor`` wraps
    it without a copy through ``__cuda_array_interface__`` and keeps this object (and so the allocation) alive."""

    def __init__(self, rows, cols):
        import ctypes
        from . import _capi
        self.shape = (int(rows), int(cols))
        self._ptr = ctypes.c_void_p()
        self._handle = ctypes.create_string_buffer(64)
        _capi.check(_capi.lib().b200rmsd_peer_alloc(max(1, rows * cols * 4), ctypes.byref(self._ptr), self._handle),
                    "b200rmsd_peer_alloc")
        self.ptr = int(self._ptr.value)
        self.handle = bytes(self._handle.raw)

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": "<f4", "data": (self.ptr, False), "version": 3, "strides": None}

    def free(self):
        from . import _capi
        if getattr(self, "ptr", 0):
            _capi.lib().b200rmsd_peer_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass


class PeerExchange:
    """Peer-visible row blocks of an (n_frames, n_frames) matrix, one per rank of ``group``, each mapped into every rank
    that writes into it (CUDA IPC over NVLink).  Building one is a collective and costs tens of milliseconds (a
    ``cudaMalloc`` of the row block, one handle exchange, one ``cudaIpcOpenMemHandle`` per peer), so it is built once per
    (group, n_frames) and reused: ``rmsd_matrix_sharded`` keeps the exchanges it has built in a small cache.

    ``block`` is this rank's (row1-row0, n_frames) CUDA tensor.  It is the memory the other ranks' kernels write into,
    so a matrix computed through this exchange lives there until the next one is computed through the same exchange
    (``.clone()`` what must outlive it).  ``ok`` is False on every rank when any rank could not allocate or map."""

    def __init__(self, n_frames: int, group=None, device=None):
        import ctypes
        import torch
        from . import _capi
        dist = _dist()
        L = _capi.lib()
        self.group, self.n_frames = group, int(n_frames)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.bounds = all_shard_bounds(self.n_frames, self.world)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.plan = symmetric_plan(self.n_frames, self.world)[self.rank]
        self.opened, self._buf, self.block, self.uses = {}, None, None, 0
        r0, r1 = self.bounds[self.rank]
        err = None
        try:
            with torch.cuda.device(self.device):
                self._buf = _PeerBuffer(r1 - r0, self.n_frames)
        except Exception as e:  # noqa: BLE001
            err = repr(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, None if self._buf is None else self._buf.handle, group)
        if all(h is not None for h in handles):
            try:
                with torch.cuda.device(self.device):
                    for dst in sorted({c[4] for c in self.plan["compute"] if c[4] is not None}):
                        p = ctypes.c_void_p()
                        _capi.check(L.b200rmsd_peer_open(handles[dst], ctypes.byref(p)), "b200rmsd_peer_open")
                        self.opened[dst] = int(p.value)
            except Exception as e:  # noqa: BLE001
                err = repr(e)
        else:
            err = err or "a peer could not allocate"
        oks = [None] * self.world
        dist.all_gather_object(oks, err, group)
        self.error = next((e for e in oks if e is not None), None)
        self.ok = self.error is None
        if self.ok:
            self.block = torch.as_tensor(self._buf, device=self.device)
            self._token = torch.zeros(1, dtype=torch.float32, device=self.device)
        else:
            self.close(collective=False)

    def handshake(self):
        """Stream-ordered meeting point of the ranks (a one-float all-reduce, no host synchronisation): the work every
        rank queued before it has completed, on every rank, before anything queued after it starts."""
        _dist().all_reduce(self._token, group=self.group)

    def close(self, collective=True):
        """Unmap the peers' blocks and free this rank's (a collective when the exchange was usable: nobody frees a
        block another rank may still be writing to)."""
        import torch
        from . import _capi
        if collective and self.ok:
            torch.cuda.synchronize(self.device)
            _dist().barrier(self.group)
        with torch.cuda.device(self.device):
            for p in self.opened.values():
                _capi.lib().b200rmsd_peer_close(p)
        self.opened = {}
        self.block = None
        if collective and self.ok:
            _dist().barrier(self.group)   # every rank has unmapped this rank's block before it is freed
        if self._buf is not None:
            self._buf.free()
            self._buf = None
        self.ok = False


_EXCHANGES = {}       # (group id, world, n_frames, device index) -> PeerExchange, most recently used last
_MAX_EXCHANGES = 2


def peer_exchange(n_frames: int, group=None, device=None):
    """The cached ``PeerExchange`` of (group, n_frames), built on first use (collective: every rank of ``group`` calls
    this with the same arguments, which ``rmsd_matrix_sharded`` guarantees).  At most two stay alive; the least recently
    used one is closed first -- the same decision on every rank, since every rank sees the same sequence of calls."""
    import torch
    dist = _dist()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (id(group) if group is not None else 0, dist.get_world_size(group), int(n_frames), dev.index)
    ex = _EXCHANGES.pop(key, None)
    if ex is None:
        while len(_EXCHANGES) >= _MAX_EXCHANGES:
            _EXCHANGES.pop(next(iter(_EXCHANGES))).close()
        ex = PeerExchange(n_frames, group, dev)
    _EXCHANGES[key] = ex
    return ex


def release_peer_exchanges():
    """Close every cached exchange (collective; call before ``destroy_process_group``)."""
    while _EXCHANGES:
        _EXCHANGES.pop(next(iter(_EXCHANGES))).close()


def _matrix_sharded_peer(prep, ex, diag_zero, precise):
    """The symmetric block plan with the transposed blocks written straight into their owners' row blocks over NVLink
    by the kernel that computes them.  Two stream-ordered handshakes bracket the kernels: nobody writes into a block
    its owner may still be reading from the previous matrix, and nobody reads its block before every writer is done."""
    from . import allpairs
    F, (r0, _) = ex.n_frames, ex.bounds[ex.rank]
    out = ex.block
    if ex.uses:
        ex.handshake()
    ex.uses += 1
    for (a0, a1, c0, c1, dst) in sorted(ex.plan["compute"], key=lambda c: c[4] is None):  # remote blocks first
        rows_view = out[a0 - r0: a1 - r0]
        if dst is None:
            allpairs.block(prep, a0, a1, c0, c1, rows_view, None, diag_zero, precise)
        else:   # element (j, i) of rank dst's block: rows of dst start at bounds[dst][0], leading dimension F
            addr = ex.opened[dst] + ((c0 - ex.bounds[dst][0]) * F + a0) * 4
            allpairs.block(prep, a0, a1, c0, c1, rows_view, None, diag_zero, precise, out_t_ptr=addr, ld_t=F)
    ex.handshake()
    return out


def rmsd_matrix_sharded(traj_dev, atom_indices=None, group=None, broadcast=True, diag_zero=True, precise=True,
                        symmetric=True, exchange="auto"):
    """Row-block sharded all-pairs matrix.  ``traj_dev``: a DeviceTrajectory with identical shape on every rank
    (rank 0's coordinates are broadcast over NCCL unless ``broadcast=False``).  Returns ``(row0, row1, block)`` where
    ``block`` is this rank's ``(row1-row0, F)`` CUDA tensor; the matrix is never gathered.

    ``symmetric=True`` (default): every unordered pair of frames is computed once (``symmetric_plan``); the ranks
    exchange transposed blocks over NVLink -- half the tensor work for F^2/(2W) floats of traffic per rank.
    ``exchange="peer"``: the kernel that computes a block writes its transposed copy straight into the owner's row block
    (a buffer mapped through CUDA IPC): the transfer rides under the tensor-core work, nothing is staged or copied on
    arrival.  The row blocks then live in a ``PeerExchange`` that is built on the first call for (group, F) and reused:
    **the returned block is overwritten by the next sharded matrix of the same size** -- ``.clone()`` it to keep it, or
    pass your own ``PeerExchange`` as ``exchange``.  ``exchange="nccl"``: the transposed blocks go through NCCL send/recv
    into a fresh tensor, one group per round of the plan, each issued while the next block is computed.  ``"auto"``
    (default) tries ``"peer"`` and falls back to ``"nccl"`` on every rank when any rank cannot map its peers.
    ``symmetric=False``: every rank computes its whole row block, no exchange.  Call ``release_peer_exchanges()`` before
    ``destroy_process_group``.
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
    plan = symmetric_plan(F, world)[rank]
    if isinstance(exchange, PeerExchange):
        ex = exchange
        if not ex.ok or ex.n_frames != F or ex.world != world:
            raise ValueError("rmsd_matrix_sharded: the PeerExchange does not match this matrix / group")
        return r0, r1, _matrix_sharded_peer(prep, ex, diag_zero, precise)
    if exchange not in ("auto", "peer", "nccl"):
        raise ValueError("exchange must be 'auto', 'peer', 'nccl' or a PeerExchange")
    if exchange != "nccl":
        ex = peer_exchange(F, group, dev)
        if ex.ok:
            return r0, r1, _matrix_sharded_peer(prep, ex, diag_zero, precise)
        if exchange == "peer":
            raise RuntimeError("rmsd_matrix_sharded(exchange='peer'): the peer mapping could not be set up on every "
                               "rank: " + str(ex.error))
    out = torch.empty((r1 - r0, F), dtype=torch.float32, device=dev)

    def compute(a0, a1, c0, c1, rows_view, out_t):
        allpairs.block(prep, a0, a1, c0, c1, rows_view, out_t, diag_zero, precise)
    symmetric_exchange(plan, rank, world, r0, out, compute, group)
    return r0, r1, out


def exchange_rounds(plan_rank, rank: int, world: int):
    """The blocks of one rank's ``symmetric_plan`` entry grouped into rounds: in round d = 1 .. W/2 the rank computes its
    block against rank + d and sends the transposed copy there, and receives the one rank - d computed for it (the
    antipodal round of an even world has half blocks going both ways).  Returns ``[(d, sends, recvs)]`` in the order
    every rank walks them, so that the send/receive groups of a round pair up across ranks."""
    by_round = {}
    for c in plan_rank["compute"]:
        if c[4] is not None:
            by_round.setdefault((c[4] - rank) % world, {"send": [], "recv": []})["send"].append(c)
    for rc in plan_rank["recv"]:
        by_round.setdefault((rank - rc[0]) % world, {"send": [], "recv": []})["recv"].append(rc)
    return [(d, by_round[d]["send"], by_round[d]["recv"]) for d in sorted(by_round)]


def symmetric_exchange(plan_rank, rank, world, row0, out, compute, group=None):
    """Fill ``out`` (this rank's (rows, F) block, any device) from the symmetric block plan with send/recv exchange.
    ``compute(a0, a1, c0, c1, rows_view, out_t)`` writes block rows [a0,a1) x columns [c0,c1) into ``rows_view`` (absolute
    column index) and, when ``out_t`` is not None, its transpose into ``out_t``.  One send and one receive per rank and
    round, issued as one group right after the block's kernel (on CUDA the group waits for it on the compute stream) while
    the next block is already being computed; the diagonal block comes last and covers the last round's transfer.
    Device-agnostic: the world_size 2-4 gloo tests run it on CPU tensors with a numpy stand-in for the kernel."""
    import torch
    dist = _dist()
    works, recvs = [], []
    for _, sends, rcvs in exchange_rounds(plan_rank, rank, world):
        ops = []
        for (a0, a1, c0, c1, dst) in sends:
            t = torch.empty((c1 - c0, a1 - a0), dtype=out.dtype, device=out.device)
            compute(a0, a1, c0, c1, out[a0 - row0: a1 - row0], t)
            ops.append(dist.P2POp(dist.isend, t, dst, group))
        for (src, a0, a1, c0, c1) in rcvs:
            buf = torch.empty((a1 - a0, c1 - c0), dtype=out.dtype, device=out.device)
            recvs.append((a0, a1, c0, c1, buf))
            ops.append(dist.P2POp(dist.irecv, buf, src, group))
        if ops:
            works.extend(dist.batch_isend_irecv(ops))
    for (a0, a1, c0, c1, dst) in plan_rank["compute"]:
        if dst is None:
            compute(a0, a1, c0, c1, out[a0 - row0: a1 - row0], None)
    for w in works:
        w.wait()
    for (a0, a1, c0, c1, buf) in recvs:
        out[a0 - row0: a1 - row0, c0:c1] = buf
    return out


def similarity_scores_sharded(traj_dev, atom_indices=None, beta=1.0, group=None, broadcast=True, precise=True):
    """``np.exp(-beta * D / D.std()).sum(axis=1)`` (``examples/centroids.ipynb:117``) for the all-pairs matrix of a
    trajectory whose matrix is sharded over the ranks of ``group`` (``rmsd_matrix_sharded``): every rank reduces its own
    row block on the device (``b200rmsd_matrix_moments_dev``, ``b200rmsd_exp_rowsum_dev``), the two moments are
    all-reduced (16 bytes), the F scores all-gathered -- the matrix (40 GB at 100k frames) never leaves the GPUs.
    Returns ``(scores float64 ndarray (F,), std)`` on every rank."""
    import math
    import torch
    from . import _capi
    from .device import _stream_ptr
    dist = _dist()
    world = dist.get_world_size(group)
    F = traj_dev.n_frames
    dev = traj_dev.device
    r0, r1, blk = rmsd_matrix_sharded(traj_dev, atom_indices, group, broadcast=broadcast, precise=precise)
    L = _capi.lib()
    moments = torch.zeros(2, dtype=torch.float64, device=dev)
    local = torch.zeros(max(r1 - r0, 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        if r1 > r0:
            _capi.check(L.b200rmsd_matrix_moments_dev(blk.data_ptr(), r1 - r0, F, blk.stride(0), moments.data_ptr(), stream),
                        "b200rmsd_matrix_moments_dev")
        dist.all_reduce(moments, group=group)
        s1, s2 = (float(v) for v in moments.cpu())
        n = float(F) * float(F)
        mean = s1 / n
        std = math.sqrt(max(s2 / n - mean * mean, 0.0))
        if not std > 0.0:
            raise ValueError("all pairwise distances are equal: distances.std() == 0")
        if r1 > r0:
            _capi.check(L.b200rmsd_exp_rowsum_dev(blk.data_ptr(), r1 - r0, F, blk.stride(0), -float(beta) / std, 0,
                                                  local.data_ptr(), stream), "b200rmsd_exp_rowsum_dev")
    bounds = all_shard_bounds(F, world)
    width = max(b - a for a, b in bounds)
    padded = torch.zeros(width, dtype=torch.float64, device=dev)
    padded[: r1 - r0] = local[: r1 - r0]
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    scores = torch.cat([p[: b - a] for p, (a, b) in zip(parts, bounds)])
    return scores.cpu().numpy(), std


def centroid_index_sharded(traj_dev, atom_indices=None, beta=1.0, group=None, broadcast=True, precise=True):
    """Index of the frame most similar to all others (``centroids.ipynb:117-118``) from the sharded matrix; the same
    integer on every rank."""
    return int(np.argmax(similarity_scores_sharded(traj_dev, atom_indices, beta, group, broadcast, precise)[0]))
