"""Consumers of the all-pairs RMSD matrix (SURVEY.md section 8(f) "next" #3).

The reference's example notebooks build the (F, F) matrix with a Python loop of ``md.rmsd`` calls and post-process
it on the host; here the matrix stays in HBM and only the small results cross PCIe:

* ``rmsd_condensed``   -- ``squareform(distances, checks=False)`` (``examples/clustering.ipynb:101``): the condensed upper
  triangle that ``scipy.cluster.hierarchy.linkage`` takes;
* ``similarity_scores`` / ``centroid_index`` -- ``np.exp(-beta * distances / distances.std()).sum(axis=1).argmax()``
  (``examples/centroids.ipynb:117``);
* ``assign_to_leaders`` -- ``np.argmin(md.rmsd(leaders, frame, 0))`` for every frame
  (``examples/two-pass-clustering.ipynb`` cell 13).

Matrices larger than ``max_matrix_bytes`` are never materialised: the reductions run over row blocks (for the
similarity scores the blocks are computed twice, once for the standard deviation and once for the row sums).
"""
from __future__ import annotations

import math

import numpy as np

from . import _capi
from . import allpairs as AP
from .device import DeviceTrajectory, _stream_ptr, _torch

_DEFAULT_MAX_MATRIX_BYTES = 32 << 30


def _as_device(traj):
    return traj if isinstance(traj, DeviceTrajectory) else DeviceTrajectory.from_trajectory(traj)


def _row_blocks(n_rows, n_cols, max_bytes):
    per = max(1, int(max_bytes // (4 * max(n_cols, 1))))
    per = min(per, AP._MAX_ROWS_PER_CALL)
    return [(r, min(n_rows, r + per)) for r in range(0, n_rows, per)]


def rmsd_condensed(traj, atom_indices=None, precise=True, as_numpy=True, dtype=np.float64):
    """Condensed all-pairs distances: ``squareform(D, checks=False)`` with ``D[i] = md.rmsd(traj, traj, i, atom_indices)``.

    Returns ``F*(F-1)/2`` values ordered (0,1), (0,2), ..., (F-2,F-1); a ``dtype`` ndarray (float64 like the notebook's
    matrix) or, with ``as_numpy=False``, the float32 CUDA tensor."""
    torch = _torch()
    dt = _as_device(traj)
    F = dt.n_frames
    D = AP.rmsd_matrix_device(dt, atom_indices, precise=precise)  # symmetric by construction
    out = torch.empty(F * (F - 1) // 2, dtype=torch.float32, device=dt.device)
    if F > 1:
        with torch.cuda.device(dt.device):
            rc = _capi.lib().b200rmsd_condense_dev(D.data_ptr(), F, D.stride(0), out.data_ptr(), _stream_ptr(torch, dt.device))
        _capi.check(rc, "b200rmsd_condense_dev")
    if not as_numpy:
        return out
    return out.cpu().numpy().astype(dtype, copy=False)


def similarity_scores(traj, atom_indices=None, beta=1.0, precise=True, max_matrix_bytes=_DEFAULT_MAX_MATRIX_BYTES):
    """``np.exp(-beta * D / D.std()).sum(axis=1)`` and ``D.std()`` for the all-pairs matrix D of ``traj``.

    Returns ``(scores, std)``: float64 ndarray (F,) and float.  ``D.std()`` is the population standard deviation over all
    F*F entries (zeros on the diagonal included), as numpy computes it in ``examples/centroids.ipynb:117``."""
    torch = _torch()
    dt = _as_device(traj)
    F = dt.n_frames
    dev = dt.device
    L = _capi.lib()
    prep = AP.prepare(dt, atom_indices)
    blocks = _row_blocks(F, F, max_matrix_bytes)
    moments = torch.zeros(2, dtype=torch.float64, device=dev)
    part = torch.empty(2, dtype=torch.float64, device=dev)
    scores = torch.empty(F, dtype=torch.float64, device=dev)
    kept = None
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        for r0, r1 in blocks:
            D = AP.rows(prep, r0, r1, precise=precise)
            _capi.check(L.b200rmsd_matrix_moments_dev(D.data_ptr(), r1 - r0, F, D.stride(0), part.data_ptr(), stream),
                        "b200rmsd_matrix_moments_dev")
            moments += part
            if len(blocks) == 1:
                kept = D
        s1, s2 = (float(v) for v in moments.cpu())
        n = float(F) * float(F)
        mean = s1 / n
        std = math.sqrt(max(s2 / n - mean * mean, 0.0))
        if not std > 0.0:
            raise ValueError("all pairwise distances are equal: distances.std() == 0")
        scale = -float(beta) / std
        for r0, r1 in blocks:
            D = kept if kept is not None else AP.rows(prep, r0, r1, precise=precise)
            _capi.check(L.b200rmsd_exp_rowsum_dev(D.data_ptr(), r1 - r0, F, D.stride(0), scale, 0,
                                                  scores[r0:r1].data_ptr(), stream), "b200rmsd_exp_rowsum_dev")
    return scores.cpu().numpy(), std


def centroid_index(traj, atom_indices=None, beta=1.0, precise=True, max_matrix_bytes=_DEFAULT_MAX_MATRIX_BYTES):
    """Index of the frame most similar to all others: ``similarity_scores(...)[0].argmax()`` (centroids.ipynb:117-118)."""
    return int(np.argmax(similarity_scores(traj, atom_indices, beta, precise, max_matrix_bytes)[0]))


def assign_to_leaders(traj, leaders, atom_indices=None, precise=True, as_numpy=True,
                      max_matrix_bytes=_DEFAULT_MAX_MATRIX_BYTES):
    """Nearest leader of every frame: ``argmin_k md.rmsd(leaders, traj[i], 0, atom_indices)[k]`` for all i, and that RMSD.

    ``leaders`` is a trajectory with the same atoms (host or device).  Returns ``(labels int32 (F,), distances float32 (F,))``.
    The (F, n_leaders) block is one rectangular call of the all-pairs kernel over the concatenated frames."""
    torch = _torch()
    dt = _as_device(traj)
    ld = leaders if isinstance(leaders, DeviceTrajectory) else DeviceTrajectory.from_trajectory(leaders, dt.device)
    if ld.n_atoms != dt.n_atoms:
        raise ValueError("Input trajectories must have same number of atoms. found %d and %d." % (dt.n_atoms, ld.n_atoms))
    F, K = dt.n_frames, ld.n_frames
    if K == 0:
        raise ValueError("leaders has no frames")
    dev = dt.device
    both = DeviceTrajectory(torch.cat([dt.xyz_dev, ld.xyz_dev.to(dev)], dim=0), dt.n_atoms)
    prep = AP.prepare(both, atom_indices)
    labels = torch.empty(F, dtype=torch.int32, device=dev)
    dist = torch.empty(F, dtype=torch.float32, device=dev)
    L = _capi.lib()
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        for r0, r1 in _row_blocks(F, K, max_matrix_bytes):
            blk = torch.empty((r1 - r0, K), dtype=torch.float32, device=dev)
            # columns are addressed absolutely: hand over the address that column 0 would have (only [F, F+K) is written)
            rc = L.b200rmsd_allpairs_block_dev(prep.workspace.data_ptr(), prep.workspace.numel(), F + K, prep.n_sel, r0, r1,
                                               F, F + K, blk.data_ptr() - 4 * F, K, None, 0,
                                               0 if precise else AP.FAST_SOLVE, stream)
            _capi.check(rc, "b200rmsd_allpairs_block_dev")
            _capi.check(L.b200rmsd_row_argmin_dev(blk.data_ptr(), r1 - r0, K, K, labels[r0:r1].data_ptr(),
                                                  dist[r0:r1].data_ptr(), stream), "b200rmsd_row_argmin_dev")
    if not as_numpy:
        return labels, dist
    return labels.cpu().numpy(), dist.cpu().numpy()
