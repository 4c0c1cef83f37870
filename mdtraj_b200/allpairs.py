"""All-pairs RMSD matrix: ``D[i, j] = md.rmsd(traj, traj, i, atom_indices=...)[j]``.

The reference obtains it with a Python loop of ``F`` one-vs-many calls
(``examples/clustering.ipynb:78-81``, ``examples/centroids.ipynb:80-82``); here the
frames are centred and laid out once (``b200rmsd_allpairs_prepare_dev``) and the
matrix comes from one tiled contraction with the QCP solve fused into its epilogue
(``b200rmsd_allpairs_rows_dev``).
"""
from __future__ import annotations

import numpy as np

from . import _capi
from .device import DeviceTrajectory, _stream_ptr, _torch

DIAG_ZERO = 1
FAST_SOLVE = 2
_MAX_ROWS_PER_CALL = 65535 * 32


def configure(min_tc_frames=None, max_refs=None, cta_pair=None):
    """Process-wide settings (``b200rmsd_allpairs_configure``): ``min_tc_frames`` -- trajectories with at least this many
    frames take the tcgen05 kernel, shorter ones the exact-fp32 SIMT kernel (default 512); ``max_refs`` -- upper bound on
    the reference structures ``prepare`` may choose (default and maximum 32).  Do not change ``min_tc_frames`` between
    ``prepare`` and the ``rows``/``block`` calls that use its workspace.  ``cta_pair`` (``b200rmsd_allpairs_set_cta_pair``):
    True (default) runs the tcgen05 kernel on 2-CTA clusters (a quarter less operand traffic, 2-15 % faster on one B200),
    False on single CTAs; the matrices are bit-identical."""
    _capi.check(_capi.lib().b200rmsd_allpairs_configure(int(min_tc_frames or 0), int(max_refs or 0)),
                "b200rmsd_allpairs_configure")
    if cta_pair is not None:
        _capi.lib().b200rmsd_allpairs_set_cta_pair(1 if cta_pair else 0)


class PreparedAllPairs:
    """Opaque device workspace (aligned operands + traces) for one trajectory/selection."""

    def __init__(self, workspace, n_frames, n_sel):
        self.workspace = workspace
        self.n_frames = int(n_frames)
        self.n_sel = int(n_sel)

    @property
    def device(self):
        return self.workspace.device

    def info(self):
        """What the prepare step chose: ``n_refs`` reference structures (their frame indices in ``ref_frames``), ``n_far``
        frames stored as they are because no reference is near them, ``cover_radius`` (largest RMSD of a frame to its
        nearest reference, nm).  All zero on the SIMT path (fewer than ``min_tc_frames`` frames)."""
        import ctypes as C
        torch = _torch()
        n_refs, n_far, rad = C.c_int(0), C.c_int(0), C.c_float(0)
        frames = (C.c_int * 32)()
        with torch.cuda.device(self.device):
            rc = _capi.lib().b200rmsd_allpairs_info_dev(self.workspace.data_ptr(), self.workspace.numel(), C.byref(n_refs),
                                                        C.byref(n_far), C.byref(rad), frames, 32,
                                                        _stream_ptr(torch, self.device))
        _capi.check(rc, "b200rmsd_allpairs_info_dev")
        return {"n_refs": n_refs.value, "n_far": n_far.value, "cover_radius": float(rad.value),
                "ref_frames": [int(frames[i]) for i in range(n_refs.value)]}


def prepare(traj: DeviceTrajectory, atom_indices=None) -> PreparedAllPairs:
    torch = _torch()
    dev = traj.device
    idx = traj._index_tensor(atom_indices, traj.n_atoms, "atom_indices")
    n_sel = traj.n_atoms if idx is None else int(idx.numel())
    if n_sel == 0:
        raise ValueError("Number of atom indices must be greater than 0")
    L = _capi.lib()
    nbytes = L.b200rmsd_allpairs_workspace_bytes(traj.n_frames, n_sel)
    ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)  # torch allocations are >= 256-byte aligned
    with torch.cuda.device(dev):
        rc = L.b200rmsd_allpairs_prepare_dev(traj.xyz_dev.data_ptr(), traj.n_frames, traj.n_atoms, traj.frame_stride,
                                             None if idx is None else idx.data_ptr(), n_sel, ws.data_ptr(),
                                             ws.numel(), _stream_ptr(torch, dev))
    _capi.check(rc, "b200rmsd_allpairs_prepare_dev")
    return PreparedAllPairs(ws, traj.n_frames, n_sel)


def rows(prep: PreparedAllPairs, row0: int, row1: int, out=None, diag_zero=True, precise=True):
    """Rows ``[row0, row1)`` as a CUDA float32 tensor of shape ``(row1-row0, F)``.

    ``precise=True`` (default) solves the QCP polynomial in float64 (closer to the float64 truth than the
    reference); ``precise=False`` uses an all-float32 solve -- the reference's own precision class -- which is
    ~1.4x faster on the tensor-core path."""
    torch = _torch()
    dev = prep.device
    F = prep.n_frames
    if not (0 <= row0 <= row1 <= F):
        raise ValueError("row block out of range")
    if out is None:
        out = torch.empty((row1 - row0, F), dtype=torch.float32, device=dev)
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (row1 - row0, F)
    L = _capi.lib()
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        for r0 in range(row0, row1, _MAX_ROWS_PER_CALL):
            r1 = min(row1, r0 + _MAX_ROWS_PER_CALL)
            sub = out[r0 - row0: r1 - row0]
            rc = L.b200rmsd_allpairs_rows_dev(prep.workspace.data_ptr(), prep.workspace.numel(), F, prep.n_sel, r0,
                                              r1, sub.data_ptr(), out.stride(0),
                                              (DIAG_ZERO if diag_zero else 0) | (0 if precise else FAST_SOLVE), stream)
            _capi.check(rc, "b200rmsd_allpairs_rows_dev")
    return out


def block(prep: PreparedAllPairs, row0, row1, col0, col1, out, out_t=None, diag_zero=True, precise=True, out_t_ptr=None,
          ld_t=None):
    """Rows [row0,row1) x columns [col0,col1) into ``out`` (a (row1-row0, >=col1) CUDA tensor, absolute column index);
    ``out_t`` (col1-col0, row1-row0), when given, also receives the transposed block (for shipping to the rank that
    owns rows [col0,col1), see distributed.rmsd_matrix_sharded).  ``out_t_ptr``/``ld_t``: the same as a raw device
    address and leading dimension -- a window of another process's buffer opened with ``b200rmsd_peer_open``."""
    torch = _torch()
    dev = prep.device
    assert out.is_cuda and out.dtype == torch.float32 and out.stride(1) == 1 and out.shape[0] == row1 - row0
    assert out.shape[1] >= col1
    if out_t is not None:
        assert out_t.is_cuda and out_t.dtype == torch.float32 and out_t.stride(1) == 1
        assert out_t.shape == (col1 - col0, row1 - row0)
    L = _capi.lib()
    with torch.cuda.device(dev):
        rc = L.b200rmsd_allpairs_block_dev(prep.workspace.data_ptr(), prep.workspace.numel(), prep.n_frames, prep.n_sel,
                                           row0, row1, col0, col1, out.data_ptr(), out.stride(0),
                                           out_t_ptr if out_t is None else out_t.data_ptr(),
                                           (ld_t or 0) if out_t is None else out_t.stride(0),
                                           (DIAG_ZERO if diag_zero else 0) | (0 if precise else FAST_SOLVE),
                                           _stream_ptr(torch, dev))
    _capi.check(rc, "b200rmsd_allpairs_block_dev")
    return out


def block_rotations(prep: PreparedAllPairs, row0, row1, col0, col1, diag_zero=True, precise=True):
    """Rows [row0,row1) x columns [col0,col1) with the superposition of every pair (``b200rmsd_allpairs_block_rot_dev``):
    returns ``(D, U)`` -- ``D`` (row1-row0, col1-col0) float32 RMSDs and ``U`` (row1-row0, col1-col0, 3, 3) float32
    rotations with ``(x_j - centroid_j) @ U[i-row0, j-col0] ~ x_i - centroid_i`` over the selected atoms: what
    ``md.rmsd`` / ``Trajectory.superpose`` find for target frame j against reference frame i.  Both come out of the same
    accumulator tile of the fused epilogue; 40 bytes per pair are written, so ask for blocks (centres x members)."""
    torch = _torch()
    dev = prep.device
    F = prep.n_frames
    if not (0 <= row0 <= row1 <= F and 0 <= col0 <= col1 <= F):
        raise ValueError("block out of range")
    nr, nc = row1 - row0, col1 - col0
    D = torch.empty((nr, nc), dtype=torch.float32, device=dev)
    U = torch.empty((nr, nc, 3, 3), dtype=torch.float32, device=dev)
    if nr == 0 or nc == 0:
        return D, U
    L = _capi.lib()
    with torch.cuda.device(dev):
        # `out` is addressed with the absolute column index: pass the address column 0 would have
        rc = L.b200rmsd_allpairs_block_rot_dev(prep.workspace.data_ptr(), prep.workspace.numel(), F, prep.n_sel, row0, row1,
                                               col0, col1, D.data_ptr() - 4 * col0, nc, U.data_ptr(),
                                               (DIAG_ZERO if diag_zero else 0) | (0 if precise else FAST_SOLVE),
                                               _stream_ptr(torch, dev))
    _capi.check(rc, "b200rmsd_allpairs_block_rot_dev")
    return D, U


def rmsd_matrix_device(traj: DeviceTrajectory, atom_indices=None, row_block=None, diag_zero=True, precise=True):
    """Full (or ``row_block=(r0, r1)``) matrix as a CUDA tensor; nothing is copied to the host."""
    prep = prepare(traj, atom_indices)
    r0, r1 = (0, traj.n_frames) if row_block is None else row_block
    return rows(prep, r0, r1, diag_zero=diag_zero, precise=precise)


def rmsd_matrix(traj, atom_indices=None, diag_zero=True, precise=True, out=None, row_block=2048) -> np.ndarray:
    """All-pairs RMSD matrix of ``traj`` (host ``Trajectory``/``mdtraj.Trajectory`` or ``DeviceTrajectory``).

    Returns a float32 ndarray ``(F, F)`` with ``D[i, j] == rmsd(traj, traj, i, atom_indices)[j]``.

    ``out``: optional C-contiguous float32 ``(F, F)`` ndarray to receive the matrix.  The device->host copy of a 20 000^2
    matrix (1.6 GB) takes three times as long as computing it, so with ``out`` the matrix is produced in blocks of
    ``row_block`` rows and each block is copied (asynchronously when ``out`` is page-locked) while the next one is being
    computed; every entry is then computed (no mirrored half), which the copy hides.
    """
    if not isinstance(traj, DeviceTrajectory):
        traj = DeviceTrajectory.from_trajectory(traj)
    if out is None:
        return rmsd_matrix_device(traj, atom_indices, None, diag_zero, precise).cpu().numpy()
    torch = _torch()
    F = traj.n_frames
    if not (isinstance(out, np.ndarray) and out.dtype == np.float32 and out.shape == (F, F) and out.flags.c_contiguous
            and out.flags.writeable):
        raise ValueError("out must be a writeable C-contiguous float32 ndarray of shape (n_frames, n_frames)")
    host = torch.from_numpy(out)
    dev = traj.device
    prep = prepare(traj, atom_indices)
    rb = max(1, min(int(row_block), F))
    with torch.cuda.device(dev):
        compute = torch.cuda.current_stream(dev)
        copy = torch.cuda.Stream(dev)
        bufs = [torch.empty((rb, F), dtype=torch.float32, device=dev) for _ in range(2)]
        freed = [None, None]
        for b, r0 in enumerate(range(0, F, rb)):
            r1 = min(F, r0 + rb)
            buf = bufs[b % 2][: r1 - r0]
            if freed[b % 2] is not None:
                compute.wait_event(freed[b % 2])  # the copy that last read this buffer
            rows(prep, r0, r1, out=buf, diag_zero=diag_zero, precise=precise)
            done = torch.cuda.Event()
            done.record(compute)
            copy.wait_event(done)
            with torch.cuda.stream(copy):
                host[r0:r1].copy_(buf, non_blocking=True)
                freed[b % 2] = torch.cuda.Event()
                freed[b % 2].record(copy)
        copy.synchronize()
    return out
