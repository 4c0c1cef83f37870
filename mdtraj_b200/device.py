"""Device-resident trajectories: coordinates staged to HBM once, reused by every call.

``DeviceTrajectory`` owns a torch CUDA tensor in the padded atom-major layout the
kernels stream (``include/b200rmsd.h``: float32, ``n_pad = 4*ceil(N/4)`` atoms per
frame, padding zero).  torch is used for allocation, streams and (in
``mdtraj_b200.distributed``) NCCL -- plumbing only; every number is produced by
``libb200rmsd.so``.

Mirrors the slice of ``mdtraj.Trajectory`` the RMSD path touches
(``core/trajectory.py:1083-1173, 2114-2137, 1376-1384``): ``xyz``-like access,
``_rmsd_traces`` caching, ``center_coordinates``, ``superpose``.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _capi


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mdtraj_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def _stream_ptr(torch, device):
    return torch.cuda.current_stream(device).cuda_stream


class _Scratch:
    """Grow-only scratch tensor for the _dev entry points, one per (device, stream): calls issued on different streams
    may run concurrently and must not share partial-sum buffers."""
    _buf = {}

    @classmethod
    def get(cls, torch, device, nbytes):
        index = device.index if device.index is not None else torch.cuda.current_device()
        key = (index, torch.cuda.current_stream(device).cuda_stream)
        buf = cls._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
            cls._buf[key] = buf
        return buf


class PreparedReference:
    """Centred, packed reference conformation + its statistics, on the device."""

    def __init__(self, ref, stats, n_sel):
        self.ref = ref
        self.stats = stats
        self.n_sel = n_sel

    @property
    def trace(self) -> float:
        return float(self.stats.cpu().view(dtype=_torch().float64)[0])

    @property
    def centroid(self) -> np.ndarray:
        return self.stats.cpu().view(dtype=_torch().float64)[4:7].numpy().copy()


def prepare_reference(frame_dev, idx_dev, n_sel, do_center=True, given_trace=0.0) -> PreparedReference:
    """Wrap b200rmsd_prepare_reference_dev.  ``frame_dev``: 1-D/2-D float32 CUDA tensor holding one frame."""
    torch = _torch()
    dev = frame_dev.device
    n_pad = (n_sel + 3) // 4 * 4
    ref = torch.empty(n_pad * 3, dtype=torch.float32, device=dev)
    stats = torch.empty(_capi.REFSTATS_BYTES, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _capi.lib().b200rmsd_prepare_reference_dev(
            frame_dev.data_ptr(), None if idx_dev is None else idx_dev.data_ptr(), n_sel, int(do_center),
            float(given_trace), ref.data_ptr(), stats.data_ptr(), _stream_ptr(torch, dev))
    _capi.check(rc, "b200rmsd_prepare_reference_dev")
    return PreparedReference(ref, stats, n_sel)


class DeviceTrajectory:
    """A trajectory whose coordinates live in HBM, padded atom-major float32."""

    _is_b200_device_trajectory = True

    def __init__(self, xyz_padded, n_atoms, traces=None, topology=None):
        torch = _torch()
        assert xyz_padded.is_cuda and xyz_padded.dtype == torch.float32 and xyz_padded.dim() == 3
        assert xyz_padded.shape[2] == 3 and xyz_padded.shape[1] % 4 == 0 and xyz_padded.is_contiguous()
        assert xyz_padded.shape[1] == (n_atoms + 3) // 4 * 4
        self._xyz = xyz_padded
        self.n_atoms = int(n_atoms)
        self._rmsd_traces = traces
        self.topology = topology

    # ---- construction -------------------------------------------------
    @classmethod
    def from_host(cls, xyz, device=None, topology=None, traces=None):
        """Stage a host (F,N,3) array once.  ``xyz`` may be pinned; float32 is required (cast otherwise)."""
        torch = _torch()
        if device is None:
            from ._rmsd import current_device
            device = torch.device("cuda", current_device())
        device = torch.device(device)
        xyz = np.asarray(xyz)
        if xyz.ndim == 2:
            xyz = xyz[None]
        if xyz.ndim != 3 or xyz.shape[2] != 3:
            raise ValueError("xyz must have shape (n_frames, n_atoms, 3)")
        src = torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.float32))
        F, N, _ = src.shape
        n_pad = (N + 3) // 4 * 4
        if n_pad == N:
            dev = src.to(device, non_blocking=True)
        else:
            dev = torch.zeros((F, n_pad, 3), dtype=torch.float32, device=device)
            dev[:, :N, :] = src.to(device, non_blocking=True)
        tr = None
        if traces is not None:
            tr = torch.from_numpy(np.ascontiguousarray(traces, dtype=np.float32)).to(device)
        return cls(dev, N, tr, topology)

    @classmethod
    def from_trajectory(cls, traj, device=None):
        return cls.from_host(traj.xyz, device=device, topology=getattr(traj, "topology", None),
                             traces=getattr(traj, "_rmsd_traces", None))

    @classmethod
    def synthetic_iid(cls, n_frames, n_atoms, seed=0, device=None):
        """N(0,1) coordinates generated on the device (bench/test helper; never touches the host)."""
        torch = _torch()
        if device is None:
            from ._rmsd import current_device
            device = torch.device("cuda", current_device())
        device = torch.device(device)
        g = torch.Generator(device=device)
        g.manual_seed(int(seed))
        n_pad = (n_atoms + 3) // 4 * 4
        x = torch.zeros((n_frames, n_pad, 3), dtype=torch.float32, device=device)
        chunk = max(1, (1 << 28) // (n_pad * 3))
        for f0 in range(0, n_frames, chunk):
            f1 = min(n_frames, f0 + chunk)
            x[f0:f1, :n_atoms, :] = torch.randn((f1 - f0, n_atoms, 3), generator=g, dtype=torch.float32, device=device)
        return cls(x, n_atoms)

    # ---- views --------------------------------------------------------
    @property
    def device(self):
        return self._xyz.device

    @property
    def n_frames(self) -> int:
        return int(self._xyz.shape[0])

    @property
    def n_pad(self) -> int:
        return int(self._xyz.shape[1])

    @property
    def frame_stride(self) -> int:
        return self.n_pad * 3

    @property
    def xyz_dev(self):
        """(F, n_pad, 3) CUDA tensor (padding atoms are zero and must stay zero)."""
        return self._xyz

    @property
    def xyz(self) -> np.ndarray:
        """Host copy (F, N, 3) float32 -- a device->host transfer, for inspection and tests."""
        return self._xyz[:, : self.n_atoms, :].cpu().numpy()

    def __len__(self):
        return self.n_frames

    def __getitem__(self, key):
        torch = _torch()
        if isinstance(key, (int, np.integer)):
            key = slice(key, key + 1) if key != -1 else slice(key, None)
        sub = self._xyz[key].contiguous()
        tr = None if self._rmsd_traces is None else self._rmsd_traces[key].contiguous()
        return DeviceTrajectory(sub, self.n_atoms, tr, self.topology)

    def to_trajectory(self):
        from .trajectory import Trajectory
        t = Trajectory(self.xyz, self.topology)
        if self._rmsd_traces is not None:
            t._rmsd_traces = self._rmsd_traces.cpu().numpy()
        return t

    # ---- operations ---------------------------------------------------
    def center_coordinates(self, mass_weighted=False):
        """In-place centring + trace caching (``Trajectory.center_coordinates``, trajectory.py:2114-2137)."""
        if mass_weighted:
            raise NotImplementedError("mass-weighted centring needs a topology; out of scope of the RMSD path")
        torch = _torch()
        traces = torch.empty(self.n_frames, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _capi.lib().b200rmsd_center_trace_dev(self._xyz.data_ptr(), self.n_frames, self.n_atoms,
                                                       self.frame_stride, traces.data_ptr(),
                                                       _stream_ptr(torch, self.device))
        _capi.check(rc, "b200rmsd_center_trace_dev")
        self._rmsd_traces = traces
        return self

    def _index_tensor(self, idx, n_atoms, name):
        torch = _torch()
        if idx is None:
            return None
        from ._rmsd import _ensure_int_1d
        idx = _ensure_int_1d(idx, name)
        if not np.all((idx >= 0) * (idx < n_atoms)):
            raise ValueError(f"{name} must be valid positive indices")
        return torch.from_numpy(idx.astype(np.int32)).to(self.device)

    def superpose(self, reference, frame=0, atom_indices=None, ref_atom_indices=None, parallel=True,
                  return_rotations=False):
        """In-place superposition onto ``reference[frame]`` (``Trajectory.superpose``, trajectory.py:1083-1173)."""
        torch = _torch()
        if atom_indices is not None and len(atom_indices) == 0:
            raise ValueError("Number of atom indices must be greater than 0")
        if ref_atom_indices is None:
            ref_atom_indices = atom_indices
        if ref_atom_indices is not None and atom_indices is not None and len(ref_atom_indices) != len(atom_indices):
            raise ValueError("Number of atoms must be consistent!")
        ref_dev = reference if isinstance(reference, DeviceTrajectory) else DeviceTrajectory.from_trajectory(
            reference, self.device)
        if atom_indices is None and ref_atom_indices is not None:
            atom_indices = np.arange(len(ref_atom_indices))
        idx = self._index_tensor(atom_indices, self.n_atoms, "atom_indices")
        ridx = self._index_tensor(ref_atom_indices, ref_dev.n_atoms, "ref_atom_indices")
        n_sel = self.n_atoms if idx is None else int(idx.numel())
        if idx is None and ref_dev.n_atoms != self.n_atoms:
            raise ValueError("Number of atoms must be consistent!")
        # the reference copies the reference frame first (trajectory.py:1129), so aliasing is safe
        ref_frame = ref_dev._xyz[frame].clone()
        prep = prepare_reference(ref_frame, ridx, n_sel, True)
        F = self.n_frames
        out = torch.empty(F, dtype=torch.float32, device=self.device)
        rot = torch.empty((F, 3, 3), dtype=torch.float32, device=self.device)
        degen = torch.zeros(1, dtype=torch.int32, device=self.device)
        L = _capi.lib()
        nbytes = L.b200rmsd_scratch_bytes(F, self.n_atoms)
        scratch = _Scratch.get(torch, self.device, nbytes)
        with torch.cuda.device(self.device):
            rc = L.b200rmsd_superpose_dev(self._xyz.data_ptr(), F, self.n_atoms, self.frame_stride,
                                          None if idx is None else idx.data_ptr(), n_sel, prep.ref.data_ptr(),
                                          prep.stats.data_ptr(), out.data_ptr(), rot.data_ptr(), degen.data_ptr(),
                                          scratch.data_ptr(), scratch.numel(), _stream_ptr(torch, self.device))
        _capi.check(rc, "b200rmsd_superpose_dev")
        # overflow / underflow guard of trajectory.py:1162-1169 (rotated centred frame 0 all zeros), evaluated after the
        # re-translation: every atom of frame 0 on one point
        if F and not bool((self._xyz[0, : self.n_atoms] - self._xyz[0, :1]).any().item()):
            raise OverflowError(
                "Encounted a potential overflow/underflow error during superpose() due to the magnitude of your `_xyz`"
                "coordinates. To circumvent this, (1) reload your trajectory and (2) divide and/or multiply your"
                "`_xyz` dataset by multiples of 10 before running superpose again. Then, (3) revert that "
                "multiplication/division after the calculations.")
        self._rmsd_traces = None  # xyz changed (trajectory.py:1029)
        self.last_superpose_rmsd = out
        self.n_degenerate_rotations = degen
        if return_rotations:
            return self, rot
        return self


def rmsd_device(target: DeviceTrajectory, reference, frame=0, atom_indices=None, ref_atom_indices=None,
                precentered=False, superpose=True, as_numpy=True):
    """``md.rmsd`` for device-resident coordinates; same validation and warnings as ``_rmsd.pyx:154-213``."""
    torch = _torch()
    from ._rmsd import _ensure_int_1d
    ref_dev = reference if isinstance(reference, DeviceTrajectory) else DeviceTrajectory.from_trajectory(
        reference, target.device)
    none_idx = atom_indices is None
    if not none_idx:
        atom_indices = _ensure_int_1d(atom_indices, "atom_indices")
        if not np.all((atom_indices >= 0) * (atom_indices < target.n_atoms)):
            raise ValueError("atom_indices must be valid positive indices")
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    else:
        if none_idx:
            raise TypeError("object of type 'slice' has no len()")  # what the reference raises (Appendix B #13)
        if len(ref_atom_indices) != len(atom_indices):
            raise ValueError("atom_indices and ref_atom_indices must have same number of atom indices. "
                             "found %d and %d." % (len(atom_indices), len(ref_atom_indices)))
    if ref_atom_indices is not None:
        ref_atom_indices = _ensure_int_1d(ref_atom_indices, "ref_atom_indices")
        if not np.all((ref_atom_indices >= 0) * (ref_atom_indices < ref_dev.n_atoms)):
            raise ValueError("ref_atom_indices must be valid positive indices")
    if none_idx and target.n_atoms != ref_dev.n_atoms:
        raise ValueError("Input trajectories must have same number of atoms. "
                         "found %d and %d." % (target.n_atoms, ref_dev.n_atoms))
    if frame >= ref_dev.n_frames:
        raise ValueError("Cannot calculate RMSD of frame %d: reference has "
                         "only %d frames." % (frame, ref_dev.n_frames))
    dev = target.device
    F = target.n_frames
    L = _capi.lib()
    idx = None if none_idx else torch.from_numpy(atom_indices.astype(np.int32)).to(dev)
    ridx = None if ref_atom_indices is None else torch.from_numpy(ref_atom_indices.astype(np.int32)).to(dev)
    n_sel = target.n_atoms if none_idx else int(idx.numel())
    out = torch.zeros(F, dtype=torch.float32, device=dev)
    stream = _stream_ptr(torch, dev)
    ref_frame = ref_dev._xyz[frame]

    use_traces = False
    if superpose:
        if precentered and target._rmsd_traces is not None and ref_dev._rmsd_traces is not None and none_idx:
            use_traces = True
        elif precentered:
            warnings.warn("in rmsd(), precentered is ignored when atom_indices != None", RuntimeWarning)
    elif precentered:
        warnings.warn("in rmsd(), precentered is ignored when superpose=False", RuntimeWarning)
    if F == 0:
        return out.cpu().numpy() if as_numpy else out

    with torch.cuda.device(dev):
        if not superpose:
            prep = prepare_reference(ref_frame, ridx, n_sel, False)
            rc = L.b200rmsd_rmsd_nosuperpose_dev(target._xyz.data_ptr(), F, target.n_atoms, target.frame_stride,
                                                 None if idx is None else idx.data_ptr(), n_sel, prep.ref.data_ptr(),
                                                 out.data_ptr(), stream)
            _capi.check(rc, "b200rmsd_rmsd_nosuperpose_dev")
        else:
            if use_traces:
                prep = prepare_reference(ref_frame, None, n_sel, False, float(ref_dev._rmsd_traces[frame]))
            else:
                prep = prepare_reference(ref_frame, ridx, n_sel, True)
            nbytes = L.b200rmsd_scratch_bytes(F, target.n_atoms)
            scratch = _Scratch.get(torch, dev, nbytes)
            rc = L.b200rmsd_rmsd_dev(target._xyz.data_ptr(), F, target.n_atoms, target.frame_stride,
                                     None if idx is None else idx.data_ptr(), n_sel, prep.ref.data_ptr(),
                                     prep.stats.data_ptr(),
                                     target._rmsd_traces.data_ptr() if use_traces else None,
                                     _capi.PRECENTERED if use_traces else 0, out.data_ptr(), None, None, None,
                                     scratch.data_ptr(), scratch.numel(), stream)
            _capi.check(rc, "b200rmsd_rmsd_dev")
            if none_idx and ref_dev is target and -F <= frame < F:
                out[frame] = 0.0  # same-memory shortcut, theobald_rmsd_sse.h:256-262
    return out.cpu().numpy() if as_numpy else out


def superpose_raw_arrays(align_target_frame, align_mobile, g_target, g_mobile, displace):
    """Body of ``superpose_atom_major`` (_rmsd.pyx:663-674) for host arrays: rotation only, in place on ``displace``."""
    torch = _torch()
    from ._rmsd import current_device
    dev = torch.device("cuda", current_device())
    mob = DeviceTrajectory.from_host(align_mobile, dev, traces=g_mobile)
    dis = DeviceTrajectory.from_host(displace, dev)
    tgt = torch.from_numpy(np.ascontiguousarray(align_target_frame, dtype=np.float32)).to(dev)
    prep = prepare_reference(tgt, None, mob.n_atoms, False, float(g_target))
    F = mob.n_frames
    L = _capi.lib()
    out = torch.empty(F, dtype=torch.float32, device=dev)
    rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
    scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, max(mob.n_atoms, dis.n_atoms)))
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        rc = L.b200rmsd_rmsd_dev(mob._xyz.data_ptr(), F, mob.n_atoms, mob.frame_stride, None, mob.n_atoms,
                                 prep.ref.data_ptr(), prep.stats.data_ptr(), mob._rmsd_traces.data_ptr(),
                                 _capi.PRECENTERED, out.data_ptr(), rot.data_ptr(), None, None, scratch.data_ptr(),
                                 scratch.numel(), stream)
        _capi.check(rc, "b200rmsd_rmsd_dev")
        rc = L.b200rmsd_rotate_dev(dis._xyz.data_ptr(), F, dis.n_atoms, dis.frame_stride, rot.data_ptr(),
                                   scratch.data_ptr(), scratch.numel(), stream)
        _capi.check(rc, "b200rmsd_rotate_dev")
    displace[...] = dis.xyz
    return rot.cpu().numpy().reshape(F, 3, 3)


def align_displace_raw_arrays(align1_frame, align2, g1, g2, displ1_frame, displ2, n_align, n_displ):
    """Body of ``getMultipleAlignDisplaceRMSDs_atom_major`` (_rmsd.pyx:736-759) for host arrays.

    The reference aligns a = ``xyz_align1[frame]`` onto b = ``xyz_align2[i]`` (rotation R_i), then measures
    ``displ1[frame] . R_i`` against ``displ2[i]``.  The streaming kernel produces the rotation of frame i onto the
    single frame, i.e. R_i^T; ``b200rmsd_rot_msd_dev(transpose=1)`` undoes that and does the measuring.
    Returns (rmsds (F,), rotations (F,3,3)).
    """
    torch = _torch()
    from ._rmsd import current_device
    dev = torch.device("cuda", current_device())
    F, na_pad = align2.shape[0], align2.shape[1]
    nd_pad = displ2.shape[1]
    L = _capi.lib()
    a2 = torch.from_numpy(np.ascontiguousarray(align2, dtype=np.float32)).to(dev)
    d2 = torch.from_numpy(np.ascontiguousarray(displ2, dtype=np.float32)).to(dev)
    a1 = torch.from_numpy(np.ascontiguousarray(align1_frame, dtype=np.float32)).to(dev)
    d1 = torch.from_numpy(np.ascontiguousarray(displ1_frame, dtype=np.float32)).to(dev)
    tr2 = torch.from_numpy(np.ascontiguousarray(g2, dtype=np.float32)).to(dev)
    out = torch.zeros(F, dtype=torch.float32, device=dev)
    rot = torch.zeros((F, 9), dtype=torch.float32, device=dev)
    rot_t = torch.zeros((F, 9), dtype=torch.float32, device=dev)
    if F == 0:
        return out.cpu().numpy(), rot.cpu().numpy().reshape(0, 3, 3)
    prep = prepare_reference(a1, None, int(n_align), False, float(g1))
    scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, int(n_align)))
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        rc = L.b200rmsd_rmsd_dev(a2.data_ptr(), F, int(n_align), na_pad * 3, None, int(n_align), prep.ref.data_ptr(),
                                 prep.stats.data_ptr(), tr2.data_ptr(), _capi.PRECENTERED, out.data_ptr(),
                                 rot_t.data_ptr(), None, None, scratch.data_ptr(), scratch.numel(), stream)
        _capi.check(rc, "b200rmsd_rmsd_dev")
        rc = L.b200rmsd_rot_msd_dev(d1.data_ptr(), d2.data_ptr(), F, int(n_displ), nd_pad * 3, rot_t.data_ptr(), 1,
                                    rot.data_ptr(), out.data_ptr(), stream)
        _capi.check(rc, "b200rmsd_rot_msd_dev")
    return out.cpu().numpy(), rot.cpu().numpy().reshape(F, 3, 3)
