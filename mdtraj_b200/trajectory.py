"""Minimal host-side ``Trajectory``: exactly the slice of ``mdtraj.Trajectory`` that the
RMSD path reads or writes (``mdtraj/core/trajectory.py``):

    xyz property + setter (casts to float32, resets ``_rmsd_traces``)   :1000-1029
    ``_rmsd_traces`` cache                                              :1376-1384
    ``superpose``                                                       :1083-1173
    ``center_coordinates``                                              :2114-2137
    slicing (``__getitem__`` / ``slice``)                               :1296-1359

Topology, unit cells, time and I/O are out of scope;
``topology`` is carried opaquely so a real ``mdtraj.Topology`` can ride along.
``mdtraj_b200.patch_mdtraj()`` installs the same methods on a real ``mdtraj.Trajectory``.
"""
from __future__ import annotations

import numpy as np

from . import _capi
from . import _rmsd


def _coerce_xyz(value):
    """ensure_type(value, float32, 3, 'xyz', shape=(None,None,3), add_newaxis_on_deficient_ndim=True)
    as the setter at trajectory.py:1019-1028 does."""
    value = np.asarray(value)
    if value.ndim == 2:
        value = value[np.newaxis]
    if value.ndim != 3:
        raise ValueError("xyz must be ndim 3. You supplied %d" % value.ndim)
    if value.shape[2] != 3:
        raise ValueError("xyz must be shape (Any, Any, 3). You supplied  %s" % (value.shape,))
    if value.dtype != np.float32 or not value.flags.c_contiguous:
        value = np.ascontiguousarray(value, dtype=np.float32)
    return value


def superpose_host(self, reference, frame=0, atom_indices=None, ref_atom_indices=None, parallel=True):
    """Superpose each conformation in this trajectory upon a reference
    (``Trajectory.superpose``, trajectory.py:1083-1173).  Returns ``self``.

    The whole numpy/C sequence of the reference -- gather, float64 centroids, shifts,
    traces, QCP rotation, rotate, re-translate -- is one call into the CUDA library.
    """
    if atom_indices is None:
        idx = None
    elif len(atom_indices) == 0:
        raise ValueError("Number of atom indices must be greater than 0")
    else:
        idx = np.asarray(atom_indices)
    if ref_atom_indices is None:
        ridx = idx
    else:
        ridx = np.asarray(ref_atom_indices)
        if idx is not None and len(ridx) != len(idx):
            raise ValueError("Number of atoms must be consistent!")

    xyz = self.xyz
    rxyz = np.asarray(reference.xyz)
    n_frames, n_atoms = xyz.shape[0], xyz.shape[1]
    if idx is None and ridx is not None:
        idx = np.arange(len(ridx))
    if idx is None and rxyz.shape[1] != n_atoms:
        raise ValueError("operands could not be broadcast together: %d vs %d atoms" % (n_atoms, rxyz.shape[1]))
    for name, arr, n in (("atom_indices", idx, n_atoms), ("ref_atom_indices", ridx, rxyz.shape[1])):
        if arr is not None and arr.size and (arr.min() < -n or arr.max() >= n):
            raise IndexError("%s out of bounds for axis with size %d" % (name, n))
    if idx is not None:
        idx = np.ascontiguousarray(np.where(idx < 0, idx + n_atoms, idx), dtype=np.int32)
        ridx = np.ascontiguousarray(np.where(ridx < 0, ridx + rxyz.shape[1], ridx), dtype=np.int32)

    work = xyz if (xyz.dtype == np.float32 and xyz.flags.c_contiguous and xyz.flags.writeable) else \
        np.array(xyz, dtype=np.float32, order="C", copy=True)
    ref_frame = np.array(rxyz[frame], dtype=np.float32, order="C", copy=True)  # reference never mutated (:1129)
    degen = _capi.C.c_uint(0)
    if n_frames:
        devs = _rmsd.current_devices()
        rc = _capi.lib().b200rmsd_superpose_host_multi(
            work.ctypes.data, n_frames, n_atoms, ref_frame.ctypes.data, ref_frame.shape[0], _capi.np_ptr(idx),
            _capi.np_ptr(ridx), 0 if idx is None else len(idx), None, None, _capi.C.byref(degen),
            devs.ctypes.data, len(devs))
        _capi.check(rc, "b200rmsd_superpose_host_multi")
        # The guard of trajectory.py:1162-1169 fires when the ROTATED, CENTRED first frame is all zeros, i.e. before the
        # reference centroid is added back (:1171).  The kernel has already re-translated, so the same condition reads:
        # every atom of frame 0 sits on one point (x' = 0 . R + o).
        if not np.any(work[0] - work[0, :1]):
            raise OverflowError(
                "Encounted a potential overflow/underflow error during superpose() due to the magnitude of your `_xyz`"
                "coordinates. To circumvent this, (1) reload your trajectory and (2) divide and/or multiply your"
                "`_xyz` dataset by multiples of 10 before running superpose again. Then, (3) revert that "
                "multiplication/division after the calculations.")
    self.xyz = work  # rebinding resets _rmsd_traces (:1029, :1172)
    self._n_degenerate_rotations = int(degen.value)
    return self


def center_coordinates_host(self, mass_weighted=False):
    """Center each trajectory frame at the origin (``Trajectory.center_coordinates``, trajectory.py:2114-2137)."""
    if mass_weighted and getattr(self, "top", None) is not None:
        raise NotImplementedError("mass-weighted centring needs mdtraj's topology/masses; outside the RMSD hot path")
    self._rmsd_traces = _rmsd._center_inplace_atom_major(self._xyz)
    return self


class Trajectory:
    """Host container duck-typing ``mdtraj.Trajectory`` for the RMSD path."""

    def __init__(self, xyz, topology=None, **_ignored):
        self.topology = topology
        self.xyz = xyz
        self._rmsd_traces = None

    @property
    def top(self):
        return self.topology

    @property
    def xyz(self):
        return self._xyz

    @xyz.setter
    def xyz(self, value):
        self._xyz = _coerce_xyz(value)
        self._rmsd_traces = None

    @property
    def n_frames(self):
        return self._xyz.shape[0]

    @property
    def n_atoms(self):
        return self._xyz.shape[1]

    def __len__(self):
        return self.n_frames

    def __getitem__(self, key):
        return self.slice(key)

    def slice(self, key, copy=True):
        """``Trajectory.slice`` (trajectory.py:1296-1359).  Unlike upstream (:1333-1334, which hands the
        *unsliced* traces to the new object -- SURVEY.md section 3.3) the cached traces are sliced with the
        frames, so ``precentered=True`` on a sliced trajectory reads the right trace."""
        if isinstance(key, (int, np.integer)):
            key = [int(key)]
        xyz = self._xyz[key]
        if copy:
            xyz = xyz.copy()
        new = Trajectory(xyz, self.topology)
        if self._rmsd_traces is not None:
            new._rmsd_traces = np.asarray(self._rmsd_traces)[key].copy()
        return new

    superpose = superpose_host
    center_coordinates = center_coordinates_host

    def to_device(self, device=None):
        from .device import DeviceTrajectory
        return DeviceTrajectory.from_trajectory(self, device)

    def __repr__(self):
        return "<mdtraj_b200.Trajectory with %d frames, %d atoms>" % (self.n_frames, self.n_atoms)
