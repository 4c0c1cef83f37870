"""Drop-in mirror of the reference's ``mdtraj._rmsd`` extension module.

Same public names, argument order, validation, warnings and exceptions as
``mdtraj/rmsd/_rmsd.pyx``; the frame loops and the C kernels underneath are
replaced by calls into ``libb200rmsd.so`` (``include/b200rmsd.h``).

Inputs may be host objects (anything with ``.xyz`` float32 (F,N,3) and
``._rmsd_traces``: ``mdtraj.Trajectory`` or ``mdtraj_b200.Trajectory``) or
``mdtraj_b200.DeviceTrajectory`` objects whose coordinates already sit in HBM in
the padded atom-major layout.  Host inputs are streamed through the device per
call; device inputs are computed in place.  No CPU path exists.

One deliberate deviation: the reference centres
``target.xyz`` and ``reference.xyz[frame]`` *in place* as a side effect of
``rmsd(..., atom_indices=None)`` (``_rmsd.pyx:71,197-213``).  Results do not
depend on it; reproducing it costs a device->host copy of the whole trajectory,
so it is off by default and enabled with ``set_inplace_centering(True)``.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

from . import _capi

__all__ = [
    "rmsd", "rmsf", "_center_inplace_atom_major", "getMultipleRMSDs_atom_major", "getMultipleRMSDs_axis_major",
    "superpose_atom_major", "getMultipleAlignDisplaceRMSDs_atom_major", "set_inplace_centering", "set_device",
    "set_devices", "set_host_pipeline", "current_device", "current_devices", "TypeCastPerformanceWarning",
]


class TypeCastPerformanceWarning(RuntimeWarning):
    """Same role as mdtraj.utils.validation.TypeCastPerformanceWarning (validation.py:36)."""


_state = {"inplace": False, "device": None, "devices": None}


def set_inplace_centering(flag: bool) -> None:
    """Reproduce (True) or skip (False, default) the reference's in-place centring side effect of rmsd()."""
    _state["inplace"] = bool(flag)


def set_device(index) -> None:
    """CUDA device used by the host-array entry points (default: $LOCAL_RANK, else 0)."""
    _state["device"] = None if index is None else int(index)


def set_devices(indices) -> None:
    """CUDA devices the host-array entry points spread their frames over (``b200rmsd_*_host_multi``).  ``None`` restores
    the default: the one device of ``set_device`` / ``$MDTRAJ_B200_DEVICE`` / ``$LOCAL_RANK`` when any of those is set (one
    process per GPU under torchrun), otherwise every visible device."""
    _state["devices"] = None if indices is None else [int(i) for i in indices]


def set_host_pipeline(copy_threads=None, chunk_mb=None, staged_chunk_mb=None, stage_piece_kb=None) -> None:
    """Settings of the host-array pipeline (``b200rmsd_host_configure``): threads of the memcpy pool that stages pageable
    memory through page-locked buffers (before its first use), MB of coordinates per chunk for page-locked (default 64)
    and for pageable (default 16) caller memory.  ``stage_piece_kb`` (``b200rmsd_host_configure_staging``): KB per piece of
    the streamed staging of pageable memory (0 = automatic, the default; -1 = stage whole chunks)."""
    _capi.check(_capi.lib().b200rmsd_host_configure(int(copy_threads or 0), int(chunk_mb or 0), int(staged_chunk_mb or 0)),
                "b200rmsd_host_configure")
    if stage_piece_kb is not None:
        _capi.check(_capi.lib().b200rmsd_host_configure_staging(int(stage_piece_kb)), "b200rmsd_host_configure_staging")


def current_device() -> int:
    if _state["device"] is not None:
        return _state["device"]
    if _state["devices"]:
        return _state["devices"][0]
    return int(os.environ.get("MDTRAJ_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))


_n_visible = None


def current_devices():
    """Devices a host-array call runs on, as an int32 array (see ``set_devices``)."""
    global _n_visible
    if _state["devices"]:
        return np.asarray(_state["devices"], dtype=np.int32)
    if _state["device"] is not None or "MDTRAJ_B200_DEVICE" in os.environ or "LOCAL_RANK" in os.environ:
        return np.asarray([current_device()], dtype=np.int32)
    if _n_visible is None:
        _n_visible = max(1, int(_capi.device_info(0)["n_devices"]))
    return np.arange(_n_visible, dtype=np.int32)


# ---------------------------------------------------------------------------
# validation helpers (restating mdtraj/utils/validation.py:44-160 for the two uses here)
# ---------------------------------------------------------------------------
def _ensure_int_1d(val, name):
    """ensure_type(np.asarray(val), dtype=int, ndim=1, name=name) as called at _rmsd.pyx:158,171."""
    val = np.asarray(val)
    if val.dtype == object or val.dtype.kind in "US":
        raise TypeError(f"{name} must be numeric array-like, got dtype {val.dtype}")
    if val.dtype != np.int64:
        warnings.warn(
            f"Casting {name} dtype={val.dtype} to {np.dtype(np.int64)} ", TypeCastPerformanceWarning, stacklevel=3)
        val = val.astype(np.int64)
    if val.ndim != 1:
        raise ValueError(f"{name} must be ndim 1. You supplied {val.ndim}")
    return np.ascontiguousarray(val)


def _is_device(obj) -> bool:
    return getattr(obj, "_is_b200_device_trajectory", False)


def _host_xyz(traj, what):
    xyz = traj.xyz
    if not isinstance(xyz, np.ndarray):
        xyz = np.asarray(xyz)
    return xyz


def _as_f32c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------------------
# rmsd
# ---------------------------------------------------------------------------
def rmsd(target, reference, frame=0, atom_indices=None, ref_atom_indices=None, parallel=True, precentered=False,
         superpose=True):
    """rmsd(target, reference, frame=0, atom_indices=None, ref_atom_indices=None,
            parallel=True, precentered=False, superpose=True)

    Compute RMSD of all conformations in target to a reference conformation
    (signature and semantics of ``mdtraj.rmsd``, ``_rmsd.pyx:65-243``).

    ``parallel`` is accepted and ignored (every call runs on the GPU).  Returns a
    float32 ndarray of shape ``(target.n_frames,)``.
    """
    frame = int(frame)
    if _is_device(target):
        from .device import rmsd_device
        return rmsd_device(target, reference, frame, atom_indices, ref_atom_indices, precentered, superpose)

    txyz = _host_xyz(target, "target")
    rxyz = _host_xyz(reference, "reference")  # a DeviceTrajectory reference hands back a host copy

    atom_indices_is_none = atom_indices is None
    if atom_indices_is_none:
        atom_indices = slice(None)
    else:
        atom_indices = _ensure_int_1d(atom_indices, "atom_indices")
        if not np.all((atom_indices >= 0) * (atom_indices < txyz.shape[1])):
            raise ValueError("atom_indices must be valid positive indices")

    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    else:
        # like the reference, len(slice) raises TypeError when atom_indices is None (Appendix B #13)
        if len(ref_atom_indices) != len(atom_indices):
            raise ValueError("atom_indices and ref_atom_indices must have same number of atom indices. "
                             "found %d and %d." % (len(atom_indices), len(ref_atom_indices)))

    if not isinstance(ref_atom_indices, slice):
        ref_atom_indices = _ensure_int_1d(ref_atom_indices, "ref_atom_indices")
        if not np.all((ref_atom_indices >= 0) * (ref_atom_indices < rxyz.shape[1])):
            raise ValueError("ref_atom_indices must be valid positive indices")

    assert (txyz.ndim == 3) and (rxyz.ndim == 3) and (txyz.shape[2] == 3) and (rxyz.shape[2] == 3)
    if not ((txyz.shape[1] == rxyz.shape[1]) or (len(ref_atom_indices) == len(atom_indices))):
        raise ValueError("Input trajectories must have same number of atoms. "
                         "found %d and %d." % (txyz.shape[1], rxyz.shape[1]))
    if frame >= rxyz.shape[0]:
        raise ValueError("Cannot calculate RMSD of frame %d: reference has "
                         "only %d frames." % (frame, rxyz.shape[0]))

    n_frames = txyz.shape[0]
    # the reference types its views as writable memoryviews: read-only buffers raise on the view path
    # (tests/test_rmsd_memmap.py:24,50); with an index list a copy is made first and it works.
    if atom_indices_is_none and not txyz.flags.writeable:
        raise ValueError("buffer source array is read-only")
    if isinstance(ref_atom_indices, slice) and not rxyz.flags.writeable:
        raise ValueError("buffer source array is read-only")

    t32 = _as_f32c(txyz)
    ref_frame = _as_f32c(rxyz[frame])
    out = np.zeros(n_frames, dtype=np.float32)
    L = _capi.lib()
    devs = current_devices()

    use_traces = False
    traces = None
    ref_trace = 0.0
    if superpose:
        t_tr = getattr(target, "_rmsd_traces", None)
        r_tr = getattr(reference, "_rmsd_traces", None)
        if precentered and (r_tr is not None) and (t_tr is not None) and atom_indices_is_none:
            use_traces = True
            traces = _as_f32c(np.asarray(t_tr))
            ref_trace = float(np.asarray(r_tr)[frame])
        elif precentered:
            warnings.warn("in rmsd(), precentered is ignored when atom_indices != None", RuntimeWarning)
    elif precentered:
        warnings.warn("in rmsd(), precentered is ignored when superpose=False", RuntimeWarning)

    if n_frames == 0:
        return out
    idx = None if atom_indices_is_none else _i32(atom_indices)
    ridx = None
    if idx is not None:
        ridx = _i32(ref_atom_indices)
    elif not isinstance(ref_atom_indices, slice):
        # only ref_atom_indices given is impossible in the reference (len(slice) TypeError above)
        ridx = _i32(ref_atom_indices)
        idx = np.arange(len(ridx), dtype=np.int32)

    rc = L.b200rmsd_rmsd_host_multi(
        t32.ctypes.data, n_frames, t32.shape[1], ref_frame.ctypes.data, ref_frame.shape[0],
        _capi.np_ptr(idx), _capi.np_ptr(ridx), 0 if idx is None else len(idx), int(bool(superpose)),
        int(use_traces), _capi.np_ptr(traces), ref_trace, out.ctypes.data, devs.ctypes.data, len(devs))
    _capi.check(rc, "b200rmsd_rmsd_host_multi")

    if atom_indices_is_none and superpose and target is reference and 0 <= (frame % rxyz.shape[0]) < n_frames:
        # the reference's same-pointer shortcut (theobald_rmsd_sse.h:256-262): a frame against itself,
        # in the same memory, is exactly 0
        out[frame % rxyz.shape[0]] = 0.0

    if _state["inplace"] and superpose and not use_traces and atom_indices_is_none:
        # reproduce the documented side effect (_rmsd.pyx:71): centre the live arrays
        if txyz.dtype == np.float32 and txyz.flags.c_contiguous:
            _capi.check(L.b200rmsd_center_host_multi(txyz.ctypes.data, n_frames, txyz.shape[1], None, devs.ctypes.data,
                                                     len(devs)), "b200rmsd_center_host_multi")
        if rxyz is not txyz and rxyz.dtype == np.float32 and rxyz[frame].flags.c_contiguous:
            fr = rxyz[frame]
            _capi.check(L.b200rmsd_center_host(fr.ctypes.data, 1, fr.shape[0], None, int(devs[0])), "b200rmsd_center_host")
    return out


# ---------------------------------------------------------------------------
# centring
# ---------------------------------------------------------------------------
def _center_inplace_atom_major(xyz):
    """Centre ``xyz`` (F,N,3) float32 C-contiguous in place, return the traces (``_rmsd.pyx:487-491``)."""
    if xyz is None:
        raise TypeError("Argument 'xyz' must not be None")
    if not (isinstance(xyz, np.ndarray) and xyz.dtype == np.float32 and xyz.ndim == 3 and xyz.flags.c_contiguous):
        raise ValueError("Buffer dtype mismatch or not C-contiguous: expected float32 (n_frames, n_atoms, 3)")
    if not xyz.flags.writeable:
        raise ValueError("buffer source array is read-only")
    assert xyz.shape[2] == 3
    traces = np.empty(xyz.shape[0], dtype=np.float32)
    if xyz.shape[0] and xyz.shape[1]:
        devs = current_devices()
        _capi.check(_capi.lib().b200rmsd_center_host_multi(xyz.ctypes.data, xyz.shape[0], xyz.shape[1], traces.ctypes.data,
                                                           devs.ctypes.data, len(devs)), "b200rmsd_center_host_multi")
    return traces


# ---------------------------------------------------------------------------
# legacy low-level entry points on raw arrays
# ---------------------------------------------------------------------------
def _check_f32_3d(a, name):
    if a is None:
        raise TypeError(f"Argument '{name}' must not be None")
    if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 3 and a.flags.c_contiguous):
        raise ValueError(f"{name}: expected a C-contiguous float32 array with 3 dimensions")
    return a


def getMultipleRMSDs_atom_major(xyz1, xyz2, g1, g2, frame, parallel=True):
    """RMSD of every frame of ``xyz2`` to ``xyz1[frame]`` with caller-supplied traces (``_rmsd.pyx:562-615``)."""
    _check_f32_3d(xyz1, "xyz1"); _check_f32_3d(xyz2, "xyz2")
    assert xyz1.shape[2] == 3 and xyz2.shape[2] == 3
    if not (xyz1.shape[1] == xyz2.shape[1]):
        raise ValueError("Input arrays must have same number of atoms. "
                         "found %d and %d." % (xyz1.shape[1], xyz2.shape[1]))
    if frame >= xyz1.shape[0]:
        raise ValueError("Cannot calculate RMSD of frame %d: xyz1 has "
                         "only %d frames." % (frame, xyz1.shape[0]))
    g1 = _as_f32c(g1); g2 = _as_f32c(g2)
    out = np.zeros(xyz2.shape[0], dtype=np.float32)
    if xyz2.shape[0] == 0:
        return out
    ref = _as_f32c(xyz1[frame])
    rc = _capi.lib().b200rmsd_rmsd_host(xyz2.ctypes.data, xyz2.shape[0], xyz2.shape[1], ref.ctypes.data, ref.shape[0],
                                        None, None, 0, 1, 1, g2.ctypes.data, float(g1[frame]), out.ctypes.data,
                                        current_device())
    _capi.check(rc, "b200rmsd_rmsd_host")
    same_memory = (xyz1 is xyz2 and np.shares_memory(g1, g2)) or \
        (xyz1.ctypes.data == xyz2.ctypes.data and frame < g2.shape[0] and g1[frame] == g2[frame])
    if same_memory and -xyz2.shape[0] <= frame < xyz2.shape[0]:
        out[frame] = 0.0  # same-pointer shortcut, theobald_rmsd_sse.h:256-262
    return out


def getMultipleRMSDs_axis_major(xyz1, xyz2, g1, g2, frame, parallel=True):
    """Axis-major twin (``_rmsd.pyx:502-557``): arrays are (n_frames, 3, n_atoms)."""
    _check_f32_3d(xyz1, "xyz1"); _check_f32_3d(xyz2, "xyz2")
    assert xyz1.shape[1] == 3 and xyz2.shape[1] == 3
    if not (xyz1.shape[2] == xyz2.shape[2]):
        raise ValueError("Input arrays must have same number of atoms. "
                         "found %d and %d." % (xyz1.shape[2], xyz2.shape[2]))
    if frame >= xyz1.shape[0]:
        raise ValueError("Cannot calculate RMSD of frame %d: xyz1 has "
                         "only %d frames." % (frame, xyz1.shape[0]))
    # the device layout is atom-major; transpose once on the host (this legacy entry has no in-tree caller)
    a1 = np.ascontiguousarray(np.transpose(xyz1[frame:frame + 1], (0, 2, 1)))
    a2 = np.ascontiguousarray(np.transpose(xyz2, (0, 2, 1)))
    g1 = _as_f32c(g1)
    out = getMultipleRMSDs_atom_major(a1, a2, g1[frame:frame + 1], g2, 0)
    if xyz1.ctypes.data == xyz2.ctypes.data and frame < xyz2.shape[0] and \
            float(g1[frame]) == float(np.asarray(g2)[frame]):
        out[frame] = 0.0
    return out


def superpose_atom_major(xyz_align_target, xyz_align_mobile, g_target, g_mobile, xyz_displace_mobile, target_frame,
                         parallel=True):
    """Rotate every frame of ``xyz_displace_mobile`` in place by the rotation that best maps
    ``xyz_align_mobile[i]`` onto ``xyz_align_target[target_frame]`` (``_rmsd.pyx:620-674``).
    All arrays are expected centred, as the reference requires."""
    _check_f32_3d(xyz_align_target, "xyz_align_target"); _check_f32_3d(xyz_align_mobile, "xyz_align_mobile")
    _check_f32_3d(xyz_displace_mobile, "xyz_displace_mobile")
    if not (xyz_align_target.shape[1] == xyz_align_mobile.shape[1]):
        raise ValueError("Input arrays must have same number of atoms. "
                         "found %d and %d." % (xyz_align_target.shape[1], xyz_align_mobile.shape[1]))
    if not xyz_align_mobile.shape[0] == xyz_displace_mobile.shape[0]:
        raise ValueError("xyz_align_mobile and xyz_displace_mobile must contain the same number of frames")
    from .device import superpose_raw_arrays
    superpose_raw_arrays(xyz_align_target[target_frame], xyz_align_mobile, np.asarray(g_target)[target_frame],
                         g_mobile, xyz_displace_mobile)
    return None


def getMultipleAlignDisplaceRMSDs_atom_major(xyz_align1, xyz_align2, g_align1, g_align2, xyz_displ1, xyz_displ2,
                                             n_atoms_align, n_atoms_displ, frame, parallel=True):
    """Align on one atom set, measure on another (``_rmsd.pyx:679-759``).  Returns (rmsds, rotations)."""
    for a, n in ((xyz_align1, "xyz_align1"), (xyz_align2, "xyz_align2"), (xyz_displ1, "xyz_displ1"),
                 (xyz_displ2, "xyz_displ2")):
        _check_f32_3d(a, n)
    assert (xyz_align1.shape[2] == 3) and (xyz_align2.shape[2] == 3)
    assert (xyz_displ1.shape[2] == 3) and (xyz_displ2.shape[2] == 3)
    if not ((xyz_align1.shape[1] % 4 == 0) & (xyz_align2.shape[1] % 4 == 0)):
        raise ValueError("Input arrays must have middle dimension of 4*n, "
                         "found %d and %d." % (xyz_align1.shape[1], xyz_align2.shape[1]))
    if not ((xyz_displ1.shape[1] % 4 == 0) & (xyz_displ2.shape[1] % 4 == 0)):
        raise ValueError("Input arrays must have middle dimension of 4*n, "
                         "found %d and %d." % (xyz_displ1.shape[1], xyz_displ2.shape[1]))
    if not xyz_align1.shape[0] == xyz_displ1.shape[0]:
        raise ValueError("xyz_align1 and xyz_displ1 must contain the same number of frames")
    if not xyz_align2.shape[0] == xyz_displ2.shape[0]:
        raise ValueError("xyz_align2 and xyz_displ2 must contain the same number of frames")
    if frame >= xyz_align1.shape[0]:
        raise ValueError("Cannot calculate RMSD of frame %d: xyz1 has "
                         "only %d frames." % (frame, xyz_align1.shape[0]))
    from .device import align_displace_raw_arrays
    return align_displace_raw_arrays(xyz_align1[frame], xyz_align2, float(np.asarray(g_align1)[frame]), g_align2,
                                     xyz_displ1[frame], xyz_displ2, int(n_atoms_align), int(n_atoms_displ))


def rmsf(target, reference, frame=0, atom_indices=None, ref_atom_indices=None, parallel=True, precentered=False,
         mode="atom"):
    """Root-mean-square fluctuation per atom (``_rmsd.pyx:247-484``).  See mdtraj_b200.rmsf_impl."""
    from .rmsf_impl import rmsf as _rmsf
    return _rmsf(target, reference, frame, atom_indices, ref_atom_indices, parallel, precentered, mode)
