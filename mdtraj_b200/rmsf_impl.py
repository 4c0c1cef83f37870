"""``md.rmsf`` on the GPU (SURVEY.md section 8(f) "next" #1; reference: ``mdtraj/rmsd/_rmsd.pyx:247-484``).

Same signature, validation, warnings and quirks as the reference:

* ``reference=None`` means "already aligned": no rotation, but the frames are still centred when no index
  list is given (the reference centres ``target.xyz`` in place through a view, ``:344,:401``);
* with an index list AND a reference the upstream result is not well defined: it hands a non-C-contiguous
  fancy-index copy (``np.array(target.xyz[:, atom_indices, :], copy=True)``, strides (12, 12*F, 4), ``:408``) to
  ``rot_atom_major`` (``:415``), which walks it as if it were contiguous and mixes frames.  We return what the
  all-atom path gives on the sliced trajectory -- ``md.rmsf(t.atom_slice(idx), ref.atom_slice(ref_idx), frame)`` --
  i.e. the RMSF of the properly superposed selection.  With ``reference=None`` (no rotation) the index-list path
  is well defined upstream (statistics of the raw selected coordinates, ``tests/test_rmsd.py:255-267``) and kept;
* ``mode='residue'`` needs a topology with ``atom(i).residue.index`` / ``element.mass`` (host-side glue, ``:463-483``).

Device work: one pass for the rotations/centroids (``b200rmsd_rmsd_dev`` with ``out_rot``/``out_centroid``) and
one pass for the per-atom statistics (``b200rmsd_rmsf_dev``, float64 sums; the reference accumulates in float32).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _capi
from ._rmsd import _ensure_int_1d, _state, current_device


def rmsf(target, reference, frame=0, atom_indices=None, ref_atom_indices=None, parallel=True, precentered=False,
         mode="atom"):
    from .device import DeviceTrajectory, _Scratch, _stream_ptr, _torch, prepare_reference
    frame = int(frame)
    prealigned = reference is None
    if prealigned:
        reference = target
    t_is_dev = getattr(target, "_is_b200_device_trajectory", False)
    n_atoms_t = target.n_atoms if t_is_dev else np.asarray(target.xyz).shape[1]
    r_is_dev = getattr(reference, "_is_b200_device_trajectory", False)
    n_atoms_r = reference.n_atoms if r_is_dev else np.asarray(reference.xyz).shape[1]
    n_frames_r = reference.n_frames if r_is_dev else np.asarray(reference.xyz).shape[0]

    none_idx = atom_indices is None
    if not none_idx:
        atom_indices = _ensure_int_1d(atom_indices, "atom_indices")
        if not np.all((atom_indices >= 0) * (atom_indices < n_atoms_t) * (atom_indices < n_atoms_r)):
            raise ValueError("atom_indices must be valid positive indices")
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    else:
        if none_idx:
            raise TypeError("object of type 'slice' has no len()")  # what the reference raises
        if len(ref_atom_indices) != len(atom_indices):
            raise ValueError("atom_indices and ref_atom_indices must have same number of atom indices. "
                             "found %d and %d." % (len(atom_indices), len(ref_atom_indices)))
    if ref_atom_indices is not None:
        ref_atom_indices = _ensure_int_1d(ref_atom_indices, "ref_atom_indices")
        if not np.all((ref_atom_indices >= 0) * (ref_atom_indices < n_atoms_t) * (ref_atom_indices < n_atoms_r)):
            raise ValueError("ref_atom_indices must be valid positive indices")
    if mode not in {"atom", "residue"}:
        raise ValueError('Mode must be one of "atom" or "residue". "%s" supplied.' % mode)
    if none_idx and n_atoms_t != n_atoms_r:
        raise ValueError("Input trajectories must have same number of atoms. "
                         "found %d and %d." % (n_atoms_t, n_atoms_r))
    if frame >= n_frames_r:
        raise ValueError("Cannot calculate RMSF of frame %d: reference has "
                         "only %d frames." % (frame, n_frames_r))

    torch = _torch()
    dev = target.device if t_is_dev else torch.device("cuda", current_device())
    dt = target if t_is_dev else DeviceTrajectory.from_trajectory(target, dev)
    ref_dev = dt if reference is target else (reference if r_is_dev else DeviceTrajectory.from_trajectory(reference, dev))
    F = dt.n_frames
    n_sel = dt.n_atoms if none_idx else len(atom_indices)
    idx = None if none_idx else torch.from_numpy(atom_indices.astype(np.int32)).to(dev)
    ridx = None if ref_atom_indices is None else torch.from_numpy(ref_atom_indices.astype(np.int32)).to(dev)

    t_tr, r_tr = getattr(target, "_rmsd_traces", None), getattr(reference, "_rmsd_traces", None)
    use_traces = bool(precentered and t_tr is not None and r_tr is not None and none_idx)
    if precentered and not use_traces:
        warnings.warn("in rmsd(), precentered is ignored when atom_indices != None", RuntimeWarning)

    L = _capi.lib()
    out = torch.zeros(n_sel, dtype=torch.float32, device=dev)
    if F == 0:
        return out.cpu().numpy()
    rot = torch.empty((F, 9), dtype=torch.float32, device=dev)
    cen = torch.empty((F, 3), dtype=torch.float64, device=dev)
    rms = torch.empty(F, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        ref_frame = ref_dev.xyz_dev[frame].clone()
        if use_traces:
            tr_dev = dt._rmsd_traces if t_is_dev and dt._rmsd_traces is not None else \
                torch.from_numpy(np.ascontiguousarray(t_tr, dtype=np.float32)).to(dev)
            r_trace = float(r_tr[frame]) if not hasattr(r_tr, "cpu") else float(r_tr[frame].item())
            prep = prepare_reference(ref_frame, None, n_sel, False, r_trace)
        else:
            tr_dev = None
            prep = prepare_reference(ref_frame, ridx, n_sel, True)
        scratch = _Scratch.get(torch, dev, max(L.b200rmsd_scratch_bytes(F, dt.n_atoms),
                                               L.b200rmsd_rmsf_scratch_bytes(F, n_sel)))
        rc = L.b200rmsd_rmsd_dev(dt.xyz_dev.data_ptr(), F, dt.n_atoms, dt.frame_stride,
                                 None if idx is None else idx.data_ptr(), n_sel, prep.ref.data_ptr(),
                                 prep.stats.data_ptr(), None if tr_dev is None else tr_dev.data_ptr(),
                                 _capi.PRECENTERED if use_traces else 0, rms.data_ptr(), rot.data_ptr(), cen.data_ptr(),
                                 None, scratch.data_ptr(), scratch.numel(), stream)
        _capi.check(rc, "b200rmsd_rmsd_dev")
        # all-atom path: frames were centred in place before the copy (:344,:408); index list + reference: centre the
        # selection too (see the module docstring); index list without reference: raw coordinates, as upstream
        centre_in_stats = (none_idx and not use_traces) or (not none_idx and not prealigned)
        rc = L.b200rmsd_rmsf_dev(dt.xyz_dev.data_ptr(), F, dt.n_atoms, dt.frame_stride,
                                 None if idx is None else idx.data_ptr(), n_sel,
                                 None if prealigned else rot.data_ptr(), cen.data_ptr() if centre_in_stats else None,
                                 scratch.data_ptr(), scratch.numel(), out.data_ptr(), stream)
        _capi.check(rc, "b200rmsd_rmsf_dev")
    fluct = out.cpu().numpy()

    if _state["inplace"] and none_idx and not use_traces and not t_is_dev:
        xyz = np.asarray(target.xyz)
        if xyz.dtype == np.float32 and xyz.flags.c_contiguous and xyz.flags.writeable:
            _capi.check(L.b200rmsd_center_host(xyz.ctypes.data, xyz.shape[0], xyz.shape[1], None, current_device()),
                        "b200rmsd_center_host")

    if mode == "atom":
        return fluct
    # mode == "residue": mass-weighted mean square fluctuation per residue (:463-483)
    top = getattr(target, "topology", None)
    if top is None:
        raise ValueError('mode="residue" needs target.topology (atoms with .residue.index and .element.mass)')
    sel = np.arange(n_sel) if none_idx else atom_indices
    res_idx = np.array([top.atom(int(i)).residue.index for i in sel], dtype=np.int64)
    masses = np.array([float(top.atom(int(i)).element.mass) for i in sel], dtype=np.float32)
    n_res = top.n_residues
    msf = (fluct.astype(np.float32) ** 2) * masses
    num = np.zeros(n_res, dtype=np.float32); den = np.zeros(n_res, dtype=np.float32)
    np.add.at(num, res_idx, msf); np.add.at(den, res_idx, masses)
    res = np.full(n_res, -1.0, dtype=np.float32)
    ok = den > 0
    res[ok] = np.sqrt(num[ok] / den[ok])
    return res
