"""``md.lprmsd`` on the GPU (last "next" row; reference: ``mdtraj/rmsd/_lprmsd.pyx:71-221``).

LP-RMSD minimises the RMSD over rotation/translation *and* over the labels of exchangeable atoms (waters, identical
ligands).  Same signature, validation and three-step procedure as the reference: (1) rotate every frame onto the reference
with the rotation optimal for the distinguishable atoms, (2) with that orientation fixed solve the assignment problem of
every permute group on squared distances, (3) QCP RMSD of the relabelled selection.  One warp per frame does all three in
shared memory (``csrc/lprmsd.cu``); the reference runs Munkres on a dense ``n x n`` float64 matrix per frame
(``euclidean_permutation.cpp``, ~0.7 s per frame at 300 atoms).

Deliberate difference: the assignment is found by an exact shortest-augmenting-path solver per group instead of Munkres
on the whole matrix.  Both return a minimum-cost matching, so they agree whenever that matching is unique (always, for
real coordinates); for exactly tied costs either optimum may come out (same cost, RMSD equal to rounding).
"""
from __future__ import annotations

import numpy as np

from . import _capi
from ._rmsd import current_device


def _unique_int_1d(val, name):
    """ensure_type(np.unique(val), dtype=int, ndim=1, warn_on_cast=False) as called at _lprmsd.pyx:234,260."""
    val = np.unique(np.asarray(val))
    if val.dtype == object or val.dtype.kind in "US":
        raise TypeError(f"{name} must be numeric array-like, got dtype {val.dtype}")
    if val.ndim != 1:
        raise ValueError(f"{name} must be ndim 1. You supplied {val.ndim}")
    return np.ascontiguousarray(val.astype(np.int64))


def _validate_atom_indices(atom_indices, n_atoms):   # _lprmsd.pyx:230-241
    if atom_indices is None:
        return np.arange(n_atoms, dtype=np.int64)
    atom_indices = _unique_int_1d(atom_indices, "atom_indices")
    if not np.all((atom_indices >= 0) * (atom_indices < n_atoms)):
        raise ValueError("atom_indices must be valid positive indices")
    return atom_indices


def _validate_permute_groups(permute_groups, atom_indices):   # _lprmsd.pyx:256-273
    if permute_groups is None:
        return [atom_indices]
    permute_groups = [_unique_int_1d(group, "permute_groups[%d]" % i) for i, group in enumerate(permute_groups)]
    for pgroup in permute_groups:
        if len(np.setdiff1d(pgroup, atom_indices)) > 0:
            raise ValueError("The elements in each permute group must be a subset of atom_indices")
    all_permutable_atoms = np.concatenate(permute_groups) if permute_groups else np.zeros(0, dtype=np.int64)
    # the reference compares np.unique(..) == np.sort(..) elementwise, which numpy >= 1.25 refuses for arrays of different
    # length ("operands could not be broadcast") before the intended message is reached; the message is what is kept
    if len(np.unique(all_permutable_atoms)) != len(all_permutable_atoms):
        raise ValueError("permute_groups must be mutually disjoint sets")
    return permute_groups


def lprmsd(target, reference, frame=0, atom_indices=None, permute_groups=None, parallel=True, superpose=False,
           return_mapping=False):
    """``md.lprmsd(target, reference, frame=0, atom_indices=None, permute_groups=None, parallel=True, superpose=False)``.

    ``target`` / ``reference``: host trajectories (anything with ``.xyz``) or ``DeviceTrajectory``.  Returns the (F,)
    float32 LP-RMSDs; with ``superpose=True`` ``target`` is centred on ALL its atoms and rotated in place like the
    reference does (``_lprmsd.pyx:217-220``: no re-translation onto the reference).  ``return_mapping=True`` (an
    extension) also returns the (F, n_sel) int32 matching: reference atom ``atom_indices[i]`` is paired with target atom
    ``atom_indices[mapping[f, i]]``."""
    from .device import DeviceTrajectory, _Scratch, _stream_ptr, _torch
    frame = int(frame)
    t_is_dev = getattr(target, "_is_b200_device_trajectory", False)
    r_is_dev = getattr(reference, "_is_b200_device_trajectory", False)
    n_atoms_t = target.n_atoms if t_is_dev else np.asarray(target.xyz).shape[1]
    n_atoms_r = reference.n_atoms if r_is_dev else np.asarray(reference.xyz).shape[1]
    n_frames_r = reference.n_frames if r_is_dev else np.asarray(reference.xyz).shape[0]
    if n_atoms_t != n_atoms_r:   # _validate_shapes, _lprmsd.pyx:244-253
        raise ValueError("Input trajectories must have same number of atoms. "
                         "found %d and %d." % (n_atoms_t, n_atoms_r))
    if frame >= n_frames_r:
        raise ValueError("Cannot calculate RMSD of frame %d: reference has "
                         "only %d frames." % (frame, n_frames_r))
    atom_indices = _validate_atom_indices(atom_indices, n_atoms_t)
    permute_groups = _validate_permute_groups(permute_groups, atom_indices)
    if len(atom_indices) == 0:
        raise ValueError("Number of atom indices must be greater than 0")

    # positions of the permutable atoms inside the selection (_lprmsd.pyx:137-141); everything else is distinguishable
    groups_rel = [np.searchsorted(atom_indices, g).astype(np.int32) for g in permute_groups if len(g)]
    flat = np.concatenate(groups_rel).astype(np.int32) if groups_rel else np.zeros(0, dtype=np.int32)
    offs = np.zeros(len(groups_rel) + 1, dtype=np.int32)
    np.cumsum([len(g) for g in groups_rel], out=offs[1:])
    dis = np.setdiff1d(np.arange(len(atom_indices)), flat).astype(np.int32)
    g_max = max([len(g) for g in groups_rel], default=0)

    n_frames_t = target.n_frames if t_is_dev else np.asarray(target.xyz).shape[0]
    if n_frames_t == 0:
        empty = np.zeros(0, dtype=np.float32)
        return (empty, np.zeros((0, len(atom_indices)), dtype=np.int32)) if return_mapping else empty
    torch = _torch()
    dev = target.device if t_is_dev else torch.device("cuda", current_device())
    dt = target if t_is_dev else DeviceTrajectory.from_trajectory(target, dev)
    F, n_sel = dt.n_frames, len(atom_indices)
    if r_is_dev:
        ref_sel = reference.xyz_dev[frame].to(dev)[torch.from_numpy(atom_indices).to(dev)].contiguous()
    else:
        ref_sel = torch.from_numpy(np.ascontiguousarray(np.asarray(reference.xyz)[frame, atom_indices, :],
                                                        dtype=np.float32)).to(dev)
    all_atoms = n_sel == n_atoms_t   # sorted unique indices covering every atom: the identity selection
    idx = None if all_atoms else torch.from_numpy(atom_indices.astype(np.int32)).to(dev)

    def to_dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev) if len(a) else None
    dis_d, flat_d, offs_d = to_dev(dis), to_dev(flat), torch.from_numpy(offs).to(dev)
    out = torch.empty(F, dtype=torch.float32, device=dev)
    rot = torch.empty((F, 9), dtype=torch.float32, device=dev) if superpose else None
    mapping = torch.empty((F, n_sel), dtype=torch.int32, device=dev) if return_mapping else None
    L = _capi.lib()
    with torch.cuda.device(dev):
        stream = _stream_ptr(torch, dev)
        rc = L.b200rmsd_lprmsd_dev(dt.xyz_dev.data_ptr(), F, dt.n_atoms, dt.frame_stride,
                                   None if idx is None else idx.data_ptr(), n_sel, ref_sel.data_ptr(),
                                   None if dis_d is None else dis_d.data_ptr(), len(dis),
                                   None if flat_d is None else flat_d.data_ptr(), offs_d.data_ptr(), len(groups_rel),
                                   int(g_max), out.data_ptr(), None if rot is None else rot.data_ptr(),
                                   None if mapping is None else mapping.data_ptr(), stream)
        _capi.check(rc, "b200rmsd_lprmsd_dev")
        if superpose and F > 0:   # centre ALL atoms, then rot1 . rot2 (_lprmsd.pyx:217-220)
            rc = L.b200rmsd_center_trace_dev(dt.xyz_dev.data_ptr(), F, dt.n_atoms, dt.frame_stride, None, stream)
            _capi.check(rc, "b200rmsd_center_trace_dev")
            scratch = _Scratch.get(torch, dev, L.b200rmsd_scratch_bytes(F, dt.n_atoms))
            rc = L.b200rmsd_rotate_dev(dt.xyz_dev.data_ptr(), F, dt.n_atoms, dt.frame_stride, rot.data_ptr(),
                                       scratch.data_ptr(), scratch.numel(), stream)
            _capi.check(rc, "b200rmsd_rotate_dev")
            if t_is_dev:
                dt._rmsd_traces = None
    distances = out.cpu().numpy()
    if superpose and not t_is_dev:
        target.xyz = dt.xyz
    if return_mapping:
        return distances, mapping.cpu().numpy()
    return distances
