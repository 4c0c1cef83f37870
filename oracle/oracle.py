"""CPU oracle for the mdtraj RMSD hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  ``mdtraj_b200`` never does,
and has no CPU fallback.

Three independent routes to the same numbers:

``port``       ``liboracle.so`` -- our plain-C restatement (oracle/oracle.c) of
               ``mdtraj/rmsd/src/{center_generic.h,theobald_rmsd_generic.h,
               theobald_rmsd.cpp,rotation_generic.h}`` and the frame loops of
               ``mdtraj/rmsd/_rmsd.pyx``.
``reference``  ``_ref/libmdtraj_rmsd_ref.so`` -- the reference's own C++ files
               compiled in place from /root/reference (oracle/Makefile) behind
               our OpenMP loop shell (oracle/ref_loops.cpp).
``truth``      float64 numpy: Kabsch via SVD (``truth_rmsd``/``truth_superpose``),
               restating the idea of ``mdtraj/geometry/alignment.py:117-181`` --
               the reference's own test oracle -- independently.

The Python-level glue of ``md.rmsd`` (``_rmsd.pyx:154-243``) and
``Trajectory.superpose`` (``core/trajectory.py:1115-1173``) -- index selection,
float64 means, float32 einsum traces, re-translation -- is restated in
``rmsd`` and ``superpose`` below and runs on either C route.

Parity status: pinned (see oracle/oracle.c header and tests/test_oracle.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = C.POINTER(C.c_float)
_i64 = C.c_int64


def build(quiet: bool = True) -> None:
    """Compile liboracle.so, and _ref/ when /root/reference is present."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_f32p)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.oracle_qcp_largest_root.restype = C.c_double
        L.oracle_qcp_largest_root.argtypes = [C.c_double] * 3
        L.oracle_msd_from_M_and_G.restype = C.c_float
        L.oracle_msd_from_M_and_G.argtypes = [_f32p, C.c_float, C.c_float, C.c_int, C.c_int, _f32p,
                                              C.POINTER(C.c_int)]
        L.oracle_center_and_trace.restype = None
        L.oracle_center_and_trace.argtypes = [_f32p, _f32p, _i64, C.c_int]
        L.oracle_msd_atom_major.restype = C.c_float
        L.oracle_msd_atom_major.argtypes = [C.c_int, _f32p, _f32p, C.c_float, C.c_float, C.c_int, _f32p]
        L.oracle_msd_axis_major.restype = C.c_float
        L.oracle_msd_axis_major.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float]
        L.oracle_rot_atom_major.restype = None
        L.oracle_rot_atom_major.argtypes = [C.c_int, _f32p, _f32p]
        L.oracle_rot_msd_atom_major.restype = C.c_float
        L.oracle_rot_msd_atom_major.argtypes = [C.c_int, _f32p, _f32p, _f32p]
        L.oracle_rmsd_one_vs_many.restype = None
        L.oracle_rmsd_one_vs_many.argtypes = [_f32p, _f32p, _i64, C.c_int, _f32p, C.c_float, _f32p]
        L.oracle_rmsd_nosuperpose.restype = None
        L.oracle_rmsd_nosuperpose.argtypes = [_f32p, _i64, C.c_int, _f32p, _f32p]
        L.oracle_superpose_atom_major.restype = None
        L.oracle_superpose_atom_major.argtypes = [_f32p, C.c_float, _f32p, _f32p, _i64, C.c_int, _f32p, C.c_int,
                                                  _f32p]
        _port = L
    return _port


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libmdtraj_rmsd_ref.so"))


def ref_lib():
    """The compiled reference (raises FileNotFoundError when it was never built)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libmdtraj_rmsd_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        L.inplace_center_and_trace_atom_major.restype = None
        L.inplace_center_and_trace_atom_major.argtypes = [_f32p, _f32p, C.c_int, C.c_int]
        L.msd_atom_major.restype = C.c_float
        L.msd_atom_major.argtypes = [C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float, C.c_int, _f32p]
        L.msd_axis_major.restype = C.c_float
        L.msd_axis_major.argtypes = [C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_float, C.c_float]
        L.rot_atom_major.restype = None
        L.rot_atom_major.argtypes = [C.c_int, _f32p, _f32p]
        L.rot_msd_atom_major.restype = C.c_float
        L.rot_msd_atom_major.argtypes = [C.c_int, C.c_int, _f32p, _f32p, _f32p]
        L.refloops_max_threads.restype = C.c_int
        L.refloops_set_threads.argtypes = [C.c_int]
        L.refloops_rmsd.restype = None
        L.refloops_rmsd.argtypes = [_f32p, _i64, C.c_int, _f32p, C.c_int, _f32p, C.c_float, C.c_int, _f32p]
        L.refloops_superpose_atom_major.restype = None
        L.refloops_superpose_atom_major.argtypes = [_f32p, C.c_float, _f32p, _f32p, _i64, C.c_int, _f32p, C.c_int,
                                                    C.c_int, _f32p]
        L.refloops_align_displace.restype = None
        L.refloops_align_displace.argtypes = [_f32p, C.c_float, _f32p, _f32p, _f32p, _f32p, _i64, C.c_int, C.c_int,
                                              C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.refloops_rmsd_nosuperpose.restype = None
        L.refloops_rmsd_nosuperpose.argtypes = [_f32p, _i64, C.c_int, _f32p, C.c_int, _f32p]
        _ref = L
    return _ref


def _c32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------
# low-level pieces on either C route
# --------------------------------------------------------------------------
def center_and_trace(xyz: np.ndarray, impl: str = "port") -> np.ndarray:
    """In-place centring + traces (center.h:7).  ``xyz`` (F,N,3) float32 C-contiguous, mutated."""
    assert xyz.dtype == np.float32 and xyz.flags.c_contiguous and xyz.ndim == 3
    F, N, _ = xyz.shape
    tr = np.empty(F, dtype=np.float32)
    if impl == "port":
        port_lib().oracle_center_and_trace(_ptr(xyz), _ptr(tr), F, N)
    else:
        ref_lib().inplace_center_and_trace_atom_major(_ptr(xyz), _ptr(tr), F, N)
    return tr


def msd_atom_major(a, b, Ga, Gb, want_rot=False, impl="port"):
    a = _c32(a); b = _c32(b)
    rot = np.zeros(9, dtype=np.float32)
    if impl == "port":
        m = port_lib().oracle_msd_atom_major(a.shape[0], _ptr(a), _ptr(b), Ga, Gb, int(want_rot), _ptr(rot))
    else:
        m = ref_lib().msd_atom_major(a.shape[0], a.shape[0], _ptr(a), _ptr(b), Ga, Gb, int(want_rot), _ptr(rot))
    return (m, rot.reshape(3, 3)) if want_rot else m


def rot_atom_major(a, rot, impl="port"):
    """rot_atom_major (rotation.h:7): (n,3) float32 C-contiguous ``a`` <- a . rot, in place."""
    assert a.dtype == np.float32 and a.flags.c_contiguous
    rot = _c32(rot).reshape(9)
    if impl == "port":
        port_lib().oracle_rot_atom_major(a.shape[0], _ptr(a), _ptr(rot))
    else:
        ref_lib().rot_atom_major(a.shape[0], _ptr(a), _ptr(rot))
    return a


def one_vs_many_centered(target, target_g, ref_frame, ref_g, impl="port", parallel=True):
    """Loop of _rmsd.pyx:217-224 on already-centred float32 data."""
    target = _c32(target); ref_frame = _c32(ref_frame); target_g = _c32(target_g)
    F, N, _ = target.shape
    out = np.zeros(F, dtype=np.float32)
    if impl == "port":
        port_lib().oracle_rmsd_one_vs_many(_ptr(target), _ptr(target_g), F, N, _ptr(ref_frame), float(ref_g), _ptr(out))
    else:
        ref_lib().refloops_rmsd(_ptr(target), F, N, _ptr(ref_frame), 1, _ptr(target_g), float(ref_g), int(parallel),
                                _ptr(out))
    return out


# --------------------------------------------------------------------------
# md.rmsd semantics (mdtraj/rmsd/_rmsd.pyx:65-243) on raw arrays
# --------------------------------------------------------------------------
def rmsd(target_xyz, ref_xyz, frame=0, atom_indices=None, ref_atom_indices=None, superpose=True,
         target_traces=None, ref_traces=None, impl="port", parallel=True, inplace=False):
    """RMSD of every target frame to ``ref_xyz[frame]``.

    ``target_traces``/``ref_traces`` given (and no index lists) == the
    ``precentered=True`` fast path (:203-205).  ``inplace=False`` works on
    copies so the caller's arrays are not centred (the reference mutates the
    view path; pass ``inplace=True`` to reproduce that side effect).
    """
    target_xyz = np.asarray(target_xyz); ref_xyz = np.asarray(ref_xyz)
    assert target_xyz.dtype == np.float32 and ref_xyz.dtype == np.float32
    if atom_indices is None:
        t = target_xyz if inplace else target_xyz.copy()
        t = np.ascontiguousarray(t)
    else:
        t = np.ascontiguousarray(target_xyz[:, np.asarray(atom_indices, dtype=np.int64), :])
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    if ref_atom_indices is None:
        r = ref_xyz[frame] if (inplace and ref_xyz[frame].flags.c_contiguous) else ref_xyz[frame].copy()
    else:
        r = np.ascontiguousarray(ref_xyz[frame, np.asarray(ref_atom_indices, dtype=np.int64), :])
    F, N, _ = t.shape
    out = np.zeros(F, dtype=np.float32)
    if not superpose:
        if impl == "port":
            port_lib().oracle_rmsd_nosuperpose(_ptr(t), F, N, _ptr(r), _ptr(out))
        else:
            ref_lib().refloops_rmsd_nosuperpose(_ptr(t), F, N, _ptr(r), int(parallel), _ptr(out))
        return out
    use_tr = target_traces is not None and ref_traces is not None and atom_indices is None
    if use_tr:
        tg = _c32(target_traces); rg = float(ref_traces[frame])
    else:
        tg = None; rg = 0.0
    if impl == "port":
        if not use_tr:
            tg = center_and_trace(t, "port")
            rg = float(center_and_trace(r.reshape(1, N, 3), "port")[0])
        port_lib().oracle_rmsd_one_vs_many(_ptr(t), _ptr(tg), F, N, _ptr(r), rg, _ptr(out))
    else:
        ref_lib().refloops_rmsd(_ptr(t), F, N, _ptr(r), int(use_tr), _ptr(tg), rg, int(parallel), _ptr(out))
    return out


# --------------------------------------------------------------------------
# Trajectory.superpose semantics (mdtraj/core/trajectory.py:1115-1173) on raw arrays
# --------------------------------------------------------------------------
def superpose(xyz, ref_xyz, frame=0, atom_indices=None, ref_atom_indices=None, impl="port", parallel=True,
              return_rot=False):
    """Returns the superposed copy of ``xyz`` (F,N,3) float32 (and the rotations (F,3,3))."""
    xyz = np.array(xyz, dtype=np.float32, order="C", copy=True)
    F = xyz.shape[0]
    if atom_indices is None:
        align = xyz  # aliases, like the view at trajectory.py:1127
    else:
        atom_indices = np.asarray(atom_indices, dtype=np.int64)
        if len(atom_indices) == 0:
            raise ValueError("Number of atom indices must be greater than 0")
        align = np.ascontiguousarray(xyz[:, atom_indices, :])
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    ref_sel = slice(None) if ref_atom_indices is None else np.asarray(ref_atom_indices, dtype=np.int64)
    ref_align = np.array(ref_xyz[frame, ref_sel, :], dtype=np.float32, order="C", copy=True).reshape(1, -1, 3)

    offset = np.mean(align, axis=1, dtype=np.float64).reshape(F, 1, 3)
    align -= offset                       # float32 -= float64: computed in f64, rounded to f32
    if align is not xyz:
        xyz -= offset
    ref_offset = ref_align[0].astype("float64").mean(0)
    ref_align[0] -= ref_offset
    self_g = np.einsum("ijk,ijk->i", align, align)          # float32 einsum (:1149)
    ref_g = np.einsum("ijk,ijk->i", ref_align, ref_align)

    rot = np.zeros((F, 9), dtype=np.float32)
    n_align = align.shape[1]
    if impl == "port":
        port_lib().oracle_superpose_atom_major(_ptr(ref_align), float(ref_g[0]), _ptr(align), _ptr(self_g), F, n_align,
                                               _ptr(xyz), xyz.shape[1], _ptr(rot))
    else:
        ref_lib().refloops_superpose_atom_major(_ptr(ref_align), float(ref_g[0]), _ptr(align), _ptr(self_g), F,
                                                n_align, _ptr(xyz), xyz.shape[1], int(parallel), _ptr(rot))
    xyz += ref_offset
    return (xyz, rot.reshape(F, 3, 3)) if return_rot else xyz


# --------------------------------------------------------------------------
# float64 truth (independent of both C routes)
# --------------------------------------------------------------------------
# --------------------------------------------------------------------------
# md.lprmsd semantics (mdtraj/rmsd/_lprmsd.pyx:71-221) on raw arrays
# --------------------------------------------------------------------------
def min_cost_matching(cost):
    """Minimum-cost perfect matching of a square float64 cost matrix: row i -> column ``result[i]``.  Stands in for
    Munkres::solve (mdtraj/rmsd/src/Munkres.cpp; used at euclidean_permutation.cpp:56) -- both are exact, so they agree
    whenever the optimum is unique; pinned to the reference's own known answer (tests/test_lprmsd.py:12-22) and to
    md.lprmsd outputs in tests/test_oracle.py."""
    from scipy.optimize import linear_sum_assignment
    rows, cols = linear_sum_assignment(np.asarray(cost, dtype=np.float64))
    out = np.empty(len(rows), dtype=np.int64)
    out[rows] = cols
    return out


def euclidean_permutation(ref, tgt, groups):
    """euclidean_permutation.cpp:16-72 with its call-site roles (``_lprmsd.pyx:210``: rows = reference atoms, columns =
    target atoms): mapping[i] = target atom matched to reference atom i; atoms in no group map to themselves.  Costs are
    float32 differences and squares summed in float64 (lines 34-38).  Groups never interact (cross-group entries are
    DBL_MAX), so each is matched on its own."""
    ref = _c32(ref); tgt = _c32(tgt)
    mapping = np.arange(ref.shape[0], dtype=np.int64)
    for g in groups:
        g = np.asarray(g, dtype=np.int64)
        if len(g) < 2:
            continue
        d = (ref[g][:, None, :] - tgt[g][None, :, :]).astype(np.float32)
        cost = (d * d).astype(np.float32).astype(np.float64).sum(-1)
        mapping[g] = g[min_cost_matching(cost)]
    return mapping


def lprmsd(target_xyz, ref_xyz, frame=0, atom_indices=None, permute_groups=None, superpose=False, impl="port",
           return_mapping=False):
    """md.lprmsd on raw (F,N,3) arrays, step by step as _lprmsd.pyx:131-227 does it.  Returns the (F,) float32 distances
    (and, with ``superpose``, the modified copy of ``target_xyz``; with ``return_mapping`` the (F,n_sel) matchings)."""
    xyz = np.array(target_xyz, dtype=np.float32, order="C", copy=True)
    ref_xyz = np.asarray(ref_xyz)
    F, N, _ = xyz.shape
    atom_indices = np.arange(N) if atom_indices is None else np.unique(np.asarray(atom_indices, dtype=np.int64))
    groups = [atom_indices] if permute_groups is None else [np.unique(np.asarray(g, dtype=np.int64)) for g in permute_groups]
    groups_rel = [np.searchsorted(atom_indices, g) for g in groups]                                        # :139
    flat = np.concatenate(groups_rel) if groups_rel else np.zeros(0, dtype=np.int64)
    dis = np.setdiff1d(np.arange(len(atom_indices)), flat)                                                 # :140-141
    n = len(atom_indices)
    ref = np.array(ref_xyz[frame, atom_indices, :], dtype=np.float32, copy=True)[None]                    # :160
    ref_g = float(center_and_trace(ref, impl)[0])                                                          # :163
    if len(dis):
        ref_dis = np.array(ref_xyz[frame, atom_indices[dis], :], dtype=np.float32, copy=True, order="C")[None]   # :176-177
        ref_g_dis = float(center_and_trace(ref_dis, impl)[0])
    out = np.zeros(F, dtype=np.float32)
    maps = np.empty((F, n), dtype=np.int64)
    for i in range(F):
        t = np.ascontiguousarray(xyz[i, atom_indices, :])[None]                                            # :188-190
        t_g = float(center_and_trace(t, impl)[0])
        rot1 = np.eye(3, dtype=np.float32)
        if len(dis):
            t_dis = np.ascontiguousarray(t[0][dis])[None]                                                  # :197-199
            t_g_dis = float(center_and_trace(t_dis, impl)[0])
            _, rot1 = msd_atom_major(t_dis[0], ref_dis[0], t_g_dis, ref_g_dis, True, impl)                 # :202-204
            rot_atom_major(t[0], rot1, impl)                                                               # :207
        mapping = euclidean_permutation(ref[0], t[0], groups_rel)                                          # :210
        maps[i] = mapping
        t2 = np.ascontiguousarray(t[0][mapping])                                                           # :212
        if superpose:
            msd, rot2 = msd_atom_major(t2, ref[0], t_g, ref_g, True, impl)                                 # :217
            whole = np.ascontiguousarray(xyz[i])[None]
            center_and_trace(whole, impl)                                                                  # :218
            rot3 = np.zeros((3, 3), dtype=np.float32)                                                      # :219 sgemm33:
            for a in range(3):                                                                             # float32, k ascending
                for b in range(3):
                    o = np.float32(0)
                    for k in range(3):
                        o = np.float32(o + np.float32(rot1[a, k] * rot2[k, b]))
                    rot3[a, b] = o
            xyz[i] = rot_atom_major(whole[0], rot3, impl)                                                  # :220
        else:
            msd = msd_atom_major(ref[0], t2, t_g, ref_g, False, impl)                                      # :222
        out[i] = np.sqrt(np.float32(max(msd, 0.0)))
    res = (out, xyz) if superpose else out
    if return_mapping:
        return (res + (maps,)) if isinstance(res, tuple) else (res, maps)
    return res


def truth_lprmsd_given_mapping(target_xyz, ref_frame, atom_indices, mapping):
    """float64 Kabsch RMSD of the relabelled selection: what step 3 should give for a fixed matching."""
    atom_indices = np.asarray(atom_indices, dtype=np.int64)
    ref = np.asarray(ref_frame, dtype=np.float64)[atom_indices]
    out = np.empty(len(target_xyz))
    for i in range(len(target_xyz)):
        t = np.asarray(target_xyz[i], dtype=np.float64)[atom_indices][np.asarray(mapping[i], dtype=np.int64)]
        out[i] = truth_kabsch(t, ref)[1]
    return out


def truth_kabsch(mobile, target):
    """Optimal rotation R (3,3) with mobile_c @ R ~ target_c, and the RMSD, float64."""
    P = np.asarray(mobile, dtype=np.float64); Q = np.asarray(target, dtype=np.float64)
    Pc = P - P.mean(0); Qc = Q - Q.mean(0)
    H = Pc.T @ Qc
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    D = np.diag([1.0, 1.0, d])
    R = U @ D @ Vt
    e0 = (Pc * Pc).sum() + (Qc * Qc).sum()
    msd = max(0.0, (e0 - 2.0 * (S[0] + S[1] + d * S[2])) / P.shape[0])
    return R, float(np.sqrt(msd))


def truth_rmsd(target_xyz, ref_xyz, frame=0, atom_indices=None, ref_atom_indices=None, superpose=True):
    ti = slice(None) if atom_indices is None else np.asarray(atom_indices, dtype=np.int64)
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    ri = slice(None) if ref_atom_indices is None else np.asarray(ref_atom_indices, dtype=np.int64)
    ref = np.asarray(ref_xyz[frame, ri, :], dtype=np.float64)
    out = np.empty(target_xyz.shape[0], dtype=np.float64)
    for i in range(target_xyz.shape[0]):
        x = np.asarray(target_xyz[i, ti, :], dtype=np.float64)
        if superpose:
            out[i] = truth_kabsch(x, ref)[1]
        else:
            out[i] = np.sqrt(((x - ref) ** 2).sum() / x.shape[0])
    return out


def truth_rmsd_batch(target_xyz, ref_frame):
    """Vectorised float64 Kabsch RMSD of every frame of ``target_xyz`` (F,N,3) against ``ref_frame`` (N,3): the same
    numbers as ``truth_rmsd`` without the Python loop (used by bench.py's parity block on 10^4 frames)."""
    X = np.asarray(target_xyz, dtype=np.float64)
    Q = np.asarray(ref_frame, dtype=np.float64)
    Xc = X - X.mean(1, keepdims=True)
    Qc = Q - Q.mean(0)
    H = np.einsum("fki,kj->fij", Xc, Qc)
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    e0 = np.einsum("fki,fki->f", Xc, Xc) + (Qc * Qc).sum()
    msd = np.maximum(0.0, (e0 - 2.0 * (S[:, 0] + S[:, 1] + d * S[:, 2])) / X.shape[1])
    return np.sqrt(msd)


def truth_superpose(xyz, ref_xyz, frame=0, atom_indices=None, ref_atom_indices=None):
    """float64 version of Trajectory.superpose; returns (xyz', R (F,3,3))."""
    X = np.asarray(xyz, dtype=np.float64)
    ti = slice(None) if atom_indices is None else np.asarray(atom_indices, dtype=np.int64)
    if ref_atom_indices is None:
        ref_atom_indices = atom_indices
    ri = slice(None) if ref_atom_indices is None else np.asarray(ref_atom_indices, dtype=np.int64)
    ref = np.asarray(ref_xyz[frame, ri, :], dtype=np.float64)
    out = np.empty_like(X); rots = np.empty((X.shape[0], 3, 3))
    for i in range(X.shape[0]):
        sel = X[i, ti, :]
        R, _ = truth_kabsch(sel, ref)
        out[i] = (X[i] - sel.mean(0)) @ R + ref.mean(0)
        rots[i] = R
    return out, rots


# --------------------------------------------------------------------------
# seeded synthetic generators shared by tests and bench
# --------------------------------------------------------------------------
def synth_iid(n_frames, n_atoms, seed=0):
    """D-iid: xyz ~ N(0,1) nm float32 (what examples/rmsd-benchmark.ipynb:45 uses)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n_frames, n_atoms, 3), dtype=np.float32)


def random_rotations(n, rng):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    a, b, c, d = q.T
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = a * a + b * b - c * c - d * d; R[:, 0, 1] = 2 * (b * c - a * d); R[:, 0, 2] = 2 * (b * d + a * c)
    R[:, 1, 0] = 2 * (b * c + a * d); R[:, 1, 1] = a * a - b * b + c * c - d * d; R[:, 1, 2] = 2 * (c * d - a * b)
    R[:, 2, 0] = 2 * (b * d - a * c); R[:, 2, 1] = 2 * (c * d + a * b); R[:, 2, 2] = a * a - b * b - c * c + d * d
    return R


def synth_md(n_frames, n_atoms, seed=0, rg=1.5, sigma=0.1, box=5.0):
    """D-md: one base structure N(0, rg^2), per-frame noise sigma, random rigid motion, +-box nm offset."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_atoms, 3)) * rg
    X = base[None] + rng.standard_normal((n_frames, n_atoms, 3)) * sigma
    R = random_rotations(n_frames, rng)
    X = np.einsum("fni,fij->fnj", X, R) + rng.uniform(-box, box, size=(n_frames, 1, 3))
    return X.astype(np.float32)


def synth_md_chain(n_frames, n_atoms, seed=0, bond=0.15, sigma=0.1, box=5.0):
    """D-chain: a random-walk polymer (bond nm per step, Rg ~ bond*sqrt(N/6)) -- consecutive atoms are spatially correlated
    as in a real protein, so a pivot or centroid estimated from the first atoms of a frame is a poor one.  Per-frame noise,
    random rigid motion, +-box nm offset."""
    rng = np.random.default_rng(seed)
    steps = rng.standard_normal((n_atoms, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    base = np.cumsum(bond * steps, 0)
    base -= base.mean(0)
    X = base[None] + rng.standard_normal((n_frames, n_atoms, 3)) * sigma
    R = random_rotations(n_frames, rng)
    X = np.einsum("fni,fij->fnj", X, R) + rng.uniform(-box, box, size=(n_frames, 1, 3))
    return X.astype(np.float32)


def synth_md_basins(n_frames, n_atoms, n_basins=3, seed=0, rg=1.0, sigma=0.1, separation=1.4, interleave=False, box=5.0):
    """Multi-basin MD-like frames: `n_basins` base structures `separation` nm of RMSD apart from the first one (and
    ~separation*sqrt(2) from each other), per-frame noise sigma, random rigid motion.  Frames of a basin are contiguous
    (a trajectory that hops) unless interleave=True (concatenated / shuffled runs).  Returns (X float32, basin index)."""
    rng = np.random.default_rng(seed)
    base0 = rng.standard_normal((n_atoms, 3)) * rg
    bases = [base0] + [base0 + rng.standard_normal((n_atoms, 3)) * (separation / np.sqrt(3.0)) for _ in range(n_basins - 1)]
    which = rng.integers(0, n_basins, n_frames) if interleave else (np.arange(n_frames) * n_basins) // n_frames
    X = np.stack(bases)[which] + rng.standard_normal((n_frames, n_atoms, 3)) * sigma
    R = random_rotations(n_frames, rng)
    X = np.einsum("fni,fij->fnj", X, R) + rng.uniform(-box, box, size=(n_frames, 1, 3))
    return X.astype(np.float32), which


def synth_md_drift(n_frames, n_atoms, seed=0, rg=1.0, step=0.02, sigma=0.01, box=5.0):
    """A trajectory that drifts: per-atom random walk of `step` nm per frame on top of the base structure (RMSD to frame 0
    grows like step*sqrt(3 t)), neighbouring frames ~step*sqrt(3) apart, small per-frame noise, random rigid motion."""
    rng = np.random.default_rng(seed)
    base = rng.standard_normal((n_atoms, 3)) * rg
    X = base[None] + np.cumsum(step * rng.standard_normal((n_frames, n_atoms, 3)), 0) + \
        sigma * rng.standard_normal((n_frames, n_atoms, 3))
    R = random_rotations(n_frames, rng)
    X = np.einsum("fni,fij->fnj", X, R) + rng.uniform(-box, box, size=(n_frames, 1, 3))
    return X.astype(np.float32)
