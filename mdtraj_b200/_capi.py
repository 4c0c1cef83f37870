"""ctypes binding of libb200rmsd.so (include/b200rmsd.h).

There is no CPU fallback: if the shared object is missing or a call fails the
error is raised, never papered over.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200rmsd.so")

OK = 0
EINVAL, ECUDA, ENODEVICE, ENOMEM = -1, -2, -3, -4
PRECENTERED = 1
REFSTATS_BYTES = 56

# every symbol include/b200rmsd.h declares: (restype, argtypes)
_vp, _i64, _i32, _f32, _u32, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_uint, C.c_size_t
SIGNATURES = {
    "b200rmsd_abi_version": (_i32, []),
    "b200rmsd_last_error": (C.c_char_p, []),
    "b200rmsd_device_info": (_i32, [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_sz), C.POINTER(_i32),
                                    C.POINTER(_i32)]),
    "b200rmsd_scratch_bytes": (_sz, [_i64, _i32]),
    "b200rmsd_center_trace_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp]),
    "b200rmsd_prepare_reference_dev": (_i32, [_vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp]),
    "b200rmsd_rmsd_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _sz,
                                 _vp]),
    "b200rmsd_rmsd_nosuperpose_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp]),
    "b200rmsd_superpose_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "b200rmsd_rotate_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _vp, _sz, _vp]),
    "b200rmsd_rmsf_scratch_bytes": (_sz, [_i64, _i32]),
    "b200rmsd_rmsf_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _sz, _vp, _vp]),
    "b200rmsd_rot_msd_dev": (_i32, [_vp, _vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _vp]),
    "b200rmsd_rmsd_host": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _f32, _vp, _i32]),
    "b200rmsd_superpose_host": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp, C.POINTER(_u32), _i32]),
    "b200rmsd_center_host": (_i32, [_vp, _i64, _i32, _vp, _i32]),
    "b200rmsd_rmsd_host_multi": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _i32]),
    "b200rmsd_superpose_host_multi": (_i32, [_vp, _i64, _i32, _vp, _i32, _vp, _vp, _i32, _vp, _vp, C.POINTER(_u32), _vp,
                                             _i32]),
    "b200rmsd_center_host_multi": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32]),
    "b200rmsd_host_configure": (_i32, [_i32, _i32, _i32]),
    "b200rmsd_host_configure_staging": (_i32, [_i32]),
    "b200rmsd_release_workspaces": (None, []),
    "b200rmsd_allpairs_workspace_bytes": (_sz, [_i64, _i32]),
    "b200rmsd_allpairs_configure": (_i32, [_i32, _i32]),
    "b200rmsd_allpairs_set_cta_pair": (_i32, [_i32]),
    "b200rmsd_allpairs_info_dev": (_i32, [_vp, _sz, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_f32), _vp, _i32, _vp]),
    "b200rmsd_lprmsd_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b200rmsd_allpairs_prepare_dev": (_i32, [_vp, _i64, _i32, _i64, _vp, _i32, _vp, _sz, _vp]),
    "b200rmsd_allpairs_block_dev": (_i32, [_vp, _sz, _i64, _i32, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _i64, _u32, _vp]),
    "b200rmsd_allpairs_block_rot_dev": (_i32, [_vp, _sz, _i64, _i32, _i64, _i64, _i64, _i64, _vp, _i64, _vp, _u32, _vp]),
    "b200rmsd_allpairs_rows_dev": (_i32, [_vp, _sz, _i64, _i32, _i64, _i64, _vp, _i64, _u32, _vp]),
    "b200rmsd_peer_alloc": (_i32, [_sz, C.POINTER(_vp), _vp]),
    "b200rmsd_peer_open": (_i32, [_vp, C.POINTER(_vp)]),
    "b200rmsd_peer_close": (_i32, [_vp]),
    "b200rmsd_peer_free": (_i32, [_vp]),
    "b200rmsd_matrix_moments_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp]),
    "b200rmsd_exp_rowsum_dev": (_i32, [_vp, _i64, _i64, _i64, _f32, _i32, _vp, _vp]),
    "b200rmsd_row_argmin_dev": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "b200rmsd_condense_dev": (_i32, [_vp, _i64, _i64, _vp, _vp]),
}


class B200RMSDError(RuntimeError):
    def __init__(self, code: int, where: str, msg: str):
        super().__init__(f"{where} failed with code {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load (once) and return the shared library with typed signatures."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing. mdtraj_b200 has no CPU fallback: build the CUDA library first "
                "(python -m mdtraj_b200.build, or __graft_entry__.build())."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here == ABI drift, fail loudly
            fn.restype = res
            fn.argtypes = args
        if L.b200rmsd_abi_version() != 2:
            raise ImportError("libb200rmsd.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, where: str) -> None:
    if rc != OK:
        raise B200RMSDError(rc, where, lib().b200rmsd_last_error().decode("utf-8", "replace"))


def device_info(device: int = 0):
    n, sm, cc1, cc2, mem = _i32(0), _i32(0), _i32(0), _i32(0), _sz(0)
    check(lib().b200rmsd_device_info(device, C.byref(n), C.byref(sm), C.byref(mem), C.byref(cc1), C.byref(cc2)),
          "b200rmsd_device_info")
    return {"n_devices": n.value, "sm_count": sm.value, "hbm_bytes": mem.value, "cc": (cc1.value, cc2.value)}


def np_ptr(a):
    """void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data
