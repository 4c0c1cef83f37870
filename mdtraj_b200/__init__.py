"""mdtraj_b200 -- B200-native (sm_100a) implementation of MDTraj's RMSD hot path.

Drop-in for the path named in BASELINE.json: ``md.rmsd``, ``Trajectory.superpose``,
``Trajectory.center_coordinates`` and the all-pairs RMSD matrix; nothing else of
mdtraj is rebuilt.  Hand-written CUDA behind a C ABI (``include/b200rmsd.h``,
``libb200rmsd.so``); no CPU fallback, no Triton, no multi-backend dispatch.

    import mdtraj_b200 as mdb
    d = mdb.rmsd(traj, traj, 0)                    # host arrays, streamed through the GPU
    dt = mdb.DeviceTrajectory.from_trajectory(traj)  # staged to HBM once
    d = mdb.rmsd(dt, dt, 0)
    D = mdb.rmsd_matrix(dt)                        # all pairs
    mdb.patch_mdtraj()                             # make a real mdtraj use these kernels
"""
from ._rmsd import (TypeCastPerformanceWarning, _center_inplace_atom_major, current_device,  # noqa: F401
                    getMultipleAlignDisplaceRMSDs_atom_major, getMultipleRMSDs_atom_major,
                    getMultipleRMSDs_axis_major, rmsd, rmsf, set_device, set_devices, set_host_pipeline, set_inplace_centering,
                    superpose_atom_major)
from .trajectory import Trajectory  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # torch-dependent pieces are imported lazily so the host-array API works without importing torch
    if name in ("DeviceTrajectory", "rmsd_device", "prepare_reference"):
        from . import device
        return getattr(device, name)
    if name == "lprmsd":
        from .lprmsd_impl import lprmsd
        return lprmsd
    if name in ("rmsd_matrix", "rmsd_matrix_device"):
        from . import allpairs
        return getattr(allpairs, name)
    if name in ("rmsd_condensed", "similarity_scores", "centroid_index", "assign_to_leaders"):
        from . import clustering
        return getattr(clustering, name)
    if name == "distributed":
        import importlib
        return importlib.import_module(".distributed", __name__)
    raise AttributeError(name)


def patch_mdtraj():
    """Swap this implementation into an importable ``mdtraj``.

    Replaces the Cython module ``mdtraj._rmsd`` (``mdtraj/rmsd/_rmsd.pyx``), the two names ``mdtraj/__init__.py:118``
    re-exports from it (``mdtraj.rmsd``, ``mdtraj.rmsf``), ``mdtraj.lprmsd`` (``_lprmsd.pyx:71``) and ``Trajectory.superpose`` (trajectory.py:1083-1173, fused into
    one call here).  ``Trajectory.center_coordinates`` is left alone: upstream's unweighted branch (trajectory.py:2130-2135)
    looks ``mdtraj._rmsd._center_inplace_atom_major`` up at call time and so lands in the CUDA library, and its
    mass-weighted branch keeps working.  Everything else of mdtraj is untouched.  tests/test_reference_integration.py runs
    the reference's own tests against the result.
    """
    import sys

    import mdtraj  # noqa: F401  (raises ImportError if the reference is not installed)

    from . import _rmsd as ours
    from .trajectory import superpose_host

    sys.modules["mdtraj._rmsd"] = ours
    mdtraj._rmsd = ours
    mdtraj.rmsd = ours.rmsd
    mdtraj.rmsf = ours.rmsf
    try:   # md.lprmsd (mdtraj/__init__.py:117 imports the name from the Cython module mdtraj._lprmsd)
        from . import lprmsd_impl
        import mdtraj._lprmsd as ref_lp
        ref_lp.lprmsd = lprmsd_impl.lprmsd
        mdtraj.lprmsd = lprmsd_impl.lprmsd
    except ImportError:  # a reference build without the optional LP-RMSD extension
        pass
    mdtraj.Trajectory.superpose = superpose_host
    return mdtraj
