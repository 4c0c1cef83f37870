"""In-tree build of libb200rmsd.so (sm_100a only).

    python -m mdtraj_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The shared object lands next to this file
(git-ignored, but shipped to the GPU box by gpurun) and is loaded with ctypes by
mdtraj_b200._capi.  CUDA runtime is linked statically so the library has no
dependency on torch's copy.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200rmsd.so")
SOURCES = ["capi.cu", "host_pipeline.cu", "one_vs_many.cu", "aux_kernels.cu", "frame_resident.cu", "allpairs.cu", "allpairs_refs.cu", "allpairs_tc144.cu", "cluster_ops.cu", "lprmsd.cu"]
HEADERS = ["host_pipeline.cu", "allpairs_refs.cu", "common.cuh", "kernels.cuh", "qcp.cuh", "allpairs_layout.cuh", "tc_ptx.cuh", "../../include/b200rmsd.h"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; mdtraj_b200 has no CPU fallback and cannot be built without CUDA")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    """dev=True adds -DB200RMSD_DEV_SWITCHES: the B200RMSD_* environment overrides (kernel / geometry selection, isolation
    runs) the sweeps under tools/ use.  A release build (the default, what the tests and the bench run) never reads the
    environment."""
    if not force and not dev and not needs_build():
        return LIB
    cmd = [nvcc_path(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-pthread",
           "-ccbin", "/usr/bin/g++", "-o", LIB, *sources()]
    if dev:
        cmd[1:1] = ["-DB200RMSD_DEV_SWITCHES"]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libb200rmsd.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, dev="--dev" in sys.argv))
