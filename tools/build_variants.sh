#!/bin/bash
# Development aid: compile variants of the one-vs-many kernels side by side into variants/*.so (git-ignored, shipped by
# gpurun) for tools/ovm_sweep.py --lib.  Usage: tools/build_variants.sh name:"-DOVM_PIVOT=0 ..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
SRC="capi.cu host_pipeline.cu one_vs_many.cu aux_kernels.cu frame_resident.cu allpairs.cu allpairs_refs.cu allpairs_tc144.cu cluster_ops.cu lprmsd.cu"
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  ( cd mdtraj_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC,-pthread \
      -ccbin /usr/bin/g++ -DB200RMSD_DEV_SWITCHES $flags -o ../../variants/$name.so $SRC 2>/dev/null && echo "built variants/$name.so" ) &
done
wait
