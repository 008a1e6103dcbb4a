#!/bin/bash
# Builds a tuning variant of the library into parallel-gps_b200/lib_var/<name>.so:
#   scripts/build_variant.sh <name> [-DPSSGP_... flags]
# Only the three scan translation units are rebuilt with the flags; the others are taken from ../build.
set -e
name=$1; shift
cd "$(dirname "$0")/../parallel-gps_b200/csrc"
mkdir -p ../lib_var ../build_var/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in filter smoother adjoint fused; do
  nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c -o ../build_var/$name/$f.o $f.cu &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc $ARCH --shared -o ../lib_var/$name.so ../build_var/$name/*.o ../build/api.o ../build/discretise.o ../build/generic.o ../build/host_utils.o
echo built lib_var/$name.so
