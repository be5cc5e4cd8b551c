#!/bin/bash
# Builds libaslp_b200_<name>.so next to the product library with extra nvcc defines for ONE source file (A/B measurements
# of a kernel variant in the same gpurun call: ASLP_B200_CUDA_LIB=<path> python tools/perf_probe.py ...).
# usage: tools/build_variant.sh <name> <file.cu> <-DFLAG=...>
set -e
cd "$(dirname "$0")/../kaldi-aslp_b200"
name=$1; src=$2; shift 2
mkdir -p build/variant_$name
objs=""
for f in csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$f" == "csrc/$src" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I ../include "$@" -c $f -o build/variant_$name/$b.o
    objs="$objs build/variant_$name/$b.o"
  else
    objs="$objs build/$b.o"
  fi
done
nvcc -shared -o libaslp_b200_$name.so $objs -gencode arch=compute_100a,code=sm_100a -lnccl -Xlinker -rpath,/usr/local/cuda/lib64
echo built libaslp_b200_$name.so
