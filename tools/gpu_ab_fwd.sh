#!/bin/bash
# same-box A/B of recurrence kernel variants (run-to-run differences between boxes are ~3 %, more than the effects measured)
# variants are built by tools/build_variant.sh <name> lstm.cu -D...; usage: gpu_ab_fwd.sh name1 name2 ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for i in 1 2; do
  python tools/ab_probe.py "product"
  for v in "$@"; do
    ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_$v.so python tools/ab_probe.py "$v"
  done
done 2>&1 | grep variant | tee gpurun_out/ab_fwd.jsonl
