#!/bin/bash
# Same-box A/B of the default GEMM split mode at the level of the cfg3 training step: in-loop 3xTF32 split vs fp16 hi/lo planes.
# usage (under gpurun): bash tools/gpu_ab_gemm_step.sh > gpurun_out/ab_gemm_step.jsonl
for i in 1 2; do
  for v in tf32 f16; do
    if [ $v == tf32 ]; then export ASLP_GEMM_SPLIT=tf32; else unset ASLP_GEMM_SPLIT; fi
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'gemm_split': '$v', 'ms_per_step': round(d['ms_per_step'],3), 'lstm_ms': d['roofline']['avg_launch_ms']}))"
  done
done
