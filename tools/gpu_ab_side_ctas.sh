#!/bin/bash
# Same-box A/B: grid cap of the side-stream (weight-gradient) GEMMs that overlap the backward recurrence of the layer below.
# usage (under gpurun): bash tools/gpu_ab_side_ctas.sh > gpurun_out/ab_side_ctas.jsonl
for i in 1 2; do
  for v in 148 68 60 40; do
    ASLP_SIDE_GEMM_CTAS=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'side_gemm_ctas': $v, 'ms_per_step': round(d['ms_per_step'],3), 'lstm_ms': d['roofline']['avg_launch_ms']}))"
  done
done
