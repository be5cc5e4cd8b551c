#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
run() { local name=$1; shift; timeout -s KILL 900 "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit=$?" >> gpurun_out/summary.txt; tail -n 12 gpurun_out/t_$name.log | cut -c1-1500 | sed "s/^/[$name] /" >> gpurun_out/summary.txt; }
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
run workers python -m pytest tests/test_gpu_workers.py -q -m gpu -p no:cacheprovider
run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3
run bench1 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline
cat gpurun_out/summary.txt
