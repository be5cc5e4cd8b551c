#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/ngpus.txt
for n in 8 4; do
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  echo "n=$n exit=$?"; tail -1 gpurun_out/bench_n$n.log | cut -c1-420
done
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1
echo "ref n=8 exit=$?"; tail -1 gpurun_out/bench_ref_n8.log | cut -c1-300
