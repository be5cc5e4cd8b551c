cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for n in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  echo "n=$n exit=$?"; grep '^{' gpurun_out/bench_n$n.log | tail -1 | cut -c1-180
done
