cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1200 python -m pytest tests/test_gpu_workers.py tests/test_gpu_async_servers.py tests/test_gpu_cli.py -q -m gpu -p no:cacheprovider -rs 2>&1 | tail -15 | tee gpurun_out/t_two_gpu_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2>&1
echo "n=2 exit=$?"; grep '^{' gpurun_out/bench_n2.log | tail -1 | cut -c1-300
