cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graph.py -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/graph_tests.log
python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench3.log
ASLP_STEP_GRAPH=0 python tools/config_bench.py 2>&1 | grep '^{' | tee gpurun_out/config_bench3_eager.log
timeout 1200 python -m pytest tests/test_gpu_cli.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize_configs.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/graph_tests2.log
