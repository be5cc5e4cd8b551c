cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== new code"; timeout 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu -k "tiny_and_widely" 2>&1 | tail -3
echo "== shuffles through fp16 (HEAD)"; ASLP_B200_CUDA_LIB=$PWD/kaldi-aslp_b200/libaslp_b200_oldshfl.so timeout 600 python -m pytest tests/test_gpu_lstm.py -q -m gpu -k "tiny_and_widely" 2>&1 | tail -4
bash tools/gpu_ab_fwd.sh nohyb
timeout -s KILL 300 python tools/perf_probe.py timing 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['op'], round(d['us_per_step'],3), d['cycles_per_step_mean'])"
timeout 900 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_nnet_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
