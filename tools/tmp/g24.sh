cd "${GRAFT_REPO_ROOT:-/root/repo}"
for w in 2 1; do echo "== split waves $w"; ASLP_GEMM_SPLIT_WAVES=$w python tools/config_bench.py 2>&1 | grep '^{' | cut -c1-200; done
timeout 600 python -m pytest tests/test_gpu_gemm.py -q -m gpu -k "bias_grad or gemm_ex" 2>&1 | tail -2
