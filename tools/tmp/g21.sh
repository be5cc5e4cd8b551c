cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dropin_mains.py tests/test_gpu_cli.py -q -m gpu -p no:cacheprovider 2>&1 | tail -6
