#!/bin/bash
# compute-sanitizer evidence: memcheck (global / shared out-of-bounds, misaligned accesses) and racecheck (shared-memory
# hazards) over the sentinel-polled persistent recurrence kernels, the one-launch CTC kernel and the GEMM epilogues.
# Writes gpurun_out/sanitizer.txt (copied to profiles/rNN_sanitizer.txt).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer.txt
LSTM='test_lstm_fwd_bwd_parity and auto and (30-16-320-0-2 or 9-5-24-0-2 or 5-3-8-4-1 or 6-16-512-0-1 or 20-16-64-32-2)'
BWD='test_backward_recurrence_forms and transposed and (40-12-320-2 or 6-16-512-1 or 7-8-384-1)'
CTC='fused and (small_known or inf_activation or ragged or vs_oracle)'
GEMM='test_gemm_ex_epilogue_matches_the_separate_steps and (256-1024-1024-0-1 or 24-36-20) or test_bias_grad_update'
echo "# compute-sanitizer $(compute-sanitizer --version | head -2 | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)" > $OUT
run() {
  local tool=$1 name=$2 file=$3 sel=$4
  echo "## --tool $tool: pytest $file -k \"$sel\"" >> $OUT
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 python -m pytest $file -q -m gpu -p no:cacheprovider -k "$sel" > gpurun_out/san_$name.log 2>&1
  echo "exit code $? (99 = the sanitizer reported errors, 137 = timed out)" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error:|Race reported|Invalid|hazard" gpurun_out/san_$name.log | sort | uniq -c | head -20 >> $OUT
}
if [ "$1" == "late" ]; then         # kernels added after the "all" pass of the round: conv / pool row kernels, cluster split-K reduce
  OUT=gpurun_out/sanitizer_late.txt
  echo "# compute-sanitizer $(compute-sanitizer --version | head -2 | tail -1) on $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1)" > $OUT
  run memcheck mem_gemm_late tests/test_gpu_gemm.py "not full_size"
  run memcheck mem_pointwise_late tests/test_gpu_pointwise.py "conv or maxpool"
  run racecheck race_pointwise_late tests/test_gpu_pointwise.py "conv or maxpool"
  run racecheck race_gemm_ex_late tests/test_gpu_gemm.py "test_gemm_ex_epilogue_matches_the_separate_steps and (256-1024-1024-0-1 or 1024-440-256 or 24-36-20)"
  run memcheck mem_workers_late tests/test_gpu_workers.py "single_rank"
elif [ "$1" == "all" ]; then          # whole test files (about ten minutes)
  ALL='not full_size and not many_utts'
  run memcheck mem_lstm_all tests/test_gpu_lstm.py "$ALL"
  run memcheck mem_ctc_all tests/test_gpu_ctc.py "$ALL"
  run memcheck mem_eesen_all tests/test_gpu_ctc_eesen.py "$ALL"
  run memcheck mem_gemm_all tests/test_gpu_gemm.py "$ALL"
  run memcheck mem_pointwise_all tests/test_gpu_pointwise.py "$ALL"
  run racecheck race_lstm_all tests/test_gpu_lstm.py "$ALL"
  run racecheck race_ctc_all tests/test_gpu_ctc.py "$ALL"
  run racecheck race_eesen_all tests/test_gpu_ctc_eesen.py "$ALL"
  run racecheck race_pointwise_all tests/test_gpu_pointwise.py "$ALL"
else
  run memcheck mem_lstm tests/test_gpu_lstm.py "$LSTM"
  run memcheck mem_bwd tests/test_gpu_lstm.py "$BWD"
  run memcheck mem_ctc tests/test_gpu_ctc.py "$CTC"
  run memcheck mem_gemm tests/test_gpu_gemm.py "$GEMM"
  run racecheck race_lstm tests/test_gpu_lstm.py "$LSTM"
  run racecheck race_bwd tests/test_gpu_lstm.py "$BWD"
  run racecheck race_ctc tests/test_gpu_ctc.py "$CTC"
fi
cat $OUT
