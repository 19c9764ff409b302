#!/bin/bash
# compute-sanitizer passes over the persistent cluster kernels: one small LSTM and one small GRU+speaker
# training step (forward + backward through every kernel of the library) under memcheck, synccheck and racecheck.
# Usage (GPU box): bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1
set -u
cd "$(dirname "$0")/.."
export ABCD_SANITIZE=1
for tool in memcheck synccheck racecheck; do
  echo "===== compute-sanitizer --tool $tool ====="
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train_lstm_gumbel or train_gru_speaker or train_plain_lstm" 2>&1 | tail -n 25
  echo "exit code of $tool: ${PIPESTATUS[0]}"
done
