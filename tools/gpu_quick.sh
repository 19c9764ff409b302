#!/bin/bash
# Quick GPU check of a kernel change: parity suites + per-kernel timing.   bash tools/gpu_quick.sh <tag> [timing env value]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-q}
TIM=${2:-1}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/${TAG}_tests.log 2>&1; echo rc=$? >> $O/${TAG}_tests.log
timeout 300 python tools/diag.py 64 256 512 > $O/${TAG}_diag.log 2>&1
ABCD_TIMING=$TIM timeout 300 python tools/diag.py 64 512 > $O/${TAG}_diag_t.log 2>&1
tail -n 3 $O/${TAG}_tests.log; cat $O/${TAG}_diag.log
