#!/bin/bash
# Quick GPU check of a kernel change: parity suites, per-phase cycle counters of the decoder kernels (two CTAs of cluster 0),
# per-kernel timing at three batch sizes.   bash tools/gpu_quick.sh <tag>
set -u
cd "$(dirname "$0")/.."
T=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $O/${T}_tests.log 2>&1; echo rc=$? >> $O/${T}_tests.log
for c in 0 13; do
  ABCD_TIMING=1 ABCD_TIMING_CTA=$c timeout 300 python tools/diag.py 512 2>&1 | grep -A1 "dec_fwd CTA0" > $O/${T}_fwd_cta$c.log
  ABCD_TIMING=3 ABCD_TIMING_CTA=$c timeout 300 python tools/diag.py 512 2>&1 | grep "dec_bwd CTA0" > $O/${T}_bwd_cta$c.log
done
timeout 300 python tools/diag.py 64 256 512 > $O/${T}_diag.log 2>&1
tail -n 3 $O/${T}_tests.log; cat $O/${T}_fwd_cta*.log $O/${T}_bwd_cta*.log $O/${T}_diag.log
