#!/bin/bash
# One GPU-box pass: GPU test suite, smoke, bench (both arms), ncu launch list and a full-set capture of the big kernels.
# Usage (GPU box, from the repo root): bash tools/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
set -u
cd "$(dirname "$0")/.."
TAG=${1:-x}
DO_TESTS=${2:-tests}
DO_NCU=${3:-ncu}
O=gpurun_out
mkdir -p $O
if [ "$DO_TESTS" = tests ]; then
  python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1
  echo "rc=$?" >> $O/${TAG}_tests.log
  python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1
  echo "rc=$?" >> $O/${TAG}_smoke.log
fi
python bench.py --steps 50 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python tools/diag.py 64 512 > $O/${TAG}_diag.log 2>&1
if [ "$DO_NCU" = ncu ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph"
  $CMD > $O/${TAG}_ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv $CMD > $O/${TAG}_ncu_a.log 2>&1
  # full-set captures: the four persistent kernels of one step with source correlation, then the GEMM / sampler kernels of a
  # step without it (the report must stay below the 64 MiB that travel back)
  $CMD > $O/${TAG}_ncu_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on \
      -k regex:'dec_bwd_push_kernel|dec_fwd_ahead_kernel|enc_fwd_kernel|enc_bwd_ts_kernel' \
      -s 8 -c 4 -o $O/${TAG}_full -f $CMD > $O/${TAG}_ncu_b.log 2>&1
  $CMD > $O/${TAG}_ncu_plain3.log 2>&1 &&
  ncu --set full --clock-control none \
      -k regex:'umma_tn_kernel|umma_rm_kernel|sampler_fused' \
      -s 34 -c 17 -o $O/${TAG}_gemm -f $CMD > $O/${TAG}_ncu_c.log 2>&1
  # only the counters of the GEMM capture travel back (at most 64 MiB of gpurun_out/ are copied)
  ncu -i $O/${TAG}_gemm.ncu-rep --page raw --csv > $O/${TAG}_gemm_raw.csv 2>/dev/null && rm -f $O/${TAG}_gemm.ncu-rep
fi
cat $O/${TAG}_tests.log $O/${TAG}_smoke.log 2>/dev/null | tail -n 6
du -sh $O; ls -la $O | tail -n 20
