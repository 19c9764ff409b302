#!/bin/bash
# Other shapes of the training step and the bulk-encode benchmark on one B200 (numbers quoted in DESIGN.md section 5).
# Usage (GPU box): bash tools/gpu_variants.sh <tag>
set -u
cd "$(dirname "$0")/.."
T=${1:-v}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline"
$B --rnn GRU --speakers 16 > $O/${T}_gru_speaker.json 2> $O/${T}_gru_speaker.err
$B --latent gaussian > $O/${T}_plain.json 2> $O/${T}_plain.err
$B --pretrain > $O/${T}_pretrain.json 2> $O/${T}_pretrain.err
$B --bins 129 > $O/${T}_f129.json 2> $O/${T}_f129.err
$B --batch 2048 > $O/${T}_b2048.json 2> $O/${T}_b2048.err
$B --batch 128 > $O/${T}_b128.json 2> $O/${T}_b128.err
$B --batch 64 > $O/${T}_b64.json 2> $O/${T}_b64.err
python tools/bench_encode.py > $O/${T}_encode.log 2>&1
for f in gru_speaker plain pretrain f129 b2048 b128 b64; do python - <<PY
import json
try:
    d = json.loads(open('$O/${T}_$f.json').read().strip().splitlines()[-1])
    print('$f', round(d['value']), 'seg/s', round(d['ms_per_step'], 3), 'ms/step; e2e', round(d['e2e']['value']))
except Exception as e:
    print('$f', 'failed', e)
PY
done
tail -n 5 $O/${T}_encode.log
