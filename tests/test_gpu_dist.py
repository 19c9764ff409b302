# coding: utf-8
"""GPU suite, part 5: data-parallel correctness on REAL GPUs (needs >= 2 visible devices, otherwise skipped; the same
logic runs on 2 gloo ranks in tests/test_cpu_dist.py).  Launches tools/check_dist_gpu.py under torch.distributed.run:
sharded steps vs the full batch in one process, ranks bit-identical, CUDA-graph replay (NCCL inside the graph) vs
host-launched steps."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_nccl_steps_match_single_process():
	if torch.cuda.device_count() < 2:
		pytest.skip('needs 2 GPUs')
	cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
			'--master-port', '29544', os.path.join(ROOT, 'tools', 'check_dist_gpu.py')]
	r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
	print(r.stdout[-3000:], r.stderr[-3000:])
	assert r.returncode == 0
