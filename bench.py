# coding: utf-8
"""Headline benchmark: ABCD-VAE training segments/s (forward + backward + clip + SGD step) on the publication
configuration (BASELINE.json configs[2]: batch 512, K=128, f=256, LSTM 256/256, F=65 log-STFT bins), synthetic
Bengalese-finch-like segment lengths.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this implementation
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1  # CPU restatement of the reference path (host cores)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` = segments/s with the batch resident in HBM; `e2e` = the same step
driven from pinned HOST buffers (H2D of the packed batch + D2H of the loss inside the timed region).
Weak scaling: every rank trains on its own 512-segment batch (global batch 512*N, gradients all-reduced with NCCL
and divided by the global batch size, see Learner.train_step).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

METRIC = 'train segments/sec (fwd+bwd+step)'
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the persistent kernels on this workload, from the
# `ncu --set full` capture summarised in profiles/r01_f_operand_images_B512.txt (B=512, F=65, LSTM)
NCU_TRAFFIC_BYTES = {
	'enc_fwd_kernel': 124.88e6 + 140.55e6, 'enc_bwd_kernel': 155.34e6 + 96.41e6,
	'dec_fwd_kernel': 11.96e6 + 80.89e6, 'dec_bwd_kernel': 144.30e6 + 75.26e6,
}
# sm__pipe_tensor_cycles_active / sm__inst_executed_pipe_fma (% of peak sustained active) from the same capture
NCU_PIPES = {
	'enc_fwd_kernel': {'tensor_pct': 11.3, 'fma_pct': 5.8}, 'enc_bwd_kernel': {'tensor_pct': 9.8, 'fma_pct': 4.3},
	'dec_fwd_kernel': {'tensor_pct': 0.0, 'fma_pct': 23.1}, 'dec_bwd_kernel': {'tensor_pct': 4.5, 'fma_pct': 4.7},
}
CFG = dict(rnn='LSTM', F=65, K=128, f=256, H=256, M=256, B=512, N_data=100000)


def parse():
	p = argparse.ArgumentParser()
	p.add_argument('--gpus', type=int, default=1)
	p.add_argument('--steps', type=int, default=100)
	p.add_argument('--warmup', type=int, default=10)
	p.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
	p.add_argument('--batch', type=int, default=CFG['B'], help='segments per GPU')
	p.add_argument('--bins', type=int, default=CFG['F'])
	p.add_argument('--rnn', type=str, default=CFG['rnn'])
	p.add_argument('--pretrain', action='store_true', help='softmax (pretraining) mode instead of Gumbel-softmax')
	p.add_argument('--no-cpu-baseline', action='store_true')
	p.add_argument('--cpu-steps', type=int, default=3)
	return p.parse_args()


# ---------------------------------------------------------------------------------------------------------
# synthetic workload (deterministic; identical for both arms)
# ---------------------------------------------------------------------------------------------------------
def make_batch(B, F, seed, rank=0, world=1):
	"""Synthetic batch of B segments for one rank.  With world > 1 the GLOBAL length-sorted batch of B*world segments is
	sharded the way training shards it (dist.shard_sorted_batch: rank r takes positions r, r+world, ...), so every rank
	sees the same length distribution; values are drawn per rank."""
	from oracle import cases as CS
	from oracle import abcd_oracle as O
	if world == 1:
		lengths = CS.synthetic_lengths(B, 1000 + seed)
	else:
		dist = importlib.import_module('seq2seq_abcd-vae_b200.dist')
		_, lengths = dist.shard_sorted_batch(CS.synthetic_lengths(B * world, 1000), rank, world)
		lengths = np.ascontiguousarray(lengths)
		seed = seed + rank
	rng = np.random.RandomState(seed)
	sumT = int(lengths.sum())
	x = np.clip(rng.normal(0.45, 0.25, size=(sumT, F)), -0.945, 1.05).astype(np.float32)
	bs = O.batch_sizes_from_lengths(lengths.tolist())
	offsets = np.concatenate([[0], np.cumsum(bs)])
	is_off = np.zeros(sumT, dtype=np.float32)
	for b, l in enumerate(lengths):
		is_off[offsets[l - 1] + b] = 1.0
	return lengths, torch.tensor(bs, dtype=torch.int64), torch.from_numpy(x), torch.from_numpy(is_off)


def flops_per_step(lengths, F, H=256, M=256, K=128, f=256, rnn='LSTM'):
	"""Algorithmic training FLOPs (3 x forward)."""
	sumT, B = int(lengths.sum()), len(lengths)
	g = 4 if rnn == 'LSTM' else 3
	per_frame = 2 * (2 * g * H * (F + H)) + 2 * (g * H * (F + H)) + 2 * (2 * H * M + 2 * M * F) + 2 * (H * M + M)
	enc_out = (2 if rnn == 'LSTM' else 1) * 2 * H
	per_seg = 2 * (enc_out * M + M * f) + 2 * f * K + 2 * K * f + 2 * f * ((2 if rnn == 'LSTM' else 1) * H)
	return 3 * (sumT * per_frame + B * per_seg)


class ClockSampler(object):
	"""nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""

	def __init__(self, gpu_index):
		self.rows = []
		self.proc = None
		self.gpu = gpu_index

	def start(self):
		q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
		try:
			self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '50'],
										stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
			self.thread = threading.Thread(target=self._read, daemon=True)
			self.thread.start()
		except OSError:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			self.rows.append([c.strip() for c in line.split(',')])

	def stop(self):
		if self.proc is None:
			return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except subprocess.TimeoutExpired:
			self.proc.kill()
		sm, mx, reasons = [], [], set()
		names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
		for r in self.rows:
			try:
				sm.append(float(r[0])); mx.append(float(r[1]))
			except (ValueError, IndexError):
				continue
			for n, v in zip(names, r[3:7]):
				if v.lower().startswith('active'):
					reasons.add(n)
		return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
				'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's batched CPU restatement of the same training step
# ---------------------------------------------------------------------------------------------------------
def cpu_step_factory(args, lengths, bs, x, is_off):
	from oracle import abcd_oracle as O
	from oracle import cases as CS
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	cfg = dict(rnn=args.rnn, F=args.bins, K=CFG['K'], f=CFG['f'])
	mods = CS.build_modules(pkg.model, cfg)
	sds = [CS.det_state(m, 500 + i) for i, m in enumerate(mods)]
	seqs = O.unpack(x, bs.tolist())
	sumT = x.shape[0]
	eps = torch.from_numpy(np.random.RandomState(7).standard_normal((sumT, args.bins)).astype(np.float32))
	gum = torch.from_numpy(np.random.RandomState(8).gumbel(size=(len(lengths), CFG['K'])).astype(np.float32))
	state = {'sds': sds}

	def step():
		enc_sd, smp_sd, dec_sd = state['sds']
		r = O.train_step(enc_sd, smp_sd, dec_sd, seqs, eps, gum, CFG['N_data'], args.pretrain, 1.0, args.rnn, 1.0, None, batched=True)
		names, params, grads = [], [], []
		for m, sd in zip(('encoder', 'feature_sampler', 'decoder'), (enc_sd, smp_sd, dec_sd)):
			for k, g in r['grads'][m].items():
				names.append((m, k)); params.append(sd[k]); grads.append(g)
		new, _, _ = O.clip_and_sgd(params, grads, 1.0, 1.0)
		for (m, k), p in zip(names, new):
			{'encoder': enc_sd, 'feature_sampler': smp_sd, 'decoder': dec_sd}[m][k] = p.detach()
		return float(r['loss'])
	return step


def run_cpu(args, steps, warmup):
	lengths, bs, x, is_off = make_batch(args.batch, args.bins, 0)
	torch.set_num_threads(os.cpu_count() or 1)
	step = cpu_step_factory(args, lengths, bs, x, is_off)
	for _ in range(warmup):
		step()
	t0 = time.perf_counter()
	for _ in range(steps):
		step()
	dt = (time.perf_counter() - t0) / max(steps, 1)
	return args.batch / dt, dt, lengths


def reference_arm(args):
	rank = int(os.environ.get('RANK', '0'))
	if rank != 0:
		return
	steps, warmup = max(1, min(args.steps, 3)), max(0, min(args.warmup, 1))
	val, dt, lengths = run_cpu(args, steps, warmup)
	cores = torch.get_num_threads()
	sample = '{} warm-up + {} timed full steps of the B={} batch (sumT={})'.format(warmup, steps, args.batch, int(lengths.sum()))
	line = {
		'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'segments/s', 'n_gpus': args.gpus, 'steps': steps,
		'warmup': warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
		'dtype': 'f32', 'data': 'synthetic',
		'config': workload_config(args, lengths, 1),
		'cpu_baseline': {'value': val, 'unit': 'segments/s', 'cores': cores, 'kind': 'port', 'sample': sample},
		'e2e': {'value': val, 'unit': 'segments/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
		'gpu_launches': 0,
	}
	emit_json(line)


def workload_config(args, lengths, world):
	return {
		'workload': 'ABCD-VAE publication config (BASELINE.json configs[2]): {} segments/GPU, K=128, f=256, {} 256/256, mlp 256, F={} bins, '
					'T in [{}, {}] mean {:.1f} frames, sumT={}/GPU, {} mode, SGD lr 1.0 clip 1.0'.format(
						args.batch, args.rnn, args.bins, int(lengths.min()), int(lengths.max()), float(lengths.mean()), int(lengths.sum()),
						'pretrain-softmax' if args.pretrain else 'Gumbel-softmax'),
		'global_batch': args.batch * world, 'parallelism': 'dp{}'.format(world),
		'l2_policy': 'per-step working set (saved activations + gate gradients, ~0.6 GB at B=512) exceeds the 126 MB L2; no explicit flush',
	}


# ---------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------
_JSON_FD = None


def emit_json(line):
	"""The one JSON line of the contract, on the process's ORIGINAL stdout."""
	sys.stdout.flush()
	os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + '\n').encode())


def main():
	global _JSON_FD
	# stdout must carry exactly one JSON line, but native libraries print there too (NCCL's version banner when the
	# environment sets NCCL_DEBUG): keep the original stdout for the JSON line and point fd 1 at stderr for everything else
	sys.stdout.flush()
	_JSON_FD = os.dup(1)
	os.dup2(2, 1)
	args = parse()
	if args.impl == 'reference':
		reference_arm(args)
		return
	rank = int(os.environ.get('RANK', '0'))
	local_rank = int(os.environ.get('LOCAL_RANK', '0'))
	world = int(os.environ.get('WORLD_SIZE', '1'))
	assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)'
	torch.cuda.set_device(local_rank)
	dev = torch.device('cuda', local_rank)
	if world > 1:
		torch.distributed.init_process_group('nccl', device_id=dev)
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	L = pkg._lib
	from oracle import cases as CS

	lengths, bs, x_host, off_host = make_batch(args.batch, args.bins, 0, rank, world)
	learner = learning.Learner(args.bins, 256, 256, 256, CFG['K'], CFG['f'], None, encoder_rnn_type=args.rnn,
								decoder_rnn_type=args.rnn, device=str(dev), seed=1111)
	# identical deterministic weights on every rank (and in the CPU arm)
	for i, m in enumerate(learner._modules()):
		m.load_state_dict({k: v.to(dev) for k, v in CS.det_state(m, 500 + i).items()})
	learner.setup_optimizer(1.0, 0.0, gradient_clip=1.0, patience=0)
	learner.feature_sampler.train()
	x_pin, off_pin = x_host.pin_memory(), off_host.pin_memory()
	x_dev, off_dev = x_pin.to(dev), off_pin.to(dev)
	packed = torch.nn.utils.rnn.PackedSequence(x_dev, bs)
	packed_off = torch.nn.utils.rnn.PackedSequence(off_dev, bs)
	gb = args.batch * world

	def step_resident():
		return learner.train_step(packed, packed_off, None, CFG['N_data'], args.pretrain, global_batch_size=gb)[0]

	# end-to-end path: every step copies its packed batch from pinned host memory and reads the loss back.  Like a data
	# loader with one batch of look-ahead, the copy of step i+1 is issued on a copy stream before step i computes (two
	# device staging slots); the loss read-back synchronises every step.
	x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
	off_stage = [torch.empty_like(off_dev), torch.empty_like(off_dev)]
	loss_host = torch.empty((), dtype=torch.float32).pin_memory()
	copy_stream = torch.cuda.Stream(device=dev)
	copied = [torch.cuda.Event(), torch.cuda.Event()]
	e2e_state = {'i': 0, 'primed': False}

	def issue_copy(slot):
		with torch.cuda.stream(copy_stream):
			x_stage[slot].copy_(x_pin, non_blocking=True)
			off_stage[slot].copy_(off_pin, non_blocking=True)
			copied[slot].record(copy_stream)

	def step_e2e():
		slot = e2e_state['i'] & 1
		if not e2e_state['primed']:
			issue_copy(slot)
			e2e_state['primed'] = True
		issue_copy(slot ^ 1)   # next step's batch (its slot was last read by the previous step, which has completed)
		torch.cuda.current_stream().wait_event(copied[slot])
		p = torch.nn.utils.rnn.PackedSequence(x_stage[slot], bs)
		po = torch.nn.utils.rnn.PackedSequence(off_stage[slot], bs)
		loss = learner.train_step(p, po, None, CFG['N_data'], args.pretrain, global_batch_size=gb)[0]
		loss_host.copy_(loss.detach(), non_blocking=True)
		torch.cuda.current_stream().synchronize()
		e2e_state['i'] += 1
		return loss_host

	def barrier():
		if world > 1:
			torch.distributed.barrier()
		torch.cuda.synchronize()

	def timed(fn, steps):
		import gc
		gc.collect()
		gc.disable()   # no collector pauses inside the timed region (they only show in the per-step-synchronised e2e loop)
		barrier()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(steps):
			fn()
		e1.record()
		barrier()
		gc.enable()
		ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
		if world > 1:
			torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
		return float(ms.item())

	warm = max(args.warmup, 3)
	for _ in range(warm):
		step_resident()
	clocks = ClockSampler(local_rank)
	if rank == 0:
		clocks.start()
	launches0 = L.load().abcd_launch_count()
	total_ms = timed(step_resident, args.steps)
	launches = L.load().abcd_launch_count() - launches0
	clock_info = clocks.stop() if rank == 0 else None
	ms_per_step = total_ms / args.steps
	value = gb / (ms_per_step * 1e-3)

	for _ in range(2):
		step_e2e()
	e2e_ms = timed(step_e2e, args.steps) / args.steps
	e2e_value = gb / (e2e_ms * 1e-3)
	h2d = x_pin.numel() * 4 + off_pin.numel() * 4
	d2h = 4

	# kernel-group timing pass (CUDA events recorded by the library on the launching stream)
	L.profile_enable(True)
	prof_steps = min(args.steps, 10)
	barrier()
	for _ in range(prof_steps):
		step_resident()
	barrier()
	L.profile_enable(False)
	prof = L.profile_read_all()
	groups = {k: {'ms_per_step': v[0] / prof_steps, 'launches_per_step': v[1] / prof_steps} for k, v in prof.items() if v[1]}
	peaks = {}
	try:
		peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except (OSError, ValueError):
		pass
	sumT = int(lengths.sum())
	g = 4 if args.rnn == 'LSTM' else 3
	F, H, M = args.bins, 256, 256
	kernel_flops = {   # algorithmic FLOPs per launch of each persistent kernel (DESIGN.md "Kernels")
		'enc_fwd_kernel': sumT * 2 * (2 * g * H * H),
		'enc_bwd_kernel': sumT * 2 * (2 * g * H * H),
		'dec_fwd_kernel': sumT * 2 * (g * H * (F + H) + 2 * H * M + 2 * M * F),
		'dec_bwd_kernel': sumT * 2 * (g * H * (F + H) + 2 * H * M + 2 * M * F),
	}
	dom = max((k for k in groups if k in kernel_flops), key=lambda k: groups[k]['ms_per_step'], default=None)
	roofline = None
	if dom is not None:
		dur = groups[dom]['ms_per_step'] / max(groups[dom]['launches_per_step'], 1) * 1e-3
		achieved = kernel_flops[dom] / dur / 1e12
		peak = peaks.get('bf16_tflops_sustained', 1400.0)
		publication = args.batch == CFG['B'] and args.bins == CFG['F'] and args.rnn == 'LSTM'
		roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
					'traffic': NCU_TRAFFIC_BYTES.get(dom) if publication else None, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1400 (B200_PROFILING.md)',
					'traffic_source': 'profiles/r01_f_operand_images_B512.txt (ncu --set full, bytes per launch)' if publication else None,
					'ncu_pipes': NCU_PIPES.get(dom) if publication else None,
					'note': 'fp32-grade recurrent kernel (tcgen05 3xTF32 products with the weights in tensor memory + FFMA phases); algorithmic FLOPs over the measured bf16 tensor peak; the fp32 FFMA peak is ~72 TFLOP/s',
					'frac_of_fp32_ffma_peak': achieved / 72.0, 'avg_launch_ms': dur * 1e3}

	n_params = learner.arena.numel
	hbm_peak = peaks.get('hbm_gbs', 6650.0)
	other = {}
	if 'clip_sgd' in groups:
		gbs = n_params * 16 / (groups['clip_sgd']['ms_per_step'] * 1e-3) / 1e9   # 16 B/param: 4 norm read + 8 read + 4 write
		other['clip_sgd'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
							'note': '27 MB arena, L2-resident between the norm and step passes'}
	if 'philox' in groups:
		gbs = (sumT * F + args.batch * CFG['K']) * 4 / (groups['philox']['ms_per_step'] * 1e-3) / 1e9
		other['philox'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak}
	# front end of the same batch: log-amplitude STFT (8 ms frame / 4 ms hop at 16 kHz -> n_fft 128, hop 64) of synthetic
	# int16 audio straight into packed order; algorithmic bytes = int16 samples read once + fp32 bins written
	try:
		du = importlib.import_module('seq2seq_abcd-vae_b200.modules.data_utils')
		n_fft, hop = 2 * (F - 1), F - 1
		nsamp = torch.tensor([(int(l) - 1) * hop + hop // 2 for l in lengths], dtype=torch.int32)
		onset = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), nsamp[:-1].long()]), 0)
		wav = torch.from_numpy(np.random.RandomState(5).randint(-3000, 3000, size=int(nsamp.sum())).astype(np.int16)).to(dev)
		desc = pkg.model.as_packed_desc(bs, dev)
		win = torch.hann_window(n_fft).to(dev)
		out = torch.empty(sumT, F, device=dev)
		onset_d, nsamp_d = onset.to(dev), nsamp.to(dev)
		for _ in range(3):
			du._stft_launch(wav, onset_d, nsamp_d, desc, n_fft, hop, True, win, 1e-15, 11.0, out)
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		torch.cuda.synchronize(); e0.record()
		for _ in range(20):
			du._stft_launch(wav, onset_d, nsamp_d, desc, n_fft, hop, True, win, 1e-15, 11.0, out)
		e1.record(); torch.cuda.synchronize()
		stft_ms = e0.elapsed_time(e1) / 20
		gbs = (wav.numel() * 2 + out.numel() * 4) / (stft_ms * 1e-3) / 1e9
		other['stft_logamp'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
								'ms': stft_ms, 'frames': sumT, 'note': 'n_fft {} hop {}; 7.7 MB working set is L2-resident between repeats'.format(n_fft, hop)}
	except Exception as e:   # the front end is not on the timed path: report, do not fail the bench
		other['stft_logamp'] = {'error': repr(e)}
	if 'rowwise' in groups:
		# sampler row kernels (softmax / Gumbel-softmax, KL rows, statistics): logits [B,K] read + y written per kernel
		other['sampler_rowwise'] = {'ms_per_step': groups['rowwise']['ms_per_step'], 'launches_per_step': groups['rowwise']['launches_per_step'],
									'note': 'latency-bound: 512 x 128 fp32 rows (256 KB) per kernel, one warp per segment'}
	cpu_baseline = None
	if rank == 0 and not args.no_cpu_baseline:
		cval, cdt, _ = run_cpu(args, args.cpu_steps, 1)
		cpu_baseline = {'value': cval, 'unit': 'segments/s', 'cores': torch.get_num_threads(), 'kind': 'port',
						'sample': '1 warm-up + {} timed full steps of the same B={} batch on the host cores ({:.2f} s/step)'.format(args.cpu_steps, args.batch, cdt)}
	if rank == 0:
		line = {
			'metric': METRIC, 'value': value, 'unit': 'segments/s', 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
			'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
			'data': 'synthetic', 'config': workload_config(args, lengths, world), 'clocks': clock_info,
			'e2e': {'value': e2e_value, 'unit': 'segments/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_ms},
			'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
			'kernel_groups': groups, 'rooflines_other': other,
			'train_flops_per_step': flops_per_step(lengths, args.bins, rnn=args.rnn),
			'achieved_tflops_whole_step': flops_per_step(lengths, args.bins, rnn=args.rnn) / (ms_per_step * 1e-3) / 1e12,
		}
		emit_json(line)
	if world > 1:
		torch.distributed.destroy_process_group()


if __name__ == '__main__':
	main()
