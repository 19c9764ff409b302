# coding: utf-8
"""Headline benchmark: ABCD-VAE training segments/s (forward + backward + clip + SGD step) on the publication
configuration (BASELINE.json configs[2]: batch 512, K=128, f=256, LSTM 256/256, F=65 log-STFT bins), synthetic
Bengalese-finch-like segment lengths.

    python bench.py --gpus 1 --steps 20 --warmup 5                 # this implementation
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1  # CPU restatement of the reference path (host cores)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --rnn GRU --speakers 16        # BASELINE configs[3];   python bench.py --latent gaussian   # configs[1] (plain/)

One JSON line on stdout (rank 0).  `value` = segments/s with the batch resident in HBM; `e2e` = the same step
driven from pinned HOST buffers (H2D of the packed batch + D2H of the loss inside the timed region).
`scaling` is "weak": every rank trains on its own 512-segment shard of a 512*N global batch (gradients all-reduced with
NCCL, divided by the global batch size, see Learner.train_step).  With N > 1 the line additionally carries `strong`: the
publication batch itself (512 segments GLOBAL) sharded over the N ranks -- the north star's "batch 512 on 8xB200".
"""
import argparse
import gc
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
CFG = dict(rnn='LSTM', F=65, K=128, f=256, H=256, M=256, B=512, N_data=100000)
PERSISTENT = ('enc_fwd_kernel', 'enc_bwd_kernel', 'dec_fwd_kernel', 'dec_bwd_kernel')
# exchanges through L2 (publish -> cluster barrier -> consume) per time step and cluster size of each persistent kernel (DESIGN.md
# section 4).  The decoder kernels' third exchange (x_t forward, the dML rows backward) is a st.async push into the consumers'
# shared memory that completes on an mbarrier: no fence, no barrier, no L2 round trip -- it is not counted here.
EXCHANGES = {'enc_fwd_kernel': (1, 8), 'enc_bwd_kernel': (1, 8), 'dec_fwd_kernel': (2, 16), 'dec_bwd_kernel': (2, 16)}


def parse():
	p = argparse.ArgumentParser()
	p.add_argument('--gpus', type=int, default=1)
	p.add_argument('--steps', type=int, default=100)
	p.add_argument('--warmup', type=int, default=10)
	p.add_argument('--impl', type=str, default='b200', choices=['b200', 'reference'])
	p.add_argument('--batch', type=int, default=CFG['B'], help='segments per GPU')
	p.add_argument('--bins', type=int, default=CFG['F'])
	p.add_argument('--rnn', type=str, default=CFG['rnn'])
	p.add_argument('--speakers', type=int, default=0, help='> 0: --speaker_embed_dim 256 with this many speakers (BASELINE configs[3])')
	p.add_argument('--latent', type=str, default='abcd', choices=['abcd', 'gaussian'], help='gaussian = the plain/ VAE (BASELINE configs[1])')
	p.add_argument('--pretrain', action='store_true', help='softmax (pretraining) mode instead of Gumbel-softmax')
	p.add_argument('--no-graph', action='store_true', help='launch every kernel from the host instead of replaying the captured CUDA graph')
	p.add_argument('--no-cpu-baseline', action='store_true')
	p.add_argument('--no-strong', action='store_true')
	p.add_argument('--cpu-steps', type=int, default=3)
	return p.parse_args()


# ---------------------------------------------------------------------------------------------------------
# synthetic workload (deterministic; identical for both arms)
# ---------------------------------------------------------------------------------------------------------
def make_batch(B, F, seed, rank=0, world=1):
	"""Synthetic batch for one rank: the GLOBAL length-sorted batch of B*world segments sharded the way training shards it
	(rank r takes positions r, r+world, ...), so every rank sees the same length distribution; values are drawn per rank."""
	S = importlib.import_module('seq2seq_abcd-vae_b200.synthetic')
	if world == 1:
		lengths = S.synthetic_lengths(B, 1000 + seed)
	else:
		lengths = np.ascontiguousarray(S.synthetic_lengths(B * world, 1000)[rank::world])
		seed = seed + rank
	bs = S.batch_sizes_of(lengths)
	x = S.synthetic_frames(int(lengths.sum()), F, seed)
	return lengths, bs, x, S.is_offset_of(lengths, bs)


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
def case_config(args):
	cfg = dict(rnn=args.rnn, F=args.bins, K=CFG['K'], f=16 if args.latent == 'gaussian' else CFG['f'])
	if args.latent == 'gaussian':
		cfg['latent'] = 'gaussian'
	if args.speakers:
		cfg.update(num_speakers=args.speakers, E=256)
	return cfg


def cpu_step_factory(args, lengths, bs, x, is_off):
	from oracle import abcd_oracle as O
	from oracle import cases as CS
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	cfg = case_config(args)
	mods = CS.build_modules(pkg.model, cfg)
	sds = [CS.det_state(m, 500 + i) for i, m in enumerate(mods)]
	seqs = O.unpack(x, bs.tolist())
	sumT = x.shape[0]
	eps = torch.from_numpy(np.random.RandomState(7).standard_normal((sumT, args.bins)).astype(np.float32))
	if args.latent == 'gaussian':
		gum = torch.from_numpy(np.random.RandomState(8).standard_normal((len(lengths), cfg['f'])).astype(np.float32))
	else:
		gum = torch.from_numpy(np.random.RandomState(8).gumbel(size=(len(lengths), CFG['K'])).astype(np.float32))
	speaker = torch.arange(len(lengths)) % args.speakers if args.speakers else None
	state = {'sds': sds}

	def step():
		enc_sd, smp_sd, dec_sd = state['sds']
		r = O.train_step(enc_sd, smp_sd, dec_sd, seqs, eps, gum, CFG['N_data'], args.pretrain, 1.0, args.rnn, 1.0, speaker, batched=True,
						latent=args.latent)
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
	variant = 'plain/ Gaussian VAE (BASELINE.json configs[1]), f=16' if args.latent == 'gaussian' else \
		('ABCD-VAE -R GRU --speaker_embed_dim 256 (BASELINE.json configs[3]), K=128, f=256' if args.speakers else
		'ABCD-VAE publication config (BASELINE.json configs[2]), K=128, f=256')
	return {
		'workload': '{}: {} segments/GPU, {} 256/256, mlp 256, F={} bins, T in [{}, {}] mean {:.1f} frames, sumT={}/GPU, {} mode, '
					'SGD lr 1.0 clip 1.0'.format(variant, args.batch, args.rnn, args.bins, int(lengths.min()), int(lengths.max()), float(lengths.mean()),
												int(lengths.sum()), 'pretrain-softmax' if args.pretrain else ('reparameterised' if args.latent == 'gaussian' else 'Gumbel-softmax')),
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


def load_json(name):
	try:
		return json.load(open(os.path.join(ROOT, name)))
	except (OSError, ValueError):
		return {}


def measure_tf32_peak(dev):
	"""Dense TF32 tensor-core throughput measured the way MEASURED_PEAKS.json measures bf16: torch.matmul 8192^3 with TF32
	enabled (cuBLAS), best of 10 (burst) and back to back for ~1 s (sustained)."""
	prev = torch.backends.cuda.matmul.allow_tf32
	torch.backends.cuda.matmul.allow_tf32 = True
	try:
		n = 8192
		a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
		for _ in range(2):
			torch.matmul(a, b)
		best = None
		for _ in range(10):
			e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
			e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
			ms = e0.elapsed_time(e1)
			best = ms if best is None else min(best, ms)
		reps = max(5, int(1000.0 / best))
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(reps):
			torch.matmul(a, b)
		e1.record(); torch.cuda.synchronize()
		sustained = e0.elapsed_time(e1) / reps
		fl = 2.0 * n ** 3
		return {'tf32_tflops': fl / (best * 1e-3) / 1e12, 'tf32_tflops_sustained': fl / (sustained * 1e-3) / 1e12,
				'how': 'torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS): best of 10 and {} back-to-back launches'.format(reps)}
	finally:
		torch.backends.cuda.matmul.allow_tf32 = prev


def measure_exchange(L, dev):
	"""ns per publish -> cluster barrier -> consume round on one cluster (abcd_measure_cluster_exchange)."""
	out = torch.zeros(2, device=dev)
	scratch = torch.zeros(16 * 32, device=dev)
	res = {}
	for cl in (8, 16):
		for mode, name in ((0, 'barrier'), (1, 'token'), (2, 'fanin')):
			L.call('abcd_measure_cluster_exchange', cl, 20, mode, L.ptr(out), L.ptr(scratch), L.stream_ptr())   # warm-up
			L.call('abcd_measure_cluster_exchange', cl, 2000, mode, L.ptr(out), L.ptr(scratch), L.stream_ptr())
			torch.cuda.synchronize()
			res['cl{}_{}_ns'.format(cl, name)] = float(out[0].item())
	return res


def build_learner(args, dev, process_group=None):
	S = importlib.import_module('seq2seq_abcd-vae_b200.synthetic')
	kw = dict(encoder_rnn_type=args.rnn, decoder_rnn_type=args.rnn, device=str(dev), seed=1111,
			num_speakers=args.speakers or None, speaker_embed_dim=256 if args.speakers else None)
	if args.latent == 'gaussian':
		learning = importlib.import_module('seq2seq_abcd-vae_b200.plain.learning')
		learner = learning.Learner(args.bins, 256, 256, 256, 16, None, **kw)
	else:
		learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
		learner = learning.Learner(args.bins, 256, 256, 256, CFG['K'], CFG['f'], None, **kw)
	# identical deterministic weights on every rank (and in the CPU arm)
	for i, m in enumerate(learner._modules()):
		m.load_state_dict({k: v.to(dev) for k, v in S.deterministic_state(m, 500 + i).items()})
	learner.setup_optimizer(1.0, 0.0, gradient_clip=1.0, patience=0)
	learner.feature_sampler.train()
	learner.use_cuda_graph = not args.no_graph
	return learner


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
	L = pkg._lib
	learner = build_learner(args, dev)

	def barrier():
		if world > 1:
			torch.distributed.barrier()
		torch.cuda.synchronize()

	def timed(fn, steps):
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

	def workload(B_local, global_B, r, w):
		"""-> (step_resident, step_e2e, lengths, h2d bytes) for this rank's shard of a global batch."""
		lengths, bs, x_host, off_host = make_batch(B_local, args.bins, 0, r, w)
		x_pin, off_pin = x_host.pin_memory(), off_host.pin_memory()
		x_dev, off_dev = x_pin.to(dev), off_pin.to(dev)
		speaker = (torch.arange(len(lengths)) % args.speakers).to(dev) if args.speakers else None
		packed = torch.nn.utils.rnn.PackedSequence(x_dev, bs)
		packed_off = torch.nn.utils.rnn.PackedSequence(off_dev, bs)

		def step_resident():
			return learner.train_step(packed, packed_off, speaker, CFG['N_data'], args.pretrain, global_batch_size=global_B)[0]

		# end-to-end path: every step copies its packed batch from pinned host memory and reads the loss back.  Like a data
		# loader with one batch of look-ahead, the copy of step i+1 is issued on a copy stream before step i computes (two
		# device staging slots); the loss read-back synchronises every step.
		x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
		off_stage = [torch.empty_like(off_dev), torch.empty_like(off_dev)]
		stage_packed = [(torch.nn.utils.rnn.PackedSequence(x_stage[i], bs), torch.nn.utils.rnn.PackedSequence(off_stage[i], bs)) for i in range(2)]
		loss_host = torch.empty((), dtype=torch.float32).pin_memory()
		copy_stream = torch.cuda.Stream(device=dev)
		copied = [torch.cuda.Event(), torch.cuda.Event()]
		st = {'i': 0, 'primed': False}

		def issue_copy(slot):
			with torch.cuda.stream(copy_stream):
				x_stage[slot].copy_(x_pin, non_blocking=True)
				off_stage[slot].copy_(off_pin, non_blocking=True)
				copied[slot].record(copy_stream)

		def step_e2e():
			slot = st['i'] & 1
			if not st['primed']:
				issue_copy(slot)
				st['primed'] = True
			issue_copy(slot ^ 1)   # next step's batch (its slot was last read by the previous step, which has completed)
			torch.cuda.current_stream().wait_event(copied[slot])
			p, po = stage_packed[slot]
			loss = learner.train_step(p, po, speaker, CFG['N_data'], args.pretrain, global_batch_size=global_B)[0]
			loss_host.copy_(loss.detach(), non_blocking=True)
			torch.cuda.current_stream().synchronize()
			st['i'] += 1
			return loss_host
		return step_resident, step_e2e, lengths, x_pin.numel() * 4 + off_pin.numel() * 4

	gb = args.batch * world
	step_resident, step_e2e, lengths, h2d = workload(args.batch, gb, rank, world)
	warm = max(args.warmup, 3)
	for _ in range(warm):
		step_resident()
	clocks = ClockSampler(local_rank)
	if rank == 0:
		clocks.start()
	launches0 = learner.kernel_launches()
	total_ms = timed(step_resident, args.steps)
	launches = learner.kernel_launches() - launches0
	clock_info = clocks.stop() if rank == 0 else None
	ms_per_step = total_ms / args.steps
	value = gb / (ms_per_step * 1e-3)

	for _ in range(6):   # both staging slots seen twice: their steps are recorded before the timed region
		step_e2e()
	e2e_ms = timed(step_e2e, args.steps) / args.steps
	e2e_value = gb / (e2e_ms * 1e-3)

	# strong scaling of the publication batch itself: 512 segments GLOBAL, sharded strided over the ranks
	strong = None
	if world > 1 and not args.no_strong and args.batch % world == 0:
		s_res, s_e2e, s_lengths, _ = workload(args.batch // world, args.batch, rank, world)
		for _ in range(warm):
			s_res()
		s_ms = timed(s_res, args.steps) / args.steps
		for _ in range(6):
			s_e2e()
		s_e2e_ms = timed(s_e2e, args.steps) / args.steps
		strong = {'global_batch': args.batch, 'segments_per_gpu': args.batch // world, 'value': args.batch / (s_ms * 1e-3), 'unit': 'segments/s',
				'ms_per_step': s_ms, 'e2e_value': args.batch / (s_e2e_ms * 1e-3), 'e2e_ms_per_step': s_e2e_ms,
				'note': 'BASELINE.json north star: the B=512 batch sharded over the ranks (rank r takes segments r, r+N, ... of the length-sorted batch)'}

	# kernel-group timing pass (CUDA events recorded by the library on the launching stream; host-launched, no graph)
	graph_was = learner.use_cuda_graph
	learner.use_cuda_graph = False
	for _ in range(2):
		step_resident()
	L.profile_enable(True)
	prof_steps = min(args.steps, 10)
	barrier()
	for _ in range(prof_steps):
		step_resident()
	barrier()
	L.profile_enable(False)
	learner.use_cuda_graph = graph_was
	prof = L.profile_read_all()
	groups = {k: {'ms_per_step': v[0] / prof_steps, 'launches_per_step': v[1] / prof_steps} for k, v in prof.items() if v[1]}

	line = None
	if rank == 0:
		peaks = load_json('MEASURED_PEAKS.json')
		ncu = load_json(os.path.join('profiles', 'ncu_constants.json'))   # written by tools/summarize_profiles.py from the committed capture
		publication = args.batch == CFG['B'] and args.bins == CFG['F'] and args.rnn == 'LSTM' and args.latent == 'abcd' and not args.speakers
		tf32 = measure_tf32_peak(dev)
		exch = measure_exchange(L, dev)
		sumT, T_max = int(lengths.sum()), int(lengths.max())
		g = 4 if args.rnn == 'LSTM' else 3
		F, H, M = args.bins, 256, 256
		kernel_flops = {   # algorithmic FLOPs per launch of each persistent kernel (DESIGN.md "Kernels")
			'enc_fwd_kernel': sumT * 2 * (2 * g * H * H),
			'enc_bwd_kernel': sumT * 2 * (2 * g * H * H),
			'dec_fwd_kernel': sumT * 2 * (g * H * (F + H) + 2 * H * M + 2 * M * F),
			'dec_bwd_kernel': sumT * 2 * (g * H * (F + H) + 2 * H * M + 2 * M * F),
		}
		peak_bf16 = peaks.get('bf16_tflops_sustained', 1400.0)
		peak_tf32 = tf32['tf32_tflops_sustained']
		per_kernel, latency = {}, {}
		for k in PERSISTENT:
			if k not in groups:
				continue
			dur = groups[k]['ms_per_step'] / max(groups[k]['launches_per_step'], 1) * 1e-3
			ach = kernel_flops[k] / dur / 1e12
			nc = ncu.get(k, {}) if publication else {}
			per_kernel[k] = {'avg_launch_ms': dur * 1e3, 'achieved_tflops': ach, 'frac_of_bf16_peak': ach / peak_bf16, 'frac_of_tf32_peak': ach / peak_tf32,
							'frac_of_3xtf32_peak': ach / (peak_tf32 / 3.0), 'frac_of_fp32_ffma_peak': ach / 72.0,
							'traffic': nc.get('dram_bytes'), 'ncu_tensor_pct': nc.get('tensor_pct'), 'ncu_fma_pct': nc.get('fma_pct')}
			n_ex, cl = EXCHANGES[k]
			floor_us = n_ex * exch['cl{}_fanin_ns'.format(cl)] * 1e-3
			us = dur * 1e6 / T_max
			latency[k] = {'us_per_timestep': us, 'exchanges_per_timestep': n_ex, 'cluster_size': cl, 'exchange_floor_us_per_timestep': floor_us,
						'frac_of_exchange_floor': floor_us / us}
		dom = max(per_kernel, key=lambda k: per_kernel[k]['avg_launch_ms'], default=None)
		roofline = None
		if dom is not None:
			pk = per_kernel[dom]
			roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': pk['achieved_tflops'], 'peak': peak_bf16, 'unit': 'TFLOP/s',
						'frac': pk['achieved_tflops'] / peak_bf16, 'traffic': pk['traffic'],
						'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback 1400 (B200_PROFILING.md)',
						'traffic_source': ncu.get('_source') if pk['traffic'] is not None else None,
						'tf32_peak_measured': tf32, 'frac_of_tf32_peak': pk['frac_of_tf32_peak'], 'frac_of_3xtf32_peak': pk['frac_of_3xtf32_peak'],
						'frac_of_fp32_ffma_peak': pk['frac_of_fp32_ffma_peak'], 'avg_launch_ms': pk['avg_launch_ms'],
						'note': 'fp32-grade recurrent kernel (tcgen05 3xTF32 products with the weights in tensor memory + FFMA phases): algorithmic '
								'FLOPs over the measured bf16 tensor peak; the achievable ceiling of a 3xTF32 kernel is tf32_peak / 3; the kernel is '
								'bound by the per-timestep latency chain, see latency_roofline'}
		latency_roofline = {'T_max': T_max, 'kernels': latency, 'exchange_ns': exch,
							'note': 'us per time step of each persistent kernel against the measured cost of its publish -> cluster barrier -> consume '
									'exchanges through L2 (abcd_measure_cluster_exchange, 16-way fan-in rows; the decoder kernels have a third, '
									'barrier-free exchange: st.async pushes into the consumers\' shared memory); the remainder is the dependent chain '
									'operand fetch -> tcgen05.mma / FFMA -> tcgen05.ld -> cell update inside a step'}
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
		if 'rowwise' in groups:
			# sampler row kernels: logits [B,K] read, y [B,K] + per-row KL terms written (forward); the same again backward
			gbs = 4 * args.batch * CFG['K'] * 4 / (groups['rowwise']['ms_per_step'] * 1e-3) / 1e9
			other['sampler_rowwise'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
										'ms_per_step': groups['rowwise']['ms_per_step'], 'launches_per_step': groups['rowwise']['launches_per_step'],
										'note': 'latency-bound: 512 x 128 fp32 rows (256 KB) per kernel, one warp per segment'}
		# front end of the same batch: log-amplitude STFT (8 ms frame / 4 ms hop at 16 kHz -> n_fft 128, hop 64) of synthetic
		# int16 audio straight into packed order; algorithmic bytes = int16 samples read once + fp32 bins written
		try:
			du = importlib.import_module('seq2seq_abcd-vae_b200.modules.data_utils')
			S = importlib.import_module('seq2seq_abcd-vae_b200.synthetic')
			n_fft, hop = 2 * (F - 1), F - 1
			nsamp = torch.tensor([(int(l) - 1) * hop + hop // 2 for l in lengths], dtype=torch.int32)
			onset = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), nsamp[:-1].long()]), 0)
			wav = torch.from_numpy(np.random.RandomState(5).randint(-3000, 3000, size=int(nsamp.sum())).astype(np.int16)).to(dev)
			desc = pkg.model.as_packed_desc(S.batch_sizes_of(lengths), dev)
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
		cpu_baseline = None
		if not args.no_cpu_baseline:
			cval, cdt, _ = run_cpu(args, args.cpu_steps, 1)
			cpu_baseline = {'value': cval, 'unit': 'segments/s', 'cores': torch.get_num_threads(), 'kind': 'port',
							'sample': '1 warm-up + {} timed full steps of the same B={} batch on the host cores ({:.2f} s/step)'.format(args.cpu_steps, args.batch, cdt)}
		fl = flops_per_step(lengths, args.bins, rnn=args.rnn)
		line = {
			'metric': METRIC, 'value': value, 'unit': 'segments/s', 'n_gpus': world, 'steps': args.steps, 'warmup': warm,
			'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
			'data': 'synthetic', 'config': workload_config(args, lengths, world), 'clocks': clock_info,
			'e2e': {'value': e2e_value, 'unit': 'segments/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4, 'ms_per_step': e2e_ms},
			'gpu_launches': int(launches), 'launch_mode': learner.launch_mode(), 'roofline': roofline, 'cpu_baseline': cpu_baseline,
			'strong': strong, 'latency_roofline': latency_roofline, 'rooflines_kernels': per_kernel,
			'kernel_groups': groups, 'rooflines_other': other,
			'train_flops_per_step': fl, 'achieved_tflops_whole_step': fl / (ms_per_step * 1e-3) / 1e12,
		}
	if world > 1:
		torch.distributed.barrier()   # ranks > 0 wait while rank 0 measures peaks and the CPU baseline
	if line is not None:
		emit_json(line)
	if world > 1:
		learner.release_graphs()   # recorded steps hold NCCL kernels: drop them before tearing the communicator down
		torch.distributed.barrier()
		torch.distributed.destroy_process_group()


if __name__ == '__main__':
	main()
