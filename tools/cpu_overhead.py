# coding: utf-8
"""Host-side cost of one training step: time to enqueue a step (no sync) vs. GPU time, and the idle gap a per-step
synchronisation exposes.  python tools/cpu_overhead.py"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
pkg = importlib.import_module('seq2seq_abcd-vae_b200')
learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
from oracle import cases as CS
dev = torch.device('cuda')
lengths, bs, x, off = B.make_batch(512, 65, 0)
ln = learning.Learner(65, 256, 256, 256, 128, 256, None, device='cuda', seed=1)
ln.setup_optimizer(1.0, 0.0, gradient_clip=1.0, patience=0)
ln.feature_sampler.train()
p = torch.nn.utils.rnn.PackedSequence(x.to(dev), bs); po = torch.nn.utils.rnn.PackedSequence(off.to(dev), bs)
for _ in range(5):
	ln.train_step(p, po, None, 100000, False)
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(30):
	torch.cuda.synchronize()
	t0 = time.perf_counter()
	loss = ln.train_step(p, po, None, 100000, False)[0]
	t1 = time.perf_counter()
	v = loss.item()
	t2 = time.perf_counter()
	enq.append(t1 - t0); tot.append(t2 - t0)
print('enqueue %.3f ms, step with read-back %.3f ms (median of 30)' % (1e3 * np.median(enq), 1e3 * np.median(tot)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
	ln.train_step(p, po, None, 100000, False)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)

# --- variants of the end-to-end loop (pinned host batch -> device every step) ---
x_pin, off_pin = x.pin_memory(), off.pin_memory()
xs = [torch.empty_like(p.data), torch.empty_like(p.data)]; os_ = [torch.empty_like(po.data), torch.empty_like(po.data)]
cs = torch.cuda.Stream(); evs = [torch.cuda.Event(), torch.cuda.Event()]
loss_host = torch.empty((), dtype=torch.float32).pin_memory()
def run(variant, n=30):
	ts = []
	for i in range(n):
		torch.cuda.synchronize(); t0 = time.perf_counter()
		s = i & 1
		if variant == 'instream':
			xs[s].copy_(x_pin, non_blocking=True); os_[s].copy_(off_pin, non_blocking=True)
		elif variant == 'prefetch':
			with torch.cuda.stream(cs):
				xs[s ^ 1].copy_(x_pin, non_blocking=True); os_[s ^ 1].copy_(off_pin, non_blocking=True); evs[s ^ 1].record(cs)
			if i > 0:
				torch.cuda.current_stream().wait_event(evs[s])
		pp = torch.nn.utils.rnn.PackedSequence(xs[s] if variant != 'none' else p.data, bs)
		ppo = torch.nn.utils.rnn.PackedSequence(os_[s] if variant != 'none' else po.data, bs)
		loss = ln.train_step(pp, ppo, None, 100000, False)[0]
		loss_host.copy_(loss.detach(), non_blocking=True)
		torch.cuda.current_stream().synchronize()
		ts.append(time.perf_counter() - t0)
	print('%-10s %.3f ms/step' % (variant, 1e3 * np.median(ts)))
xs[0].copy_(x_pin); os_[0].copy_(off_pin); xs[1].copy_(x_pin); os_[1].copy_(off_pin)
for v in ('none', 'instream', 'prefetch', 'none'):
	run(v)
