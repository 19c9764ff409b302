# coding: utf-8
"""GPU diagnostics: cluster residency and per-kernel-group timing of one training step at several batch sizes.
    python tools/diag.py [B ...]"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	L = pkg._lib
	from oracle import cases as CS
	dev = torch.device('cuda', 0)
	L.ensure_device(dev)
	lib = L.load()
	print('max active clusters (enc_fwd, enc_bwd, dec_fwd, dec_bwd) F=65:', [lib.abcd_max_active_clusters(i, 65) for i in range(4)],
		' F=129:', [lib.abcd_max_active_clusters(i, 129) for i in range(4)])
	sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512]
	for B in sizes:
		lengths, bs, x, off = bench.make_batch(B, 65, 0)
		learner = learning.Learner(65, 256, 256, 256, 128, 256, None, device='cuda:0')
		for i, m in enumerate(learner._modules()):
			m.load_state_dict({k: v.to(dev) for k, v in CS.det_state(m, 500 + i).items()})
		learner.setup_optimizer(1.0, 0.0, gradient_clip=1.0)
		p = torch.nn.utils.rnn.PackedSequence(x.to(dev), bs)
		po = torch.nn.utils.rnn.PackedSequence(off.to(dev), bs)
		for _ in range(3):
			learner.train_step(p, po, None, 100000, False)
		torch.cuda.synchronize()
		L.profile_enable(True)
		n = 5
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		e0.record()
		for _ in range(n):
			learner.train_step(p, po, None, 100000, False)
		e1.record()
		torch.cuda.synchronize()
		L.profile_enable(False)
		prof = L.profile_read_all()
		T = int(lengths.max())
		print('B={} sumT={} T_max={} step {:.2f} ms :: '.format(B, int(lengths.sum()), T, e0.elapsed_time(e1) / n) +
			', '.join('{} {:.3f} ms ({:.1f} us/step)'.format(k, v[0] / n, v[0] / n / T * 1e3) for k, v in prof.items() if v[1]))


if __name__ == '__main__':
	main()


def enc_timing():
	import ctypes as C
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	out = (C.c_ulonglong * 10)()
	pkg._lib.load().abcd_debug_enc_timing(out)
	v = list(out)
	if v[4]:
		print('enc_fwd CTA0 cycles/step: setup %.0f gemm %.0f epilogue %.0f barrier %.0f (steps %d)' % (v[0] / v[4], v[1] / v[4], v[2] / v[4], v[3] / v[4], v[4]))
		print("   phase_gemm per step: stage-issue %.0f wait+sync %.0f mma %.0f tail-sync %.0f chunks/step %.2f" % (v[5] / v[4], v[6] / v[4], v[7] / v[4], v[9] / v[4], v[8] / v[4]))
		print("   (tensor-core phase: prologue %.0f stage-wait %.0f split %.0f fence+sync %.0f issue %.0f)" % tuple(x / v[4] for x in v[5:10]))


def dec_timing():
	import ctypes as C
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	out = (C.c_ulonglong * 16)()
	pkg._lib.load().abcd_debug_dec_timing(out)
	v = list(out)
	if v[15]:
		n = float(v[15])
		if os.environ.get('ABCD_DEC_FWD', 'ahead') == 'ahead':   # one-step-ahead pair kernel (default)
			names = ['wait C', 'P1 x product', 'P1 cell + publish', 'barrier A', 'P2 gemm', 'P2 store + P3a', 'read-back + send', 'barrier B', 'P3b + x push', 'arrive C + saves + recv']
		else:
			names = ['P1 setup', 'P1 gemm', 'P1 cell', 'barrier', 'P2 gemm', 'P2 store', 'barrier', 'P3 setup+gemm', 'P3 sample', 'barrier']
		print('dec_fwd CTA0 cycles/step: ' + ', '.join('%s %.0f' % (nm, x / n) for nm, x in zip(names, v[:10])) + ' (steps %d, total %.0f)' % (v[15], sum(v[:10]) / n))
		print('   slots 10..14 (ahead: MMA issue cycles of the issuing thread, arrive A, arrive B, arrive C / others: P1 gemm detail): ' + ' '.join('%.0f' % (x / n) for x in v[10:15]))


if os.environ.get('ABCD_TIMING') == '3':
	# the last dec launch was the backward kernel
	import ctypes as C
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	out = (C.c_ulonglong * 16)()
	pkg._lib.load().abcd_debug_dec_timing(out)
	v = list(out); n = float(max(v[15], 1))
	names = ['prefetch', 'B3 gemm', 'B3 store', 'B2 local gemm', 'barrier', 'B2 reduce', 'pointwise', 'B1 local gemm', 'barrier', 'B1 reduce', 'barrier']
	print('dec_bwd CTA0 cycles/step: ' + ', '.join('%s %.0f' % (nm, x / n) for nm, x in zip(names, v[:11])) + ' (steps %d, total %.0f)' % (v[15], sum(v[:11]) / n))
	print('   B1 detail (push kernel; slots 10, 11 are part of the step but not of the list above): relayout (<= 32 rows) %.0f, MMA issue -> done %.0f' % (v[10] / n, v[11] / n))
elif os.environ.get('ABCD_TIMING'):
	enc_timing()
	dec_timing()
