# coding: utf-8
"""CPU suite, part 1: the oracle restatement against the committed golden vectors (produced by the reference
itself, oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import abcd_oracle as O
from oracle import cases as CS


class _Shapes(object):
	"""Minimal stand-ins exposing state_dict() key/shape sets of the reference modules (no CUDA needed)."""

	def __init__(self, abcd, cfg):
		self.enc, self.smp, self.dec = CS.build_modules(abcd.model, cfg)


def _run_oracle(abcd, cfg, dtype):
	mods = CS.build_modules(abcd.model, cfg)
	enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
	cast = lambda d: {k: (v.to(dtype) if v.dtype.is_floating_point else v) for k, v in d.items()}
	return O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dtype) for s in seqs], eps.to(dtype), gumbel.to(dtype),
						cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, speaker, latent=cfg.get('latent', 'abcd'))


@pytest.mark.parametrize('name', sorted(CS.TRAIN_CASES) + sorted(CS.PLAIN_CASES))
def test_oracle_reproduces_reference_golden(abcd, golden_dir, name):
	cfg = CS.TRAIN_CASES[name] if name in CS.TRAIN_CASES else CS.PLAIN_CASES[name]
	g = np.load(os.path.join(golden_dir, name + '.npz'))
	o = _run_oracle(abcd, cfg, torch.float64)
	for k in ('emission', 'offset', 'loss'):
		assert abs(o[k].item() - float(g[k])) <= 2e-6 * abs(float(g[k])), k
	# the reference evaluates lgamma/digamma in fp32: its KL carries ~1e-5 relative rounding noise
	assert abs(o['kl'].item() - float(g['kl'])) <= 5e-5 * abs(float(g['kl'])) + 1e-4
	np.testing.assert_allclose(o['logits'].numpy(), g['logits'], rtol=0, atol=2e-6 * np.abs(g['logits']).max())
	np.testing.assert_allclose(CS.sample_vec(o['last_hidden']), g['last_hidden_sample'], rtol=0, atol=1e-6)
	np.testing.assert_allclose(CS.sample_vec(o['flatten_out']), g['flatten_out_sample'], rtol=0, atol=2e-6 * np.abs(g['flatten_out_sample']).max())
	for m in o['grads']:
		for k, gr in o['grads'][m].items():
			ref_norm = float(g['gradnorm/{}/{}'.format(m, k)])
			ref_s = g['gradsample/{}/{}'.format(m, k)]
			assert abs(gr.norm().item() - ref_norm) <= 2e-5 * ref_norm + 1e-9, (m, k)
			np.testing.assert_allclose(CS.sample_vec(gr), ref_s, rtol=0, atol=5e-6 * max(np.abs(ref_s).max(), ref_norm, 1e-12), err_msg='{}/{}'.format(m, k))


def test_stft_oracle_matches_reference_front_end(golden_dir):
	g = np.load(os.path.join(golden_dir, 'stft_toy.npz'))
	frame, hop = int(g['frame']), int(g['hop'])
	assert (frame, hop) == (128, 64)
	for i in range(2):
		amp = O.stft_amplitude(g['wav_%d' % i], frame, hop)
		ref = g['amp_%d' % i]
		assert amp.shape == ref.shape == (1 + len(g['wav_%d' % i]) // hop, frame // 2 + 1)
		np.testing.assert_allclose(amp, ref, rtol=0, atol=1e-6 * ref.max())
		# log domain: fp32 FFT rounding of the reference is amplified at near-silent bins
		logn = O.log_normalize(amp, 2.0 ** -15, 11.0)
		np.testing.assert_allclose(logn, g['logn_%d' % i], rtol=0, atol=2e-3)


def test_discrete_bounds_round_half_even():
	assert O.discrete_bounds(0.383629, 0.81, 16000) == (6138, 12960, 6822)
	assert O.discrete_bounds(0.00003125, 0.00009375, 16000)[:2] == (0, 2)   # 0.5 -> 0, 1.5 -> 2


def test_pack_matches_torch_pack_sequence():
	seqs = [torch.randn(l, 3) for l in (7, 5, 5, 2, 1)]
	data, bs, offsets = O.pack(seqs)
	ref = torch.nn.utils.rnn.pack_sequence(seqs)
	assert bs == ref.batch_sizes.tolist()
	assert torch.equal(data, ref.data)
	back = O.unpack(data, bs)
	assert all(torch.equal(a, b) for a, b in zip(back, seqs))
	io = O.is_offset_packed(bs)
	ref_off = torch.nn.utils.rnn.pack_sequence([torch.tensor([0.0] * (s.size(0) - 1) + [1.0]) for s in seqs]).data
	assert torch.equal(io, ref_off)


def test_kl_closed_form_gradients_match_autograd():
	"""Closed forms used by the CUDA backward vs autograd on the oracle, fp64."""
	torch.manual_seed(0)
	B, K, N, c = 9, 16, 57.0, 0.7
	logits = torch.randn(B, K, dtype=torch.float64, requires_grad=True)
	phi = torch.randn(K, dtype=torch.float64, requires_grad=True)
	kl = O.sampler_kl(logits, {'posterior_shape_logits': phi}, N, c)
	kl.backward()
	with torch.no_grad():
		s = torch.softmax(phi, 0)
		alpha = s * N + c
		E = torch.digamma(alpha) - torch.digamma(alpha.sum())
		q = torch.softmax(logits, -1); lq = torch.log_softmax(logits, -1)
		inner = (q * (lq - E)).sum(-1, keepdim=True)
		dlogits = q * ((lq - E) - inner)
		h = torch.polygamma(1, alpha) * (B * s - q.sum(0))
		dphi = N * s * (h - (s * h).sum())
	assert torch.allclose(dlogits, logits.grad, atol=1e-12)
	assert torch.allclose(dphi, phi.grad, atol=1e-10)


def test_clip_and_sgd_matches_torch_golden(golden_dir):
	g = np.load(os.path.join(golden_dir, 'optim.npz'))
	shapes = [(300, 7), (1024,), (33, 65), (5,)]
	for mom, tag in ((0.0, 'plain'), (0.9, 'momentum')):
		params = [O.det_tensor(s, 900 + i, 'uniform', 0.5) for i, s in enumerate(shapes)]
		bufs = None
		for step in range(2):
			grads = [O.det_tensor(tuple(p.shape), 950 + 10 * step + i, 'normal', 0.3) for i, p in enumerate(params)]
			params, norm, bufs = O.clip_and_sgd(params, grads, 0.7, 1.0, mom, bufs)
			assert abs(norm - float(g['{}_norm_{}'.format(tag, step)])) < 1e-4
			if mom == 0.0:
				bufs = None
		flat = torch.cat([p.reshape(-1) for p in params]).numpy()
		np.testing.assert_allclose(flat, g[tag + '_params'], rtol=0, atol=2e-6)


def test_det_generators_are_stable():
	"""Bit-stability of the case generator across machines: integer hash -> fixed known values."""
	u = O.det_uniform((4,), 123)
	np.testing.assert_allclose(u, O.det_uniform((4,), 123))
	assert 0.0 < u.min() and u.max() < 1.0
	assert len(set(np.round(u, 12))) == 4


def test_batched_oracle_equals_per_sequence_oracle(abcd):
	"""The time-major batched restatement (CPU-baseline timing path) is the same function."""
	for name in ('train_lstm_gumbel', 'train_gru_speaker'):
		cfg = CS.TRAIN_CASES[name]
		mods = CS.build_modules(abcd.model, cfg)
		enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
		dt = torch.float64
		cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
		args = (cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dt) for s in seqs], eps.to(dt), gumbel.to(dt),
				cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, speaker)
		a = O.train_step(*args)
		b = O.train_step(*args, batched=True)
		assert abs(a['loss'].item() - b['loss'].item()) < 1e-10 * abs(a['loss'].item())
		for m in a['grads']:
			for k in a['grads'][m]:
				assert torch.allclose(a['grads'][m][k], b['grads'][m][k], atol=1e-10), (m, k)
