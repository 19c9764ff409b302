# coding: utf-8
"""GPU suite: the CUDA path (through the C ABI, via the nn.Module surface) against the CPU oracle on the same
seeded inputs, and against the committed golden vectors of the reference.

Tolerances (fp32 mode): losses 2e-5 relative; activations / gradients 1e-4 of the tensor's max magnitude
(BASELINE.json north_star: "1e-4 relative in fp32 mode"); packing / arg-max bit-exact.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import abcd_oracle as O
from oracle import cases as CS

pytestmark = pytest.mark.gpu

TOL = 1e-4


def rel_to_max(a, b):
	a = a.detach().double().cpu(); b = b.detach().double().cpu()
	return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _to_cuda_modules(abcd, cfg):
	mods = CS.build_modules(abcd.model, cfg)
	enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
	enc, smp, dec = mods
	enc.load_state_dict(enc_sd); smp.load_state_dict(smp_sd); dec.load_state_dict(dec_sd)
	for m in mods:
		m.cuda()
	smp.temperature = cfg['tau']
	return mods, (enc_sd, smp_sd, dec_sd), seqs, eps, gumbel, speaker


def _cuda_step(abcd, cfg):
	(enc, smp, dec), sds, seqs, eps, gumbel, speaker = _to_cuda_modules(abcd, cfg)
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	is_off = O.is_offset_packed(packed.batch_sizes.tolist())
	data = packed.data.cuda()
	packed_cuda = torch.nn.utils.rnn.PackedSequence(data, packed.batch_sizes)
	last_hidden = enc(packed_cuda)
	if cfg.get('latent') == 'gaussian':   # plain/ variant: (mean, log-variance) -> fused reparameterised sample + KL
		params = smp(last_hidden)
		smp.injected_noise = gumbel.cuda()
		feats, kl = smp.sample_and_kl(params)
		logits = params[0]
	else:
		logits = smp(last_hidden)
		smp.injected_gumbel_noise = gumbel.cuda()
		feats = smp.sample(logits, no_sample=cfg['pretrain'])
		kl = smp.kl_divergence(logits, cfg['N_data'])
	dec.injected_emission_noise = eps.cuda()
	em, off, x_all, (mean, lv), off_logit = dec(feats, batch_sizes=packed.batch_sizes, speaker=speaker.cuda() if speaker is not None else None,
												ground_truth_out=data, ground_truth_offset=is_off.cuda())
	loss = (em + off + kl) / packed.batch_sizes[0]
	loss.backward()
	torch.cuda.synchronize()
	grads = {'encoder': {k: p.grad for k, p in enc.named_parameters()},
			'feature_sampler': {k: p.grad for k, p in smp.named_parameters()},
			'decoder': {k: p.grad for k, p in dec.named_parameters()}}
	out = dict(emission=em, offset=off, kl=kl, loss=loss, logits=logits, last_hidden=last_hidden, features=feats,
				flatten_out=x_all, mean=mean, logvar=lv, offset_logit=off_logit, grads=grads)
	return out, sds, seqs, eps, gumbel, speaker


ALL_CASES = dict(CS.TRAIN_CASES, **CS.EXTRA_CASES, **CS.PLAIN_CASES)


@pytest.mark.parametrize('name', sorted(ALL_CASES))
def test_train_step_matches_oracle_and_golden(abcd, golden_dir, name):
	cfg = ALL_CASES[name]
	out, (enc_sd, smp_sd, dec_sd), seqs, eps, gumbel, speaker = _cuda_step(abcd, cfg)
	dt = torch.float64
	cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
	o = O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dt) for s in seqs], eps.to(dt), gumbel.to(dt),
					cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, speaker, latent=cfg.get('latent', 'abcd'))
	report = {}
	for k in ('emission', 'offset', 'kl', 'loss'):
		report[k] = abs(out[k].item() - o[k].item()) / (abs(o[k].item()) + 1e-30)
	for k in ('last_hidden', 'logits', 'features', 'flatten_out', 'mean', 'logvar', 'offset_logit'):
		report[k] = rel_to_max(out[k], o[k])
	worst = ('', 0.0)
	for m in o['grads']:
		for k, g in o['grads'][m].items():
			got = out['grads'][m][k]
			assert got is not None, 'missing gradient {}/{}'.format(m, k)
			r = rel_to_max(got, g)
			report['grad/{}/{}'.format(m, k)] = r
			if r > worst[1]:
				worst = ('{}/{}'.format(m, k), r)
	print(name, {k: '%.1e' % v for k, v in report.items() if not k.startswith('grad/')}, 'worst grad', worst)
	for k in ('emission', 'offset', 'kl', 'loss'):
		assert report[k] < 2e-5, (k, report[k], out[k].item(), o[k].item())
	for k, v in report.items():
		if k not in ('emission', 'offset', 'loss', 'kl'):
			assert v < TOL, (k, v)
	# arg-max category assignment is bit-exact (oracle fp64 margin is far above fp32 noise in these cases)
	if cfg.get('latent') != 'gaussian':
		assert torch.equal(out['logits'].argmax(-1).cpu(), o['logits'].argmax(-1))
	# golden vectors from the reference itself (the extra multi-tile / odd-size cases are oracle-only)
	if name not in CS.TRAIN_CASES and name not in CS.PLAIN_CASES:
		return
	g = np.load(os.path.join(golden_dir, name + '.npz'))
	for k in ('emission', 'offset', 'loss'):
		assert abs(out[k].item() - float(g[k])) <= 3e-5 * abs(float(g[k])), k
	np.testing.assert_allclose(out['logits'].detach().cpu().numpy(), g['logits'], rtol=0, atol=TOL * np.abs(g['logits']).max())
	for m in out['grads']:
		for k, gr in out['grads'][m].items():
			ref_norm = float(g['gradnorm/{}/{}'.format(m, k)])
			assert abs(gr.norm().item() - ref_norm) <= TOL * ref_norm + 1e-9, (m, k)
			np.testing.assert_allclose(CS.sample_vec(gr), g['gradsample/{}/{}'.format(m, k)], rtol=0, atol=TOL * ref_norm + 1e-12)


def test_learner_train_steps_match_oracle(abcd):
	"""learning.Learner.train_step: backward kernels accumulating straight into the flat gradient arena, fused global-norm
	clip + SGD (+ zeroing) -- two consecutive steps against the fp64 oracle (reference: learning.py:147-163)."""
	import importlib
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	cfg = CS.TRAIN_CASES['train_lstm_gumbel']
	F = cfg['F']
	learner = learning.Learner(F, 256, 256, 256, cfg['K'], cfg['f'], None, device='cuda', seed=3)
	mods = learner._modules()
	enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
	for m, sd in zip(mods, (enc_sd, smp_sd, dec_sd)):
		m.load_state_dict({k: v.cuda() for k, v in sd.items()})
	learner.setup_optimizer(0.5, 0.0, gradient_clip=1.0, patience=0)
	learner.feature_sampler.train()
	learner.feature_sampler.temperature = cfg['tau']
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	is_off = O.is_offset_packed(packed.batch_sizes.tolist())
	p_in = torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes)
	p_off = torch.nn.utils.rnn.PackedSequence(is_off.cuda(), packed.batch_sizes)
	dt = torch.float64
	state = [{k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in sd.items()} for sd in (enc_sd, smp_sd, dec_sd)]
	names = ('encoder', 'feature_sampler', 'decoder')
	for step in range(2):
		learner.feature_sampler.injected_gumbel_noise = gumbel.cuda()
		learner.decoder.injected_emission_noise = eps.cuda()
		loss = learner.train_step(p_in, p_off, None, cfg['N_data'], False)[0]
		o = O.train_step(state[0], state[1], state[2], [x.to(dt) for x in seqs], eps.to(dt), gumbel.to(dt), cfg['N_data'], False,
						cfg['tau'], cfg['rnn'], 1.0, None)
		assert abs(loss.item() - o['loss'].item()) < 2e-5 * abs(o['loss'].item()), step
		keys, params, grads = [], [], []
		for i, n in enumerate(names):
			for k, g in o['grads'][n].items():
				keys.append((i, k)); params.append(state[i][k]); grads.append(g)
		new, total, _ = O.clip_and_sgd(params, grads, 0.5, 1.0)
		for (i, k), p in zip(keys, new):
			state[i][k] = p.detach()
		assert abs(learner.arena.norm.item() - total) < 1e-4 * total
		assert float(learner.arena.grad.abs().max()) == 0.0          # the fused step leaves the gradient arena zeroed
		for i, m in enumerate(mods):
			for k, v in m.state_dict().items():
				if v.dtype.is_floating_point and k != 'prior_concentration':
					assert rel_to_max(v, state[i][k]) < TOL, (step, names[i], k)


def test_encoder_large_ragged_batch_matches_oracle(abcd):
	"""More than one row tile (B > 64), ties, length-1 segments; inference (no-save) path; LSTM and GRU."""
	for rnn, seed in (('LSTM', 11), ('GRU', 12)):
		lengths = sorted([1, 1, 2, 3, 3, 5, 8, 8, 8, 13, 21, 21, 34] * 6 + [40, 37], reverse=True)   # B = 80
		F = 65
		enc = abcd.model.RNN_Variational_Encoder(F, 256, rnn_type=rnn)
		sd = CS.det_state(enc, seed)
		enc.load_state_dict(sd)
		enc.cuda()
		seqs = CS.make_inputs(lengths, F, seed)
		packed = torch.nn.utils.rnn.pack_sequence(seqs)
		with torch.no_grad():
			got = enc(torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes))
		P = {k[len('rnn.'):]: v.double() for k, v in sd.items()}
		want = O.encoder_forward([s.double() for s in seqs], P, rnn)
		assert rel_to_max(got, want) < TOL


@pytest.mark.parametrize('tensorcore', [1, 0])
def test_encoder_recurrence_paths_match_oracle(abcd, tensorcore):
	"""Encoder forward with the per-step gate contraction on tcgen05 (kind::tf32, recurrent weights resident in tensor
	memory, 3xTF32 split; rnn_ts.cuh -- the default) and on the FFMA path: the training (saving) path -- checked through
	the full backward -- and the inference path; same fp32-grade tolerance for both."""
	lib = abcd._lib.load()
	prev = lib.abcd_set_rnn_tensorcore(tensorcore)
	try:
		assert lib.abcd_set_rnn_tensorcore(-1) == tensorcore
		for rnn, seed in (('LSTM', 21), ('GRU', 22)):
			lengths = sorted([1, 2, 3, 5, 8, 8, 13, 21, 34, 55] * 9 + [60, 47, 3], reverse=True)   # B = 93: two row tiles
			F = 65
			enc = abcd.model.RNN_Variational_Encoder(F, 256, rnn_type=rnn)
			sd = CS.det_state(enc, seed)
			enc.load_state_dict(sd)
			enc.cuda()
			seqs = CS.make_inputs(lengths, F, seed)
			packed = torch.nn.utils.rnn.pack_sequence(seqs)
			P = {k[len('rnn.'):]: v.double() for k, v in sd.items()}
			want = O.encoder_forward([s.double() for s in seqs], P, rnn)
			with torch.no_grad():
				got = enc(torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes))
			assert rel_to_max(got, want) < TOL
			# training path: forward saves activations, gradient of sum(out * w) w.r.t. the weights
			wsum = torch.from_numpy(O.det_uniform(tuple(want.shape), seed + 100, -1.0, 1.0))
			out = enc(torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes))
			assert rel_to_max(out.detach(), want) < TOL
			(out * wsum.float().cuda()).sum().backward()
			Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
			(O.encoder_forward([s.double() for s in seqs], Pg, rnn) * wsum.double()).sum().backward()
			for k, v in Pg.items():
				assert rel_to_max(dict(enc.named_parameters())['rnn.' + k].grad, v.grad) < TOL, k
	finally:
		lib.abcd_set_rnn_tensorcore(prev)


def test_encode_argmax_bit_exact(abcd):
	"""Encoder.encode (encode.py:28-32): probabilities within tolerance, category arg-max identical."""
	cfg = dict(rnn='LSTM', F=65, K=128, f=256)
	lengths = CS.encode_lengths(24, 5)
	lengths = np.minimum(lengths, 60)
	lengths = np.sort(lengths)[::-1]
	mods = CS.build_modules(abcd.model, cfg)
	enc, smp, _ = mods
	enc_sd, smp_sd = CS.det_state(enc, 31), CS.det_state(smp, 32)
	enc.load_state_dict(enc_sd); smp.load_state_dict(smp_sd)
	enc.cuda(); smp.cuda()
	seqs = CS.make_inputs([int(l) for l in lengths], 65, 33)
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	with torch.no_grad():
		logits = smp(enc(torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes)))
		B, K = logits.shape
		probs = torch.empty_like(logits)
		amax = torch.empty(B, dtype=torch.int32, device='cuda')
		L = abcd._lib
		L.call('abcd_softmax_argmax', B, K, L.ptr(logits), L.ptr(probs), L.ptr(amax), L.stream_ptr())
	P = {k[len('rnn.'):]: v.double() for k, v in enc_sd.items()}
	lh = O.encoder_forward([s.double() for s in seqs], P, 'LSTM')
	ologits = O.sampler_logits(lh, {k: v.double() for k, v in smp_sd.items()})
	oprobs = O.softmax_rows(ologits)
	top2 = ologits.topk(2, -1).values
	margin = (top2[:, 0] - top2[:, 1]).min().item()
	assert rel_to_max(probs, oprobs) < TOL
	assert margin > 1e-4, 'case too close to call in fp32; pick another seed'
	assert torch.equal(amax.cpu().long(), ologits.argmax(-1))


def test_stft_matches_golden_and_oracle(abcd, golden_dir):
	g = np.load(os.path.join(golden_dir, 'stft_toy.npz'))
	frame, hop = int(g['frame']), int(g['hop'])
	stft = abcd.data_utils.STFT(frame, hop, device='cuda')
	for i in range(2):
		wav = torch.from_numpy(g['wav_%d' % i].astype(np.float32))
		amp = stft(wav).cpu().numpy()
		ref = g['amp_%d' % i]
		assert amp.shape == ref.shape
		np.testing.assert_allclose(amp, ref, rtol=0, atol=2e-6 * ref.max())
		np.testing.assert_allclose(amp, O.stft_amplitude(g['wav_%d' % i], frame, hop), rtol=0, atol=2e-6 * ref.max())
		logn = abcd.data_utils.LogSTFT(stft, 2.0 ** -15, 11.0)(torch.from_numpy(g['wav_%d' % i]))   # int16 path
		np.testing.assert_allclose(logn.cpu().numpy(), g['logn_%d' % i], rtol=0, atol=2e-3)


def test_sgemm_layouts(abcd):
	L = abcd._lib
	L.ensure_device(torch.device('cuda'))
	torch.manual_seed(0)
	for (M, N, K) in ((70, 33, 65), (257, 130, 1000), (1024, 65, 3000), (5, 256, 1)):
		A = torch.randn(M, K, device='cuda'); Bm = torch.randn(K, N, device='cuda')
		want = (A.double() @ Bm.double()).float()
		for a_km, b_km in ((1, 1), (1, 0), (0, 0)):
			Ain = A.contiguous() if a_km else A.t().contiguous()
			Bin = Bm.t().contiguous() if b_km else Bm.contiguous()
			for splits in (1, 4):
				Cout = torch.zeros(M, N, device='cuda')
				part = torch.empty(splits * M * N, device='cuda')
				L.call('abcd_sgemm', M, N, K, L.ptr(Ain), Ain.shape[1], a_km, L.ptr(Bin), Bin.shape[1], b_km, L.ptr(Cout), N,
					1.0, None, 0, 0, splits, L.ptr(part), L.stream_ptr())
				torch.cuda.synchronize()
				assert rel_to_max(Cout, want) < 2e-6 * max(1, K ** 0.5), (M, N, K, a_km, b_km, splits)


def test_umma_rowmajor_gemm_3xtf32(abcd):
	"""tcgen05 3xTF32 GEMM for tall row-major activations (NT: input projection / MLP layers, NN: data gradients), fused
	bias + scale + tanh epilogue, ragged last row tile and column block."""
	L = abcd._lib
	L.ensure_device(torch.device('cuda'))
	torch.manual_seed(2)
	for (M, N, K, act) in ((14755, 1024, 68, 0), (4100, 256, 256, 1), (2048, 324, 1024, 0), (3000, 64, 32, 1)):
		A = torch.randn(M, K, device='cuda'); Bm = torch.randn(K, N, device='cuda') * 0.2
		bias = torch.randn(N, device='cuda')
		want = 0.5 * (A.double() @ Bm.double()) + bias.double()
		scale = float(want.abs().max())   # error budget is relative to the largest pre-activation
		if act:
			want = torch.tanh(want)
		for b_km in (1, 0):
			Bin = Bm.t().contiguous() if b_km else Bm.contiguous()
			Cout = torch.full((M, N), 7.0, device='cuda')
			L.call('abcd_sgemm', M, N, K, L.ptr(A), K, 1, L.ptr(Bin), Bin.shape[1], b_km, L.ptr(Cout), N,
				0.5, L.ptr(bias), act, 0, 1, None, L.stream_ptr())
			torch.cuda.synchronize()
			err = float((Cout.double() - want).abs().max()) / scale
			assert err < 3e-6 + 1.5e-8 * K, (M, N, K, act, b_km, err)


def test_clip_sgd_matches_golden(abcd, golden_dir):
	L = abcd._lib
	L.ensure_device(torch.device('cuda'))
	g = np.load(os.path.join(golden_dir, 'optim.npz'))
	shapes = [(300, 7), (1024,), (33, 65), (5,)]
	npart = L.load().abcd_optim_partial_count()
	for mom, tag in ((0.0, 'plain'), (0.9, 'momentum')):
		flat = torch.cat([O.det_tensor(s, 900 + i, 'uniform', 0.5).reshape(-1) for i, s in enumerate(shapes)]).cuda()
		mbuf = torch.zeros_like(flat) if mom else None
		partial = torch.empty(npart, device='cuda'); norm = torch.empty((), device='cuda')
		for step in range(2):
			gr = torch.cat([O.det_tensor(s, 950 + 10 * step + i, 'normal', 0.3).reshape(-1) for i, s in enumerate(shapes)]).cuda()
			L.call('abcd_gradnorm_clip_sgd', L.ptr(flat), L.ptr(gr), L.ptr(mbuf), flat.numel(), 0.7, mom, 1.0, 1, L.ptr(norm), L.ptr(partial), L.stream_ptr())
			assert abs(norm.item() - float(g['{}_norm_{}'.format(tag, step)])) < 1e-4
			assert float(gr.abs().max()) == 0.0   # zero_grads
		np.testing.assert_allclose(flat.cpu().numpy(), g[tag + '_params'], rtol=0, atol=2e-6)


def test_plain_gaussian_latent(abcd, golden_dir):
	g = np.load(os.path.join(golden_dir, 'plain_latent.npz'))
	f = 16
	smp = abcd.model.Sampler(1024, 256, f)
	sd = CS.det_state(smp, 77)
	smp.load_state_dict(sd)
	smp.cuda()
	x = O.det_tensor((5, 1024), 771, 'uniform', 1.0).cuda().requires_grad_(True)
	eps = O.det_tensor((5, f), 772, 'normal', 1.0).cuda()
	w = O.det_tensor((5, f), 773, 'uniform', 1.0).cuda()
	mean, lv = smp(x)
	z, kl = abcd.model._GaussLatentFn.apply(mean, lv, eps)
	((z * w).sum() + kl).backward()
	np.testing.assert_allclose(z.detach().cpu().numpy(), g['z'], rtol=0, atol=TOL * np.abs(g['z']).max())
	assert abs(kl.item() - float(g['kl'])) < 2e-5 * abs(float(g['kl']))
	np.testing.assert_allclose(x.grad.cpu().numpy(), g['dx'], rtol=0, atol=TOL * np.abs(g['dx']).max())


def test_philox_moments(abcd):
	n = 1 << 20
	t = torch.empty(n, device='cuda')
	a = abcd.model.noise_like(t)
	b = abcd.model.noise_like(t)
	assert abs(a.mean().item()) < 5e-3 and abs(a.std().item() - 1.0) < 5e-3
	assert not torch.equal(a, b)
	gmb = abcd.model.noise_like(t, kind='gumbel')
	assert abs(gmb.mean().item() - 0.5772) < 5e-3


def test_unsupported_sizes_fail_loudly(abcd):
	enc = abcd.model.RNN_Variational_Encoder(65, 128, rnn_type='LSTM').cuda()
	x = torch.nn.utils.rnn.pack_sequence([torch.randn(4, 65), torch.randn(2, 65)])
	with pytest.raises(abcd._lib.AbcdError):
		enc(torch.nn.utils.rnn.PackedSequence(x.data.cuda(), x.batch_sizes))
	with pytest.raises(NotImplementedError):
		abcd.model.RNN_Variational_Encoder(65, 256, rnn_type='ESN')


def test_umma_tn_gemm_3xtf32(abcd):
	"""tcgen05 3xTF32 weight-gradient GEMM (TN form, row gather, ragged N) keeps fp32 fidelity."""
	L = abcd._lib
	L.ensure_device(torch.device('cuda'))
	torch.manual_seed(1)
	# (M, N, K, ldb, gather, lda, first column of A): the last two exercise a ragged last M tile inside a wider row
	# (the emission head's weight gradient: M = F = 65 columns of the [sumT, 2 * 68] gradient buffer)
	for (M, N, K, ldb, gather, lda, a_col) in ((1024, 256, 5000, 256, True, 1024, 0), (1024, 65, 3001, 65, False, 1024, 0),
												(256, 256, 777, 256, False, 256, 0), (128, 17, 300, 20, True, 128, 0),
												(1024, 65, 14755, 68, True, 1024, 0), (65, 256, 14755, 256, False, 136, 68),
												(200, 130, 1000, 132, False, 200, 0)):
		Afull = torch.randn(K, lda, device='cuda')
		A = Afull[:, a_col:a_col + M]
		Bfull = torch.randn(K + 7, ldb, device='cuda')
		idx = torch.randint(-1, K + 7, (K,), device='cuda', dtype=torch.int32) if gather else None
		rows = Bfull[:K] if idx is None else torch.where((idx >= 0)[:, None], Bfull[idx.clamp(min=0).long()], torch.zeros_like(Bfull[:K]))
		want = (A.double().t() @ rows[:, :N].double())
		splits = L.load().abcd_sgemm  # noqa: F841
		for nsplit in (1, 5, 37):
			Cout = torch.full((M, N), 0.5, device='cuda')
			part = torch.empty(nsplit * M * N, device='cuda')
			import ctypes as C
			lib = L.load()
			# abcd_sgemm has no gather argument: exercise the gather through the encoder/decoder paths; here plain rows
			if idx is None:
				L.call('abcd_sgemm', M, N, K, L.ptr(Afull) + 4 * a_col, lda, 0, L.ptr(Bfull), ldb, 0, L.ptr(Cout), N, 2.0, None, 0, 1, nsplit, L.ptr(part), L.stream_ptr())
				torch.cuda.synchronize()
				got = Cout.double() - 0.5
				err = float((got - 2.0 * want).abs().max() / (2.0 * want).abs().max())
				# tensor-memory accumulation truncates (RZ): the error grows with the rows accumulated per CTA; the
				# production split keeps that near 400 rows (a few 1e-6)
				assert err < 3e-6 + 1.5e-8 * (K / nsplit), (M, N, K, nsplit, err)


def test_full_size_shards_add_up_and_step_is_deterministic(abcd):
	"""BASELINE.json's bench configuration (B = 512 segments of 0.03-0.3 s, F = 65, H = M = f = 256, K = 128), too large for
	the fp64 oracle to finish in seconds, checked through size-independent properties of the reference's loss
	(learning.py:147-163: sums over segments / batch size):
	  * determinism: the same step twice gives bit-identical losses and gradients (no atomics on the product path);
	  * additivity over segments: losses and gradients of the full batch equal the sums over its two strided shards
	    (dist.shard_packed -- exactly what two ranks compute before the all-reduce), to fp32 summation-order round-off."""
	import importlib
	dist = importlib.import_module('seq2seq_abcd-vae_b200.dist')
	F, K, B = 65, 128, 512
	lengths = CS.synthetic_lengths(B, 1000)
	rng = np.random.RandomState(11)
	sumT = int(lengths.sum())
	bs = torch.tensor(O.batch_sizes_from_lengths(lengths.tolist()), dtype=torch.int64)
	x = torch.from_numpy(np.clip(rng.normal(0.45, 0.25, size=(sumT, F)), -0.945, 1.05).astype(np.float32))
	eps = torch.from_numpy(rng.normal(size=(sumT, F)).astype(np.float32))
	gumbel = torch.from_numpy(rng.gumbel(size=(B, K)).astype(np.float32))
	is_off = O.is_offset_packed(bs.tolist())
	torch.manual_seed(5)
	enc = abcd.model.RNN_Variational_Encoder(F, 256).cuda()
	smp = abcd.model.ABCDSampler(enc.hidden_size_total, 256, K, 256).cuda()
	dec = abcd.model.RNN_Variational_Decoder(F, 256, 256, 256).cuda()
	smp.train(); smp.temperature = 0.8
	mods = (enc, smp, dec)

	def run(rank, world):
		for m in mods:
			m.zero_grad(set_to_none=True)
		d, b = dist.shard_packed(x, bs, rank, world)
		e, _ = dist.shard_packed(eps, bs, rank, world)
		o, _ = dist.shard_packed(is_off, bs, rank, world)
		g = gumbel[dist.shard_indices(B, rank, world)]
		d = d.cuda()
		logits = smp(enc(torch.nn.utils.rnn.PackedSequence(d, b)))
		smp.injected_gumbel_noise = g.cuda()
		feats = smp.sample(logits)
		kl = smp.kl_divergence(logits, 40000)
		dec.injected_emission_noise = e.cuda()
		em, off = dec(feats, batch_sizes=b, ground_truth_out=d, ground_truth_offset=o.cuda())[:2]
		((em + off + kl) / B).backward()
		torch.cuda.synchronize()
		losses = torch.stack([em.detach(), off.detach(), kl.detach()]).double().cpu()
		grads = [p.grad.detach().clone() for m in mods for p in m.parameters() if p.grad is not None]
		run.names = [type(m).__name__ + '.' + k for m in mods for k, p in m.named_parameters() if p.grad is not None]
		return losses, grads

	l_full, g_full = run(0, 1)
	l_again, g_again = run(0, 1)
	assert torch.equal(l_full, l_again)
	unequal = [(n, rel_to_max(a, b)) for n, a, b in zip(run.names, g_full, g_again) if not torch.equal(a, b)]
	assert not unequal, unequal
	l0, g0 = run(0, 2)
	l1, g1 = run(1, 2)
	assert len(g_full) == len(g0) == len(g1) > 20
	# the Dirichlet term of the KL is weighted by (segments in the batch) / N_data (model.py:327-329), so it adds up too
	assert torch.allclose(l0 + l1, l_full, rtol=2e-6, atol=0), (l0 + l1, l_full)
	worst = max(rel_to_max(a + b, f) for a, b, f in zip(g0, g1, g_full))
	assert worst < 2e-5, worst


@pytest.mark.parametrize('B,K,f,mode', [(93, 128, 256, 1), (512, 128, 256, 0), (7, 33, 40, 1), (200, 256, 96, 1), (5, 16, 512, 0)])
def test_fused_sampler_matches_oracle_and_unfused(abcd, B, K, f, mode):
	"""abcd_sampler_fused_fwd / _bwd (one kernel each way: logits, Gumbel-softmax or softmax, codebook features, Dirichlet KL)
	against the fp64 oracle (model.py:581-639) and against the separate forward / sample / kl_divergence calls."""
	tau, N = 0.7, 3000.0
	smp = abcd.model.ABCDSampler(64, 32, K, f).cuda()
	sd = CS.det_state(smp, 400 + B)
	smp.load_state_dict(sd)
	smp.cuda()
	smp.temperature = tau
	assert smp.fused_supported()
	x = O.det_tensor((B, 64), 4100 + B, 'uniform', 1.0)
	gum = O.det_tensor((B, K), 4200 + B, 'gumbel', 1.0)
	wf = O.det_tensor((B, f), 4300 + B, 'uniform', 1.0)

	def grads():
		return {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in smp.named_parameters()}
	# fused
	smp.zero_grad(set_to_none=True)
	xg = x.cuda().requires_grad_(True)
	smp.injected_gumbel_noise = gum.cuda()
	feats, kl, logits = smp.forward_sample_kl(xg, N, no_sample=(mode == 0))
	((feats * wf.cuda()).sum() + 0.37 * kl).backward()
	fused = dict(feats=feats.detach(), kl=kl.detach(), logits=logits.detach(), dx=xg.grad.clone(), grads=grads())
	# separate calls
	smp.zero_grad(set_to_none=True)
	xg2 = x.cuda().requires_grad_(True)
	lg = smp(xg2)
	smp.injected_gumbel_noise = gum.cuda()
	ft = smp.sample(lg, no_sample=(mode == 0))
	k2 = smp.kl_divergence(lg, N)
	((ft * wf.cuda()).sum() + 0.37 * k2).backward()
	sep = dict(feats=ft.detach(), kl=k2.detach(), logits=lg.detach(), dx=xg2.grad.clone(), grads=grads())
	# oracle, fp64
	P = {k: v.double().clone().requires_grad_(v.dtype.is_floating_point and k != 'prior_concentration') for k, v in sd.items()}
	xo = x.double().clone().requires_grad_(True)
	ol = O.sampler_logits(xo, P)
	of, _ = O.sampler_sample(ol, P, mode == 0, tau, gum.double())
	ok = O.sampler_kl(ol, P, N, 1.0)
	((of * wf.double()).sum() + 0.37 * ok).backward()
	for name, got in (('fused', fused), ('separate', sep)):
		assert rel_to_max(got['logits'], ol) < 1e-5, name
		assert rel_to_max(got['feats'], of) < 1e-5, name
		assert abs(got['kl'].item() - ok.item()) < 2e-5 * abs(ok.item()), (name, got['kl'].item(), ok.item())
		assert rel_to_max(got['dx'], xo.grad) < 1e-4, name
		for k, g in got['grads'].items():
			assert g is not None, (name, k)
			assert rel_to_max(g, P[k].grad) < 1e-4, (name, k, rel_to_max(g, P[k].grad))
	# in-kernel Philox noise draws the same numbers as noise_like() on a [B, K] tensor
	if mode == 1:
		abcd.model.manual_noise_seed(99)
		with torch.no_grad():
			f1, _, _ = smp.forward_sample_kl(x.cuda(), N)
		abcd.model.manual_noise_seed(99)
		with torch.no_grad():
			f2 = smp.sample(smp(x.cuda()))
		assert rel_to_max(f1, f2) < 1e-5
