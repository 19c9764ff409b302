# coding: utf-8
"""GPU suite, part 3: parity AT BASELINE.json's own configurations ("do this" item 1).

  * configs[2] / configs[3]: the exact publication batch bench.py times (512 segments of 0.03-0.3 s, sumT = 14 755, LSTM
    Gumbel-softmax) and the GRU + speaker-embedding batch of the same size, one full training step through the C ABI against
    the oracle's batched fp64 restatement (reference: ABCD-VAE/learning.py:147-163): losses 2e-5, every activation and every
    gradient 1e-4 of the tensor maximum, arg-max categories bit-exact outside fp32 noise (top-2 margin reported);
  * configs[4]: bulk encode over >= 10 k segments of 0.05-2 s (T up to 501, no clipping) against
    encoder_forward_batched + sampler_logits in fp64 (reference: ABCD-VAE/encode.py:23-35): class probabilities within 1e-4,
    arg-max agreement with the oracle's top-2 margin histogram, failing only where the margin exceeds fp32 noise;
  * learning.py:171-178: the per-batch perplexities kernel against the oracle;
  * 20 consecutive optimizer steps (Learner.train_step: arena accumulation + fused clip/SGD) against oracle + clip_and_sgd.
"""
import importlib
import os
import time

import numpy as np
import pytest
import torch

from oracle import abcd_oracle as O
from oracle import cases as CS

pytestmark = pytest.mark.gpu

TOL = 1e-4
FULL_CASES = {
	# BASELINE.json configs[2]: the batch bench.py runs (synthetic_lengths(512, 1000): T in [8, 75], sumT = 14 755)
	'publication_lstm_gumbel': dict(rnn='LSTM', F=65, K=128, f=256, pretrain=False, tau=0.8, N_data=100000, seed=41),
	# BASELINE.json configs[3]: -R GRU --speaker_embed_dim 256, 16 speakers
	'publication_gru_speaker': dict(rnn='GRU', F=65, K=128, f=256, pretrain=False, tau=0.9, N_data=100000, seed=42,
									num_speakers=16, E=256),
	# pretraining epochs of configs[2] (softmax in place of the Gumbel-softmax sample)
	'publication_lstm_pretrain': dict(rnn='LSTM', F=65, K=128, f=256, pretrain=True, tau=1.0, N_data=100000, seed=43),
}


def rel_to_max(a, b):
	a = a.detach().double().cpu(); b = b.detach().double().cpu()
	return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def margin_histogram(margins):
	"""Counts of the oracle's top-2 logit margin per decade (printed with every arg-max comparison)."""
	edges = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, np.inf]
	hist, _ = np.histogram(np.asarray(margins, dtype=np.float64), bins=edges)
	return {'<{:g}'.format(hi): int(n) for hi, n in zip(edges[1:], hist)}


def check_argmax(got_argmax, oracle_logits, noise):
	"""Bit-exact category assignment wherever the oracle's decision is outside fp32 noise."""
	ol = oracle_logits.double()
	top2 = ol.topk(2, -1).values
	margin = (top2[:, 0] - top2[:, 1]).numpy()
	want = ol.argmax(-1).numpy()
	got = np.asarray(got_argmax).astype(np.int64)
	wrong = np.nonzero(got != want)[0]
	report = dict(segments=len(want), agree=int(len(want) - len(wrong)), min_margin=float(margin.min()),
				noise_floor=float(noise), margin_histogram=margin_histogram(margin),
				disagreements=[(int(i), float(margin[i])) for i in wrong[:10]])
	bad = [int(i) for i in wrong if margin[i] > noise]
	return report, bad


@pytest.mark.parametrize('name', sorted(FULL_CASES))
def test_publication_batch_matches_oracle(abcd, name):
	from test_gpu_parity import _cuda_step
	cfg = dict(FULL_CASES[name])
	cfg['lengths'] = [int(v) for v in CS.synthetic_lengths(512, 1000)]
	assert len(cfg['lengths']) == 512 and sum(cfg['lengths']) == 14755
	out, (enc_sd, smp_sd, dec_sd), seqs, eps, gumbel, speaker = _cuda_step(abcd, cfg)
	dt = torch.float64
	cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
	t0 = time.time()
	o = O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dt) for s in seqs], eps.to(dt), gumbel.to(dt),
					cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, speaker, batched=True)
	oracle_s = time.time() - t0
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
	noise = TOL * float(o['logits'].abs().max())
	am, bad = check_argmax(out['logits'].argmax(-1).cpu().numpy(), o['logits'], noise)
	print(name, 'oracle fp64 %.1f s;' % oracle_s, {k: '%.1e' % v for k, v in report.items() if not k.startswith('grad/')},
		'worst grad', worst, 'argmax', am)
	for k in ('emission', 'offset', 'kl', 'loss'):
		assert report[k] < 2e-5, (k, report[k])
	for k, v in report.items():
		if k not in ('emission', 'offset', 'loss', 'kl'):
			assert v < TOL, (k, v)
	assert not bad, ('arg-max differs where the oracle margin is above fp32 noise', bad, am)
	assert am['agree'] == am['segments'] or am['min_margin'] <= noise


def test_bulk_encode_argmax_vs_oracle(abcd):
	"""BASELINE.json configs[4]: segments of 0.05-2 s at 16 kHz (T = 1 + floor(dur*fs)//64 in [13, 501]) in length-sorted
	batches of 512 through encoder (inference variant) -> ABCD sampler logits -> softmax / arg-max, no clipping of T."""
	n_seg = int(os.environ.get('ABCD_ENCODE_PARITY_SEGMENTS', '10240'))
	F, K, bsz = 65, 128, 512
	lengths = CS.encode_lengths(n_seg, 5)
	assert lengths.max() >= 495 and lengths.min() <= 16
	cfg = dict(rnn='LSTM', F=F, K=K, f=256)
	enc, smp, _ = CS.build_modules(abcd.model, cfg)
	enc_sd, smp_sd = CS.det_state(enc, 31), CS.det_state(smp, 32)
	enc.load_state_dict(enc_sd); smp.load_state_dict(smp_sd)
	enc.cuda(); smp.cuda()
	P = {k[len('rnn.'):]: v.double() for k, v in enc_sd.items()}
	S = {k: v.double() for k, v in smp_sd.items()}
	rng = np.random.RandomState(77)
	L = abcd._lib
	all_margin, n_wrong, bad_total = [], 0, []
	worst_prob, worst_logit, t_oracle = 0.0, 0.0, 0.0
	for start in range(0, n_seg, bsz):
		ls = lengths[start:start + bsz]
		bs = O.batch_sizes_from_lengths(ls.tolist())
		x = np.clip(rng.normal(0.45, 0.25, size=(int(ls.sum()), F)), -0.945, 1.05).astype(np.float32)
		data = torch.from_numpy(x)
		with torch.no_grad():
			logits = smp(enc(torch.nn.utils.rnn.PackedSequence(data.cuda(), torch.tensor(bs, dtype=torch.int64))))
			B = logits.shape[0]
			probs = torch.empty_like(logits)
			amax = torch.empty(B, dtype=torch.int32, device='cuda')
			L.call('abcd_softmax_argmax', B, K, L.ptr(logits), L.ptr(probs), L.ptr(amax), L.stream_ptr())
			t0 = time.time()
			ol = O.sampler_logits(O.encoder_forward_batched(data.double(), bs, P, 'LSTM'), S)
			t_oracle += time.time() - t0
		op = O.softmax_rows(ol)
		worst_prob = max(worst_prob, float((probs.double().cpu() - op).abs().max() / op.max()))
		worst_logit = max(worst_logit, rel_to_max(logits, ol))
		noise = TOL * float(ol.abs().max())
		rep, bad = check_argmax(amax.cpu().numpy(), ol, noise)
		n_wrong += rep['segments'] - rep['agree']
		bad_total += [(start + i) for i in bad]
		top2 = ol.topk(2, -1).values
		all_margin.append((top2[:, 0] - top2[:, 1]).numpy())
	margins = np.concatenate(all_margin)
	print('bulk encode parity: {} segments, T in [{}, {}], frames {}; arg-max agreement {}/{}; min top-2 margin {:.3e}; '
		'margin histogram {}; worst prob diff {:.2e} of max, worst logit diff {:.2e} of max; oracle fp64 {:.0f} s'.format(
			n_seg, int(lengths.min()), int(lengths.max()), int(lengths.sum()), n_seg - n_wrong, n_seg, float(margins.min()),
			margin_histogram(margins), worst_prob, worst_logit, t_oracle))
	assert worst_logit < TOL and worst_prob < TOL
	assert not bad_total, ('arg-max differs where the oracle margin is above fp32 noise', bad_total[:10])


@pytest.mark.parametrize('B,K', [(512, 128), (7, 16), (93, 128)])
def test_sampler_stats_match_oracle_perplexities(abcd, B, K):
	"""Learner.train's log line (learning.py:171-178): cluster / batch-mean / Dirichlet-shape perplexities."""
	smp = abcd.model.ABCDSampler(32, 32, K, 16).cuda()
	logits = (O.det_tensor((B, K), 600 + B, 'normal', 2.5)).cuda()
	phi = O.det_tensor((K,), 601 + B, 'normal', 1.0)
	with torch.no_grad():
		smp.posterior_shape_logits.copy_(phi.cuda())
	got = smp.perplexities(logits).double().cpu()
	want = torch.stack(O.perplexities(logits.double().cpu(), phi.double()))
	rel = ((got - want).abs() / want).max().item()
	print('perplexities', got.tolist(), want.tolist(), 'rel %.1e' % rel)
	assert rel < 1e-5


def test_learner_20_step_drift_vs_oracle(abcd):
	"""20 consecutive Learner.train_step calls (backward kernels accumulating into the flat gradient arena, fused global-norm
	clip + SGD with momentum) against the fp64 oracle step + clip_and_sgd: the parameter deviation must stay at fp32
	round-off level over the whole trajectory (reference: learning.py:147-163)."""
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	lengths = sorted([(i * 11) % 29 + 2 for i in range(48)], reverse=True)
	cfg = dict(rnn='LSTM', F=65, K=128, f=256, lengths=lengths, pretrain=False, tau=0.9, N_data=5000, seed=9)
	learner = learning.Learner(65, 256, 256, 256, cfg['K'], cfg['f'], None, device='cuda', seed=3)
	mods = learner._modules()
	enc_sd, smp_sd, dec_sd, seqs, _, _, _ = CS.build_case(cfg, mods)
	for m, sd in zip(mods, (enc_sd, smp_sd, dec_sd)):
		m.load_state_dict({k: v.cuda() for k, v in sd.items()})
	lr, mom, clip = 0.3, 0.9, 1.0
	learner.setup_optimizer(lr, mom, gradient_clip=clip, patience=0)
	learner.feature_sampler.train()
	learner.feature_sampler.temperature = cfg['tau']
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	is_off = O.is_offset_packed(packed.batch_sizes.tolist())
	p_in = torch.nn.utils.rnn.PackedSequence(packed.data.cuda(), packed.batch_sizes)
	p_off = torch.nn.utils.rnn.PackedSequence(is_off.cuda(), packed.batch_sizes)
	dt = torch.float64
	state = [{k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in sd.items()} for sd in (enc_sd, smp_sd, dec_sd)]
	names = ('encoder', 'feature_sampler', 'decoder')
	sumT = sum(lengths)
	bufs = None
	curve = []
	for step in range(20):
		eps = O.det_tensor((sumT, 65), 7000 + step, 'normal', 1.0)
		gumbel = O.det_tensor((len(lengths), cfg['K']), 8000 + step, 'gumbel', 1.0)
		learner.feature_sampler.injected_gumbel_noise = gumbel.cuda()
		learner.decoder.injected_emission_noise = eps.cuda()
		loss = learner.train_step(p_in, p_off, None, cfg['N_data'], False)[0]
		o = O.train_step(state[0], state[1], state[2], [x.to(dt) for x in seqs], eps.to(dt), gumbel.to(dt), cfg['N_data'], False,
						cfg['tau'], cfg['rnn'], 1.0, None, batched=True)
		keys, params, grads = [], [], []
		for i, n in enumerate(names):
			for k, g in o['grads'][n].items():
				keys.append((i, k)); params.append(state[i][k]); grads.append(g)
		new, total, bufs = O.clip_and_sgd(params, grads, lr, clip, mom, bufs)
		for (i, k), p in zip(keys, new):
			state[i][k] = p.detach()
		bufs = [b.detach() for b in bufs]
		loss_rel = abs(loss.item() - o['loss'].item()) / abs(o['loss'].item())
		worst = 0.0
		for i, m in enumerate(mods):
			for k, v in m.state_dict().items():
				if v.dtype.is_floating_point and k != 'prior_concentration':
					worst = max(worst, rel_to_max(v, state[i][k]))
		curve.append((loss_rel, abs(learner.arena.norm.item() - total) / total, worst))
	print('20-step drift (loss rel, grad-norm rel, worst parameter rel-to-max):')
	for s, c in enumerate(curve):
		print('  step %2d  %.1e  %.1e  %.1e' % ((s,) + c))
	for s, (lr_, nr, pr) in enumerate(curve):
		assert lr_ < 5e-5 and nr < TOL and pr < TOL, (s, lr_, nr, pr)


@pytest.mark.parametrize('B', [64, 200])
def test_cuda_graph_replay_equals_host_launched_steps(abcd, B):
	"""Learner.use_cuda_graph: a step recorded into a CUDA graph and replayed (with the decoder weight gradients on the side
	stream and the Philox position on the device) must leave bit-identical parameters to the same steps launched from the
	host -- same kernels, same order of floating-point operations, same noise stream."""
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	S = importlib.import_module('seq2seq_abcd-vae_b200.synthetic')
	lengths = S.synthetic_lengths(B, 77)
	bs = S.batch_sizes_of(lengths)
	x = S.synthetic_frames(int(lengths.sum()), 65, 3).cuda()
	off = S.is_offset_of(lengths, bs).cuda()
	p_in = torch.nn.utils.rnn.PackedSequence(x, bs)
	p_off = torch.nn.utils.rnn.PackedSequence(off, bs)
	results = []
	for use_graph in (False, True):
		learner = learning.Learner(65, 256, 256, 256, 128, 256, None, device='cuda', seed=5)
		for i, m in enumerate(learner._modules()):
			m.load_state_dict({k: v.cuda() for k, v in S.deterministic_state(m, 900 + i).items()})
		abcd.model.manual_noise_seed(1234)
		learner.setup_optimizer(0.5, 0.9, gradient_clip=1.0, patience=0)
		learner.feature_sampler.train()
		learner.use_cuda_graph = use_graph
		losses = []
		for step in range(6):
			losses.append(learner.train_step(p_in, p_off, None, 5000, False)[0].detach().clone())
		torch.cuda.synchronize()
		results.append((torch.stack(losses).cpu(), learner.arena.flat.detach().cpu().clone(), learner.arena.momentum_buf.cpu().clone(),
						dict(abcd.model._noise_state), learner.launch_mode(), learner.kernel_launches()))
	(l0, p0, m0, n0, mode0, k0), (l1, p1, m1, n1, mode1, k1) = results
	print('graph parity B=%d: losses' % B, l0.tolist(), l1.tolist(), mode0, '|', mode1, 'launches', k0, k1)
	assert 'graph' in mode1 and 'graph' not in mode0
	assert n0['offset'] == n1['offset'] > 0
	assert len(set(l0.tolist())) == 6              # fresh noise every step (also under replay)
	assert torch.equal(l0, l1)
	assert torch.equal(p0, p1) and torch.equal(m0, m1)
