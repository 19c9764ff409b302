# coding: utf-8
"""CPU oracle for the ABCD-VAE hot path -- TEST INFRASTRUCTURE ONLY.

A from-scratch restatement of the reference algorithm (tkc-morita/seq2seq_abcd-vae) in explicit per-sequence /
per-step tensor arithmetic on the CPU, in float32 or float64.  It is what the CUDA path is checked against on the
GPU box, where /root/reference does not exist.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors, so parity is pinned by running the
reference's own modules here (oracle/make_golden.py imports /root/reference/ABCD-VAE/modules/model.py, feeds both
the same weights, inputs and noise, asserts agreement and commits the vectors under tests/golden/).

Every function cites the reference lines it restates (paths relative to /root/reference/ABCD-VAE/).
Only elementary tensor ops are used (matmul, sigmoid, tanh, exp, log, lgamma, digamma): no nn.LSTM/GRU, no
nn.Linear, no torch.stft, no pack_sequence, no clip_grad_norm_, no optim.SGD.  Gradients come from autograd over
these elementary ops.
"""
import math

import numpy as np
import torch

LOG_2PI = math.log(2.0 * math.pi)


# ---------------------------------------------------------------------------------------------------
# deterministic, platform-independent pseudo-random tensors (integer hash -> uniform -> Box-Muller)
# ---------------------------------------------------------------------------------------------------
def _splitmix64(x):
	x = (x + np.uint64(0x9E3779B97F4A7C15))
	z = x
	z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
	z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
	return z ^ (z >> np.uint64(31))


def det_uniform(shape, seed, lo=0.0, hi=1.0):
	"""float64 uniforms in (lo, hi), a pure function of (shape, seed)."""
	n = int(np.prod(shape)) if len(shape) else 1
	with np.errstate(over='ignore'):
		idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
		bits = _splitmix64(_splitmix64(idx))
	u = ((bits >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
	return (lo + (hi - lo) * u).reshape(shape)


def det_normal(shape, seed):
	u1 = det_uniform(shape, seed * 2 + 1)
	u2 = det_uniform(shape, seed * 2 + 2)
	return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def det_gumbel(shape, seed):
	return -np.log(-np.log(det_uniform(shape, seed)))


def det_tensor(shape, seed, kind='uniform', scale=1.0, dtype=torch.float32):
	if kind == 'uniform':
		a = det_uniform(shape, seed, -scale, scale)
	elif kind == 'normal':
		a = det_normal(shape, seed) * scale
	else:
		a = det_gumbel(shape, seed) * scale
	return torch.from_numpy(np.asarray(a, dtype=np.float64)).to(dtype)


# ---------------------------------------------------------------------------------------------------
# packing (modules/data_utils.py:81-82,165-182; torch.nn.utils.rnn.pack_sequence semantics)
# ---------------------------------------------------------------------------------------------------
def sort_by_length_desc(lengths):
	"""Stable descending order of sample lengths (data_utils.py:81-82 sorts descending; ties are
	implementation-defined there, stable here)."""
	return np.argsort(-np.asarray(lengths), kind='stable')


def batch_sizes_from_lengths(lengths):
	"""batch_sizes[t] = number of sequences longer than t (lengths descending)."""
	lengths = [int(v) for v in lengths]
	assert all(a >= b for a, b in zip(lengths, lengths[1:])), 'lengths must be descending'
	return [sum(1 for v in lengths if v > t) for t in range(lengths[0])]


def pack(seqs):
	"""Time-major interleave of already length-sorted [T_i, ...] tensors -> (data, batch_sizes, offsets)."""
	lengths = [int(s.shape[0]) for s in seqs]
	bs = batch_sizes_from_lengths(lengths)
	rows = []
	for t, n in enumerate(bs):
		for b in range(n):
			rows.append(seqs[b][t])
	offsets = [0]
	for n in bs:
		offsets.append(offsets[-1] + n)
	return torch.stack(rows), bs, offsets


def unpack(data, batch_sizes):
	"""Inverse of pack: list of per-sequence [T_i, ...] tensors."""
	B = batch_sizes[0]
	out = [[] for _ in range(B)]
	off = 0
	for n in batch_sizes:
		for b in range(n):
			out[b].append(data[off + b])
		off += n
	return [torch.stack(o) for o in out]


def is_offset_packed(batch_sizes):
	"""Packed one-hot of each segment's last frame (data_utils.py:178,180)."""
	B = batch_sizes[0]
	lengths = [sum(1 for n in batch_sizes if n > b) for b in range(B)]
	seqs = [torch.tensor([0.0] * (l - 1) + [1.0]) for l in lengths]
	return pack(seqs)[0]


# ---------------------------------------------------------------------------------------------------
# front end (modules/data_utils.py:69-79, 88-103, 131-139; learning.py:433-434, 466)
# ---------------------------------------------------------------------------------------------------
def discrete_bounds(onset_sec, offset_sec, fs):
	"""round-half-even(seconds * fs) (pandas Series.round), data_utils.py:73-79."""
	on = int(np.round(np.float64(onset_sec) * fs))
	off = int(np.round(np.float64(offset_sec) * fs))
	return on, off, off - on


def hann_periodic(n, dtype=np.float64):
	return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(dtype)


def stft_amplitude(x, n_fft, hop, window=None, center=True):
	"""|STFT| as [T, F] in float64 by explicit DFT: reflect padding of n_fft//2 (torch.stft center=True),
	frames of n_fft samples every hop samples, window, onesided DFT, magnitude (data_utils.py:132-139)."""
	x = np.asarray(x, dtype=np.float64)
	if window is None:
		window = hann_periodic(n_fft)
	window = np.asarray(window, dtype=np.float64)
	if center:
		p = n_fft // 2
		assert x.shape[0] > p, 'reflect padding needs more than n_fft//2 samples'
		x = np.concatenate([x[1:p + 1][::-1], x, x[-p - 1:-1][::-1]])
	T = 1 + (x.shape[0] - n_fft) // hop
	frames = np.stack([x[t * hop:t * hop + n_fft] for t in range(T)]) * window[None, :]
	k = np.arange(n_fft // 2 + 1)
	n = np.arange(n_fft)
	ang = -2.0 * np.pi * np.outer(n, k) / n_fft
	re = frames @ np.cos(ang)
	im = frames @ np.sin(ang)
	return np.sqrt(re * re + im * im)


def log_normalize(amp, eps=2.0 ** -15, normalizer=1.0):
	"""log(x + eps) / N (learning.py:466)."""
	return np.log(amp + eps) / normalizer


# ---------------------------------------------------------------------------------------------------
# recurrent cells (torch.nn.LSTM / GRU / LSTMCell / GRUCell arithmetic; model.py:53, 295)
# ---------------------------------------------------------------------------------------------------
def lstm_step(x_proj, h, c, w_hh, b_hh):
	"""gates = x_proj + h W_hh^T + b_hh, order i,f,g,o."""
	H = h.shape[-1]
	g = x_proj + h @ w_hh.t() + b_hh
	i = torch.sigmoid(g[..., 0:H]); f = torch.sigmoid(g[..., H:2 * H])
	gg = torch.tanh(g[..., 2 * H:3 * H]); o = torch.sigmoid(g[..., 3 * H:4 * H])
	c2 = f * c + i * gg
	return o * torch.tanh(c2), c2


def gru_step(x_proj, h, w_hh, b_hh):
	"""r,z,n order; n = tanh(x_n + r * (W_hn h + b_hn)); h' = (1-z) n + z h."""
	H = h.shape[-1]
	hp = h @ w_hh.t() + b_hh
	r = torch.sigmoid(x_proj[..., 0:H] + hp[..., 0:H])
	z = torch.sigmoid(x_proj[..., H:2 * H] + hp[..., H:2 * H])
	n = torch.tanh(x_proj[..., 2 * H:3 * H] + r * hp[..., 2 * H:3 * H])
	return (1.0 - z) * n + z * h


def encoder_forward(seqs, P, rnn_type='LSTM', bidirectional=True):
	"""RNN_Variational_Encoder.forward (model.py:60-66): run each (length-sorted) sequence through a 1-layer
	(bi)directional RNN from a zero state -- the reverse direction starts at the sequence's own last frame -- and
	return the final states per sequence as [h_f | c_f | h_b | c_b] (LSTM) or [h_f | h_b] (GRU).
	P: dict with the torch.nn.LSTM/GRU key names (weight_ih_l0, ..., bias_hh_l0_reverse)."""
	outs = []
	H = P['weight_hh_l0'].shape[1]
	for x in seqs:
		per_dir = []
		for suffix, order in (('', range(x.shape[0])), ('_reverse', range(x.shape[0] - 1, -1, -1))):
			if suffix and not bidirectional:
				continue
			w_ih, w_hh = P['weight_ih_l0' + suffix], P['weight_hh_l0' + suffix]
			b_ih, b_hh = P['bias_ih_l0' + suffix], P['bias_hh_l0' + suffix]
			h = x.new_zeros(H); c = x.new_zeros(H)
			for t in order:
				xp = w_ih @ x[t] + b_ih
				if rnn_type == 'LSTM':
					h, c = lstm_step(xp, h, c, w_hh, b_hh)
				else:
					h = gru_step(xp, h, w_hh, b_hh)
			per_dir.append(torch.cat([h, c]) if rnn_type == 'LSTM' else h)
		outs.append(torch.cat(per_dir))
	return torch.stack(outs)


def encoder_forward_batched(data, batch_sizes, P, rnn_type='LSTM', bidirectional=True):
	"""Same arithmetic as encoder_forward, organised the way the reference's nn.LSTM/GRU runs it on a
	PackedSequence: one batched cell step per time index over the active prefix of the batch (used for the CPU
	baseline timing; checked against the per-sequence version in tests/test_cpu_oracle.py)."""
	H = P['weight_hh_l0'].shape[1]
	B, T = batch_sizes[0], len(batch_sizes)
	offs = [0]
	for n in batch_sizes:
		offs.append(offs[-1] + n)
	per_dir = []
	for suffix in ('', '_reverse'):
		if suffix and not bidirectional:
			continue
		w_ih, w_hh = P['weight_ih_l0' + suffix], P['weight_hh_l0' + suffix]
		b_ih, b_hh = P['bias_ih_l0' + suffix], P['bias_hh_l0' + suffix]
		xp_all = data @ w_ih.t() + b_ih
		h = data.new_zeros(0, H); c = data.new_zeros(0, H)
		finals_h, finals_c = [], []
		order = range(T) if not suffix else range(T - 1, -1, -1)
		for t in order:
			bs = batch_sizes[t]
			if not suffix:
				if t == 0:
					h = data.new_zeros(bs, H); c = data.new_zeros(bs, H)
				elif bs < h.shape[0]:   # rows bs.. finished at t-1
					finals_h.append(h[bs:]); finals_c.append(c[bs:])
					h = h[:bs]; c = c[:bs]
			else:
				if bs > h.shape[0]:     # rows start at their own last frame with a zero state
					pad = data.new_zeros(bs - h.shape[0], H)
					h = torch.cat([h, pad]); c = torch.cat([c, pad])
			xp = xp_all[offs[t]:offs[t] + bs]
			if rnn_type == 'LSTM':
				h, c = lstm_step(xp, h, c, w_hh, b_hh)
			else:
				h = gru_step(xp, h, w_hh, b_hh)
		if not suffix:
			finals_h.append(h); finals_c.append(c)
			hf = torch.cat(finals_h[::-1]); cf = torch.cat(finals_c[::-1])
		else:
			hf, cf = h, c
		per_dir.append(torch.cat([hf, cf], dim=-1) if rnn_type == 'LSTM' else hf)
	return torch.cat(per_dir, dim=-1)


def mlp(x, P, prefix):
	"""Linear -> Tanh -> Linear (model.py:326-334)."""
	h = torch.tanh(x @ P[prefix + 'whole_network.0.weight'].t() + P[prefix + 'whole_network.0.bias'])
	return h @ P[prefix + 'whole_network.2.weight'].t() + P[prefix + 'whole_network.2.bias']


# ---------------------------------------------------------------------------------------------------
# ABCD sampler (model.py:581-639)
# ---------------------------------------------------------------------------------------------------
def sampler_logits(last_hidden, P):
	"""model.py:588-589."""
	x = mlp(last_hidden, P, 'to_code_like.')
	return (x @ P['codebook']) / math.sqrt(x.shape[-1])


def softmax_rows(v):
	m = v.max(dim=-1, keepdim=True).values
	e = torch.exp(v - m)
	return e / e.sum(dim=-1, keepdim=True)


def log_softmax_rows(v):
	m = v.max(dim=-1, keepdim=True).values
	return v - m - torch.log(torch.exp(v - m).sum(dim=-1, keepdim=True))


def sampler_sample(logits, P, no_sample, tau=1.0, gumbel=None):
	"""model.py:601-605. gumbel = -log(Exp(1)) draws (torch.nn.functional.gumbel_softmax, soft)."""
	if no_sample:
		y = softmax_rows(logits)
	else:
		y = softmax_rows((logits + gumbel) / tau)
	return y @ P['codebook'].t(), y


def sampler_kl(logits, P, entire_data_size, prior_concentration=1.0):
	"""model.py:615-639."""
	phi = P['posterior_shape_logits']
	K = phi.shape[0]
	s = softmax_rows(phi[None, :])[0]
	alpha = s * entire_data_size + prior_concentration
	A = alpha.sum()
	E = torch.digamma(alpha) - torch.digamma(A)
	c = torch.as_tensor(prior_concentration, dtype=logits.dtype)
	q_pi = torch.lgamma(A) - torch.lgamma(alpha).sum() + ((alpha - 1.0) * E).sum()
	p_pi = torch.lgamma(c * K) - torch.lgamma(c) * K + ((c - 1.0) * E).sum()
	q = softmax_rows(logits)
	lq = log_softmax_rows(logits)
	B = logits.shape[0]
	return (q_pi - p_pi) * (B / entire_data_size) + (q * lq).sum() - (q * E[None, :]).sum()


def perplexities(logits, phi):
	"""learning.py:171-178."""
	q = softmax_rows(logits)
	clus = torch.exp((-(q * torch.log(q)).sum(-1)).mean())
	m = q.mean(0); m = m / m.sum()
	bat = torch.exp(-(m * torch.log(m)).sum())
	s = softmax_rows(phi[None, :])[0]
	shp = torch.exp(-(s * torch.log(s)).sum())
	return clus, bat, shp


# ---------------------------------------------------------------------------------------------------
# decoder (model.py:147-196, 297-300, 683-697, 17-37)
# ---------------------------------------------------------------------------------------------------
def decoder_forward(features, batch_sizes, P, eps, gt=None, is_offset=None, rnn_type='LSTM', speaker=None):
	"""_forward_unidirectional with recorded emission noise `eps` ([sumT, F], packed order).
	P uses the decoder's state_dict key names.  Returns (emission_loss, offset_loss, flatten_out, mean, logvar,
	offset_logits)."""
	if 'embed_speaker.weight' in P and speaker is not None:
		features = torch.cat([features, P['embed_speaker.weight'][speaker]], dim=-1)   # model.py:155-157
	hid0 = features @ P['feature2hidden.weight'].t() + P['feature2hidden.bias']        # model.py:169
	H = P['rnn_cell.cell.weight_hh'].shape[1]
	if rnn_type == 'LSTM':
		h = hid0[:, 0::2]; c = hid0[:, 1::2]    # view(-1,H,2)[...,0] / [...,1]  (model.py:100, 262-266)
	else:
		h = hid0; c = None
	w_ih, w_hh = P['rnn_cell.cell.weight_ih'], P['rnn_cell.cell.weight_hh']
	b_ih, b_hh = P['rnn_cell.cell.bias_ih'], P['rnn_cell.cell.bias_hh']
	F = w_ih.shape[1]
	x = features.new_zeros(batch_sizes[0], F)       # model.py:174
	hs, means, lvs, xs = [], [], [], []
	off = 0
	for bs in batch_sizes:                           # model.py:175-183
		xp = x[:bs] @ w_ih.t() + b_ih
		if rnn_type == 'LSTM':
			h, c = lstm_step(xp, h[:bs], c[:bs], w_hh, b_hh)
		else:
			h = gru_step(xp, h[:bs], w_hh, b_hh)
		mean = mlp(h, P, 'emission_sampler.to_parameters.mlps.0.')
		lv = mlp(h, P, 'emission_sampler.to_parameters.mlps.1.')
		x = mean + torch.exp(0.5 * lv) * eps[off:off + bs]     # model.py:19
		hs.append(h); means.append(mean); lvs.append(lv); xs.append(x)
		off += bs
	h_all, mean_all, lv_all, x_all = torch.cat(hs), torch.cat(means), torch.cat(lvs), torch.cat(xs)
	emission_loss = None
	if gt is not None:                               # model.py:30-37,190
		d = gt - mean_all
		emission_loss = 0.5 * (LOG_2PI + lv_all + d * torch.exp(-lv_all) * d).sum()
	logit = mlp(h_all, P, 'offset_predictor.')[:, 0]  # model.py:191
	offset_loss = None
	if is_offset is not None:                        # BCEWithLogitsLoss(reduction='sum'), model.py:122,195
		offset_loss = (torch.clamp(logit, min=0) - logit * is_offset + torch.log1p(torch.exp(-logit.abs()))).sum()
	return emission_loss, offset_loss, x_all, mean_all, lv_all, logit


# ---------------------------------------------------------------------------------------------------
# plain/ Gaussian latent (plain/modules/model.py:17-28, 538-567)
# ---------------------------------------------------------------------------------------------------
def gauss_latent(last_hidden, P, eps):
	mean = mlp(last_hidden, P, 'to_parameters.mlps.0.')
	lv = mlp(last_hidden, P, 'to_parameters.mlps.1.')
	z = mean + torch.exp(0.5 * lv) * eps
	kl = -0.5 * (1 + lv - mean * mean - torch.exp(lv)).sum()
	return z, kl, mean, lv


# ---------------------------------------------------------------------------------------------------
# optimiser (learning.py:161-163, 256): clip_grad_norm_ then SGD(momentum, no dampening)
# ---------------------------------------------------------------------------------------------------
def clip_and_sgd(params, grads, lr, max_norm, momentum=0.0, bufs=None):
	total = math.sqrt(sum(float((g.double() ** 2).sum()) for g in grads))
	coef = min(1.0, max_norm / (total + 1e-6))
	new_params, new_bufs = [], []
	for i, (p, g) in enumerate(zip(params, grads)):
		g = g * coef
		if momentum != 0.0:
			b = g if bufs is None or bufs[i] is None else momentum * bufs[i] + g
			new_bufs.append(b)
			g = b
		new_params.append(p - lr * g)
	return new_params, total, new_bufs


# ---------------------------------------------------------------------------------------------------
# one full training step (learning.py:147-163)
# ---------------------------------------------------------------------------------------------------
def train_step(enc_P, smp_P, dec_P, seqs, eps, gumbel, entire_data_size, is_pretraining, tau=1.0, rnn_type='LSTM',
				prior_concentration=1.0, speaker=None, global_batch=None, batched=False, latent='abcd'):
	"""Forward + backward of one batch; returns dict(losses, logits, grads per module).
	latent='gaussian' is the plain/ variant (plain/learning.py:147-149): smp_P holds the two MLPs of plain Sampler and
	`gumbel` carries the N(0,1) draws [B, f] of the latent sample."""
	leaves = {}
	for name, P in (('encoder', enc_P), ('feature_sampler', smp_P), ('decoder', dec_P)):
		leaves[name] = {k: v.clone().requires_grad_(v.dtype.is_floating_point and k != 'prior_concentration') for k, v in P.items()}
	data, bs, _ = pack(seqs)
	enc_keys = {k[len('rnn.'):]: v for k, v in leaves['encoder'].items()}
	if batched:
		last_hidden = encoder_forward_batched(data, bs, enc_keys, rnn_type)
	else:
		last_hidden = encoder_forward(seqs, enc_keys, rnn_type)
	if latent == 'gaussian':
		feats, kl, z_mean, _ = gauss_latent(last_hidden, leaves['feature_sampler'], gumbel)
		logits = z_mean
	else:
		logits = sampler_logits(last_hidden, leaves['feature_sampler'])
		feats, y = sampler_sample(logits, leaves['feature_sampler'], is_pretraining, tau, gumbel)
		kl = sampler_kl(logits, leaves['feature_sampler'], entire_data_size, prior_concentration)
	is_off = is_offset_packed(bs).to(data.dtype)
	em, off, x_all, mean, lv, off_logit = decoder_forward(feats, bs, leaves['decoder'], eps, data, is_off, rnn_type, speaker)
	B = global_batch if global_batch is not None else bs[0]
	loss = (em + off + kl) / B
	loss.backward()
	grads = {m: {k: (v.grad if v.grad is not None else None) for k, v in d.items() if v.requires_grad} for m, d in leaves.items()}
	return dict(emission=em.detach(), offset=off.detach(), kl=kl.detach(), loss=loss.detach(), logits=logits.detach(),
				last_hidden=last_hidden.detach(), features=feats.detach(), flatten_out=x_all.detach(), mean=mean.detach(),
				logvar=lv.detach(), offset_logit=off_logit.detach(), grads=grads, batch_sizes=bs)
