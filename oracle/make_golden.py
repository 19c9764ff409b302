# coding: utf-8
"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference, which
exists only in the build container) on deterministic weights / inputs / noise, check that oracle/abcd_oracle.py
reproduces them, and store compact vectors for the GPU box (where the reference is absent).

    python oracle/make_golden.py            # writes tests/golden/*.npz, prints oracle-vs-reference deviations

Compat shims applied to the reference at import time (never edits /root/reference):
  * STFT: torch>=2 needs return_complex=True  -> we call torch.stft ourselves exactly as data_utils.py:132-139 does;
  * Embedding(sparse=True) -> dense grads (same math) so that gradients can be compared.
Noise hooks: decoder.emission_sampler._sampler is replaced by a function consuming recorded N(0,1) draws in packed
order; torch.nn.functional.gumbel_softmax is wrapped to consume recorded Gumbel draws.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import abcd_oracle as O  # noqa: E402
from oracle import cases as CS  # noqa: E402
from oracle.cases import TRAIN_CASES, det_state, sample_vec  # noqa: E402

REF = os.environ.get('ABCD_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def load_reference_model(package='ABCD-VAE'):
	path = os.path.join(REF, package, 'modules', 'model.py')
	spec = importlib.util.spec_from_file_location('ref_model_' + package.replace('-', '_'), path)
	mod = importlib.util.module_from_spec(spec)
	spec.loader.exec_module(mod)
	return mod


def run_reference_step(ref, cfg, enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker):
	gaussian = cfg.get('latent') == 'gaussian'
	enc = ref.RNN_Variational_Encoder(cfg['F'], 256, rnn_type=cfg['rnn'])
	smp = ref.Sampler(enc.hidden_size_total, 256, cfg['f']) if gaussian else ref.ABCDSampler(enc.hidden_size_total, 256, cfg['K'], cfg['f'])
	dec = ref.RNN_Variational_Decoder(cfg['F'], 256, 256, cfg['f'], rnn_type=cfg['rnn'],
										num_speakers=cfg.get('num_speakers'), speaker_embed_dim=cfg.get('E'))
	if dec.embed_speaker is not None:
		dec.embed_speaker = torch.nn.Embedding(cfg['num_speakers'], cfg['E'], sparse=False)
	enc.load_state_dict(enc_sd); smp.load_state_dict(smp_sd); dec.load_state_dict(dec_sd)
	smp.temperature = cfg['tau']
	if gaussian:
		smp._sampler = lambda mean, log_variance: mean + (0.5 * log_variance).exp() * gumbel
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	is_off = torch.nn.utils.rnn.pack_sequence([torch.tensor([0.0] * (s.size(0) - 1) + [1.0]) for s in seqs])
	cursor = {'i': 0}

	def recorded_sampler(mean, log_variance):
		n = mean.shape[0]
		e = eps[cursor['i']:cursor['i'] + n]
		cursor['i'] += n
		return mean + (0.5 * log_variance).exp() * e
	dec.emission_sampler._sampler = recorded_sampler
	orig = torch.nn.functional.gumbel_softmax

	def recorded_gumbel(logits, tau=1, hard=False, eps=1e-10, dim=-1):
		return ((logits + gumbel) / tau).softmax(dim)
	torch.nn.functional.gumbel_softmax = recorded_gumbel
	try:
		last_hidden = enc(packed)
		if gaussian:   # plain/learning.py:147-149
			params = smp(last_hidden)
			feats = smp.sample(params)
			kl = smp.kl_divergence(params)
			logits = params[0]
		else:
			logits = smp(last_hidden)
			feats = smp.sample(logits, no_sample=cfg['pretrain'])
			kl = smp.kl_divergence(logits, cfg['N_data'])
		em, off, x_all, (mean, lv), off_logit = dec(feats, batch_sizes=packed.batch_sizes, speaker=speaker,
													ground_truth_out=packed.data, ground_truth_offset=is_off.data)
		loss = (em + off + kl) / packed.batch_sizes[0]
		loss.backward()
	finally:
		torch.nn.functional.gumbel_softmax = orig
	grads = {'encoder': {k: p.grad for k, p in enc.named_parameters()},
			'feature_sampler': {k: p.grad for k, p in smp.named_parameters()},
			'decoder': {k: p.grad for k, p in dec.named_parameters()}}
	return dict(emission=em.detach(), offset=off.detach(), kl=kl.detach(), loss=loss.detach(), logits=logits.detach(),
				last_hidden=last_hidden.detach(), features=feats.detach(), flatten_out=x_all.detach(), mean=mean.detach(),
				logvar=lv.detach(), offset_logit=off_logit.detach(), grads=grads), (enc, smp, dec)


def build_case(cfg, ref):
	return CS.build_case(cfg, CS.build_modules(ref, cfg))


def max_rel(a, b):
	a, b = a.double(), b.double()
	return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def gen_train_cases(ref, cases=None):
	for name, cfg in (cases if cases is not None else TRAIN_CASES).items():
		enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = build_case(cfg, ref)
		r, _ = run_reference_step(ref, cfg, enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker)
		dev = {}
		for dt in (torch.float32, torch.float64):
			cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
			o = O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dt) for s in seqs], eps.to(dt), gumbel.to(dt),
							cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, speaker, latent=cfg.get('latent', 'abcd'))
			tag = 'f32' if dt == torch.float32 else 'f64'
			for k in ('emission', 'offset', 'kl', 'loss', 'logits', 'last_hidden', 'flatten_out', 'mean', 'logvar', 'offset_logit'):
				dev['{}_{}'.format(k, tag)] = max_rel(o[k], r[k])
			worst = 0.0
			for m in r['grads']:
				for k, g in r['grads'][m].items():
					worst = max(worst, max_rel(o['grads'][m][k], g))
			dev['grads_' + tag] = worst
		print(name, {k: '%.2e' % v for k, v in dev.items()})
		assert dev['grads_f64'] < 2e-4 and dev['loss_f64'] < 1e-5, 'oracle does not reproduce the reference'
		out = {'cfg_keys': np.array(sorted(cfg.keys())), 'lengths': np.array(cfg['lengths'])}
		for k in ('emission', 'offset', 'kl', 'loss'):
			out[k] = np.float64(r[k].item())
		out['logits'] = r['logits'].numpy()
		out['last_hidden_sample'] = sample_vec(r['last_hidden'])
		out['flatten_out_sample'] = sample_vec(r['flatten_out'])
		out['mean_sample'] = sample_vec(r['mean'])
		out['offset_logit'] = r['offset_logit'].numpy()
		for m in r['grads']:
			for k, g in r['grads'][m].items():
				out['gradnorm/{}/{}'.format(m, k)] = np.float64(g.double().norm().item())
				out['gradsample/{}/{}'.format(m, k)] = sample_vec(g)
		for k, v in dev.items():
			out['oracle_dev/' + k] = np.float64(v)
		np.savez_compressed(os.path.join(GOLD, name + '.npz'), **out)


def gen_stft(ref_root):
	"""Reference front end on two toy segments: exactly data_utils.STFT.__call__ + log_and_normalize, with the
	return_complex shim."""
	import scipy.io.wavfile as spw
	fs, wav = spw.read(os.path.join(ref_root, 'toy_data', '20170806-080002_89.2-94.22.1ch.wav'))
	frame, hop = int(np.floor(0.008 * fs)), int(np.floor(0.004 * fs))
	segs = [(1.903, 2.149), (2.559, 2.912)]
	out = {'fs': np.int64(fs), 'frame': np.int64(frame), 'hop': np.int64(hop)}
	window = torch.hann_window(frame)
	for i, (on, off) in enumerate(segs):
		a, b, _ = O.discrete_bounds(on, off, fs)
		x = wav[a:b]
		xt = torch.from_numpy(x.astype(np.float32))
		tr = torch.view_as_real(xt.stft(frame, hop_length=hop, window=window, center=True, return_complex=True))
		amp = tr.pow(2).sum(-1).sqrt().transpose(0, 1).contiguous()
		logn = (amp + 2 ** (-15)).log() / 11.0
		oa = O.stft_amplitude(x, frame, hop)
		print('stft seg', i, 'frames', amp.shape, 'oracle-vs-ref amp rel dev %.2e' % (np.abs(oa - amp.numpy()).max() / amp.numpy().max()))
		out['wav_%d' % i] = x.astype(np.int16)
		out['amp_%d' % i] = amp.numpy()
		out['logn_%d' % i] = logn.numpy()
		out['bounds_%d' % i] = np.array([a, b])
	np.savez_compressed(os.path.join(GOLD, 'stft_toy.npz'), **out)


def gen_plain(ref_plain):
	"""plain/ Gaussian latent: Sampler(1024,256,f) sample + KL and their gradients."""
	f = 16
	smp = ref_plain.Sampler(1024, 256, f)
	sd = det_state(smp, 77)
	smp.load_state_dict(sd)
	x = O.det_tensor((5, 1024), 771, 'uniform', 1.0).requires_grad_(True)
	eps = O.det_tensor((5, f), 772, 'normal', 1.0)
	smp._sampler = lambda mean, lv: mean + (0.5 * lv).exp() * eps
	params = smp(x)
	z = smp.sample(params)
	kl = smp.kl_divergence(params)
	w = O.det_tensor((5, f), 773, 'uniform', 1.0)
	((z * w).sum() + kl).backward()
	out = {'z': z.detach().numpy(), 'kl': np.float64(kl.item()), 'dx': x.grad.numpy()}
	for k, p in smp.named_parameters():
		out['grad/' + k] = p.grad.numpy() if p.grad.numel() < 5000 else sample_vec(p.grad)
	P = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
	xo = x.detach().clone().requires_grad_(True)
	zo, klo, _, _ = O.gauss_latent(xo, P, eps)
	((zo * w).sum() + klo).backward()
	print('plain: oracle-vs-ref z %.2e kl %.2e dx %.2e' % (max_rel(zo, z), max_rel(klo, kl), max_rel(xo.grad, x.grad)))
	np.savez_compressed(os.path.join(GOLD, 'plain_latent.npz'), **out)


def gen_optim():
	"""clip_grad_norm_ + SGD on deterministic tensors via the real torch implementations."""
	shapes = [(300, 7), (1024,), (33, 65), (5,)]
	out = {}
	for mom, tag in ((0.0, 'plain'), (0.9, 'momentum')):
		params = [torch.nn.Parameter(O.det_tensor(s, 900 + i, 'uniform', 0.5)) for i, s in enumerate(shapes)]
		opt = torch.optim.SGD(params, lr=0.7, momentum=mom)
		for step in range(2):
			for i, p in enumerate(params):
				p.grad = O.det_tensor(tuple(p.shape), 950 + 10 * step + i, 'normal', 0.3)
			norm = torch.nn.utils.clip_grad_norm_(params, 1.0)
			opt.step()
			out['{}_norm_{}'.format(tag, step)] = np.float64(norm.item())
		out[tag + '_params'] = torch.cat([p.detach().reshape(-1) for p in params]).numpy()
	np.savez_compressed(os.path.join(GOLD, 'optim.npz'), **out)


if __name__ == '__main__':
	assert os.path.isdir(REF), 'the reference checkout is needed to (re)generate golden vectors'
	os.makedirs(GOLD, exist_ok=True)
	torch.set_num_threads(4)
	torch.manual_seed(0)
	ref = load_reference_model('ABCD-VAE')
	gen_train_cases(ref)
	gen_stft(REF)
	ref_plain = load_reference_model('plain')
	gen_plain(ref_plain)
	gen_train_cases(ref_plain, CS.PLAIN_CASES)
	gen_optim()
	print('golden vectors written to', GOLD)
