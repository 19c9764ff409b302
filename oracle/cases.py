# coding: utf-8
"""Deterministic parity cases shared by oracle/make_golden.py (which runs the reference on them) and tests/
(which run the oracle and the CUDA path on them).  TEST INFRASTRUCTURE ONLY -- pure functions of the case config,
so the GPU box rebuilds bit-identical weights / inputs / noise without the reference or any stored tensor."""
import numpy as np
import torch

from . import abcd_oracle as O

TRAIN_CASES = {
	'train_lstm_pretrain': dict(rnn='LSTM', F=65, K=128, f=256, lengths=[11, 9, 9, 6, 3, 1], pretrain=True, tau=1.0, N_data=6, seed=1),
	'train_lstm_gumbel': dict(rnn='LSTM', F=65, K=128, f=256, lengths=[13, 12, 7, 7, 7, 2, 2], pretrain=False, tau=0.8, N_data=100000, seed=2),
	'train_gru_speaker': dict(rnn='GRU', F=65, K=128, f=256, lengths=[10, 8, 5, 5, 4], pretrain=False, tau=1.0, N_data=40, seed=3,
								num_speakers=3, E=256),
	'train_lstm_f129': dict(rnn='LSTM', F=129, K=128, f=256, lengths=[9, 6, 6, 2], pretrain=True, tau=1.0, N_data=4, seed=4),
}
# Additional GPU-vs-oracle cases (no golden file: the oracle itself is pinned to the reference by the cases above):
# several row tiles, ragged tiles, narrow / odd feature sizes.
EXTRA_CASES = {
	'train_lstm_3tiles': dict(rnn='LSTM', F=65, K=128, f=256, lengths=sorted([(i * 7) % 23 + 1 for i in range(170)], reverse=True),
								pretrain=False, tau=0.9, N_data=5000, seed=5),
	'train_gru_f33_2tiles': dict(rnn='GRU', F=33, K=64, f=128, lengths=sorted([(i * 5) % 17 + 1 for i in range(85)], reverse=True),
								pretrain=True, tau=1.0, N_data=85, seed=6),
	'train_lstm_f5': dict(rnn='LSTM', F=5, K=16, f=32, lengths=[7, 5, 5, 3, 1], pretrain=False, tau=1.0, N_data=50, seed=7),
	'train_gru_f129_speaker': dict(rnn='GRU', F=129, K=32, f=64, lengths=[12, 12, 9, 4, 4, 4, 2, 1, 1], pretrain=False, tau=0.7, N_data=9, seed=8,
								num_speakers=4, E=256),
	# 88 < F <= 128: the output size is too wide for the two-barrier decoder kernels, the library falls back to the barrier-exchange
	# ones (forward: CTA-pair tcgen05 kernel at >= 48 rows per tile, FFMA below; backward: dec_bwd_ts_kernel with 32-wide staging)
	'train_lstm_f97_7tiles': dict(rnn='LSTM', F=97, K=32, f=64, lengths=sorted([(i * 7) % 11 + 1 for i in range(343)], reverse=True),
								pretrain=False, tau=0.9, N_data=2000, seed=9),
	'train_gru_f97': dict(rnn='GRU', F=97, K=32, f=64, lengths=[9, 9, 7, 4, 4, 2, 1], pretrain=True, tau=1.0, N_data=7, seed=10),
}
# plain/ variant (Gaussian latent, BASELINE.json configs[1]); `f` is the latent size (plain/learning.py -f, default 16)
PLAIN_CASES = {
	'train_plain_lstm': dict(rnn='LSTM', F=65, f=16, lengths=[12, 9, 9, 5, 2, 1], seed=11, latent='gaussian', pretrain=False, tau=1.0, N_data=6, K=16),
	'train_plain_gru_speaker': dict(rnn='GRU', F=65, f=16, lengths=[10, 7, 7, 3], seed=12, latent='gaussian', pretrain=False, tau=1.0, N_data=4, K=16,
									num_speakers=3, E=256),
}
SAMPLE_STRIDE = 101


def det_state(module, seed):
	"""Deterministic weights with PyTorch-like scales keyed by state_dict name; depends only on the (sorted) key
	names and shapes, which are identical for the reference's modules and the B200 package's."""
	sd = {}
	for i, (k, v) in enumerate(sorted(module.state_dict().items())):
		if k == 'prior_concentration':
			sd[k] = v.clone()
			continue
		if k in ('codebook', 'posterior_shape_logits') or k.startswith('embed_speaker'):
			sd[k] = O.det_tensor(tuple(v.shape), seed * 1000 + i, 'normal', 1.0)
		else:
			fan = v.shape[-1] if v.dim() > 1 else v.shape[0]
			sd[k] = O.det_tensor(tuple(v.shape), seed * 1000 + i, 'uniform', 1.0 / np.sqrt(max(fan, 1)))
	return sd


def make_inputs(lengths, F, seed):
	"""Synthetic log-STFT-like frames: clip(N(0.45, 0.25^2), -0.945, 1.05)."""
	seqs = []
	for i, l in enumerate(lengths):
		x = O.det_tensor((l, F), seed * 100 + i, 'normal', 0.25) + 0.45
		seqs.append(x.clamp(-0.945, 1.05))
	return seqs


def build_modules(model_module, cfg):
	"""Instantiate encoder / ABCD sampler / decoder from `model_module` (the reference's model.py or the B200
	package's) for a case."""
	enc = model_module.RNN_Variational_Encoder(cfg['F'], 256, rnn_type=cfg['rnn'])
	if cfg.get('latent') == 'gaussian':
		smp = model_module.Sampler(enc.hidden_size_total, 256, cfg['f'])
	else:
		smp = model_module.ABCDSampler(enc.hidden_size_total, 256, cfg['K'], cfg['f'])
	dec = model_module.RNN_Variational_Decoder(cfg['F'], 256, 256, cfg['f'], rnn_type=cfg['rnn'],
												num_speakers=cfg.get('num_speakers'), speaker_embed_dim=cfg.get('E'))
	return enc, smp, dec


def build_case(cfg, modules):
	enc, smp, dec = modules
	s = cfg['seed']
	enc_sd, smp_sd, dec_sd = det_state(enc, s * 10 + 1), det_state(smp, s * 10 + 2), det_state(dec, s * 10 + 3)
	seqs = make_inputs(cfg['lengths'], cfg['F'], s)
	sumT = sum(cfg['lengths'])
	eps = O.det_tensor((sumT, cfg['F']), s * 7 + 5, 'normal', 1.0)
	if cfg.get('latent') == 'gaussian':   # N(0,1) draws of the latent sample take the place of the Gumbel noise
		gumbel = O.det_tensor((len(cfg['lengths']), cfg['f']), s * 7 + 6, 'normal', 1.0)
	else:
		gumbel = O.det_tensor((len(cfg['lengths']), cfg['K']), s * 7 + 6, 'gumbel', 1.0)
	speaker = None
	if cfg.get('num_speakers'):
		speaker = torch.tensor([(i * 2 + 1) % cfg['num_speakers'] for i in range(len(cfg['lengths']))])
	return enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker


def sample_vec(t):
	f = t.detach().reshape(-1).double().cpu().numpy()
	return f[::SAMPLE_STRIDE].astype(np.float64)


def synthetic_lengths(B, seed, lo_s=0.03, hi_s=0.3, step_s=0.004):
	"""Finch-syllable-like durations ~ LogUniform(lo, hi) s -> frames T = 1 + floor(dur/step), sorted descending
	(T in [8, 75], mean ~30)."""
	u = O.det_uniform((B,), seed)
	dur = np.exp(np.log(lo_s) + u * (np.log(hi_s) - np.log(lo_s)))
	T = 1 + np.floor(dur / step_s).astype(np.int64)
	return np.sort(T)[::-1].copy()


def encode_lengths(B, seed, lo_s=0.05, hi_s=2.0, fs=16000, hop=64):
	"""Bulk-encode durations ~ U(0.05, 2.0) s at 16 kHz -> T = 1 + floor(dur*fs)//hop in [13, 501]."""
	u = O.det_uniform((B,), seed)
	dur = lo_s + u * (hi_s - lo_s)
	T = 1 + np.floor(dur * fs).astype(np.int64) // hop
	return np.sort(T)[::-1].copy()
