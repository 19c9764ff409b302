# coding: utf-8
"""Synthetic workloads of the publication shape for bench.py and tools/: deterministic functions of an
integer seed (SplitMix64 hash -> uniform), identical on every machine and rank, independent of torch's generators.

  synthetic_lengths   training segments: duration ~ LogUniform(0.03, 0.3) s, T = 1 + floor(dur / 4 ms) in [8, 75]
  encode_lengths      bulk-encode segments: duration ~ U(0.05, 2.0) s at 16 kHz, T = 1 + floor(dur * fs) // 64 in [13, 501]
  synthetic_frames    log-STFT-like frames clip(N(0.45, 0.25^2), -0.945, 1.05)
  deterministic_state weights with PyTorch-like scales keyed by state_dict name
(tests/test_cpu_lib.py checks that these agree with the test infrastructure's own generators in oracle/cases.py.)
"""
import numpy as np
import torch


def _splitmix64(x):
	x = x + np.uint64(0x9E3779B97F4A7C15)
	z = x
	z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
	z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
	return z ^ (z >> np.uint64(31))


def hash_uniform(shape, seed, lo=0.0, hi=1.0):
	"""float64 uniforms in (lo, hi): a pure function of (shape, seed)."""
	n = int(np.prod(shape)) if len(shape) else 1
	with np.errstate(over='ignore'):
		idx = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x100000001B3)
		bits = _splitmix64(_splitmix64(idx))
	u = ((bits >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)
	return (lo + (hi - lo) * u).reshape(shape)


def hash_normal(shape, seed):
	u1 = hash_uniform(shape, seed * 2 + 1)
	u2 = hash_uniform(shape, seed * 2 + 2)
	return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def synthetic_lengths(B, seed, lo_s=0.03, hi_s=0.3, step_s=0.004):
	u = hash_uniform((B,), seed)
	dur = np.exp(np.log(lo_s) + u * (np.log(hi_s) - np.log(lo_s)))
	T = 1 + np.floor(dur / step_s).astype(np.int64)
	return np.sort(T)[::-1].copy()


def encode_lengths(B, seed, lo_s=0.05, hi_s=2.0, fs=16000, hop=64):
	u = hash_uniform((B,), seed)
	dur = lo_s + u * (hi_s - lo_s)
	T = 1 + np.floor(dur * fs).astype(np.int64) // hop
	return np.sort(T)[::-1].copy()


def batch_sizes_of(lengths):
	"""PackedSequence.batch_sizes of descending frame counts."""
	lengths = np.asarray(lengths, dtype=np.int64)
	counts = np.bincount(lengths, minlength=int(lengths[0]) + 1)
	return torch.from_numpy((len(lengths) - np.cumsum(counts)[:int(lengths[0])]).astype(np.int64))


def synthetic_frames(sumT, F, seed):
	rng = np.random.RandomState(seed)
	return torch.from_numpy(np.clip(rng.normal(0.45, 0.25, size=(sumT, F)), -0.945, 1.05).astype(np.float32))


def is_offset_of(lengths, batch_sizes):
	"""Packed one-hot of each segment's last frame."""
	offsets = np.concatenate([[0], np.cumsum(np.asarray(batch_sizes))])
	out = np.zeros(int(offsets[-1]), dtype=np.float32)
	for b, l in enumerate(lengths):
		out[offsets[int(l) - 1] + b] = 1.0
	return torch.from_numpy(out)


def deterministic_state(module, seed):
	sd = {}
	for i, (k, v) in enumerate(sorted(module.state_dict().items())):
		if k == 'prior_concentration':
			sd[k] = v.clone()
			continue
		if k in ('codebook', 'posterior_shape_logits') or k.startswith('embed_speaker'):
			a = hash_normal(tuple(v.shape), seed * 1000 + i)
		else:
			fan = v.shape[-1] if v.dim() > 1 else v.shape[0]
			s = 1.0 / np.sqrt(max(fan, 1))
			a = hash_uniform(tuple(v.shape), seed * 1000 + i, -s, s)
		sd[k] = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(torch.float32)
	return sd
