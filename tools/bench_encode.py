# coding: utf-8
"""Bulk-encode throughput (BASELINE.json configs[4]): batches of B variable-length segments (0.05-2 s at 16 kHz,
T in [13, 501] frames) through encoder (inference variant, nothing saved) -> ABCD sampler logits -> softmax / arg-max.

    python tools/bench_encode.py [--batch 512] [--steps 20]

Prints one JSON line: segments/s and frames/s with the packed log-STFT batch resident in HBM, plus the same with the
STFT front end included (int16 audio resident)."""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--batch', type=int, default=512)
	ap.add_argument('--steps', type=int, default=20)
	ap.add_argument('--bins', type=int, default=65)
	args = ap.parse_args()
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	du = importlib.import_module('seq2seq_abcd-vae_b200.modules.data_utils')
	from oracle import cases as CS
	from oracle import abcd_oracle as O
	L = pkg._lib
	dev = torch.device('cuda')
	L.ensure_device(dev)
	F = args.bins
	lengths = CS.encode_lengths(args.batch, 77)
	sumT = int(lengths.sum())
	bs = torch.tensor(O.batch_sizes_from_lengths(lengths.tolist()), dtype=torch.int64)
	enc = pkg.model.RNN_Variational_Encoder(F, 256).cuda().eval()
	smp = pkg.model.ABCDSampler(enc.hidden_size_total, 256, 128, 256).cuda().eval()
	x = torch.from_numpy(np.clip(np.random.RandomState(0).normal(0.45, 0.25, size=(sumT, F)), -0.945, 1.05).astype(np.float32)).to(dev)
	packed = torch.nn.utils.rnn.PackedSequence(x, bs)
	# audio for the front end
	n_fft, hop = 2 * (F - 1), F - 1
	nsamp = torch.tensor([(int(l) - 1) * hop + hop // 2 for l in lengths], dtype=torch.int32)
	onset = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.int64), nsamp[:-1].long()]), 0).to(dev)
	wav = torch.from_numpy(np.random.RandomState(5).randint(-3000, 3000, size=int(nsamp.sum())).astype(np.int16)).to(dev)
	nsamp = nsamp.to(dev)
	desc = pkg.model.as_packed_desc(bs, dev)
	win = torch.hann_window(n_fft).to(dev)
	spec = torch.empty(sumT, F, device=dev)
	probs = torch.empty(args.batch, 128, device=dev)
	amax = torch.empty(args.batch, dtype=torch.int32, device=dev)

	def encode(p):
		with torch.no_grad():
			logits = smp(enc(p))
			L.call('abcd_softmax_argmax', args.batch, 128, L.ptr(logits), L.ptr(probs), L.ptr(amax), L.stream_ptr())

	def encode_from_audio():
		du._stft_launch(wav, onset, nsamp, desc, n_fft, hop, True, win, 1e-15, 11.0, spec)
		encode(torch.nn.utils.rnn.PackedSequence(spec, bs))

	out = {}
	for name, fn in (('resident_spectrogram', lambda: encode(packed)), ('from_int16_audio', encode_from_audio)):
		for _ in range(3):
			fn()
		e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
		torch.cuda.synchronize(); e0.record()
		for _ in range(args.steps):
			fn()
		e1.record(); torch.cuda.synchronize()
		ms = e0.elapsed_time(e1) / args.steps
		out[name] = {'ms_per_batch': ms, 'segments_per_s': args.batch / (ms * 1e-3), 'frames_per_s': sumT / (ms * 1e-3)}
	print(json.dumps({'metric': 'encode segments/sec', 'n_gpus': 1, 'batch': args.batch, 'sumT': sumT, 'T_max': int(lengths.max()),
					'T_min': int(lengths.min()), 'dtype': 'f32', 'data': 'synthetic', **out}))


if __name__ == '__main__':
	main()
