# coding: utf-8
"""Bulk encoding with a trained model: class probabilities (encode.py), raw logits (encode_logit.py) or
to_code_like features (encode_features.py) per segment, mirroring /root/reference ABCD-VAE/encode*.py.

B200 path: batched GPU STFT -> persistent encoder kernel (inference variant, no activations saved) -> sampler MLP
+ codebook logits -> one softmax/arg-max kernel.  Rows are streamed to the CSV in the reference's long format
(data_ix, category_ix, prob [+ annotation columns]) with numpy instead of per-batch pandas melt/merge."""
import argparse
import os
import sys

import numpy as np
import torch

from . import _lib as L
from . import learning
from .modules import data_utils

OUTPUT_KINDS = {
	'prob': ('category_ix', 'prob'),
	'logit': ('dimension', 'logit'),
	'feature': ('dimension', 'feature_value'),
}


class Encoder(learning.Learner):
	def __init__(self, model_config_path, device='cuda'):
		if not str(device).startswith('cuda'):
			raise ValueError('the B200 hot path has no CPU fallback: pass -d cuda')
		self.device = torch.device(device)
		self.process_group = None
		self.save_dir = None
		self.retrieve_model(checkpoint_path=model_config_path, device=device)
		for param in self.parameters():
			param.requires_grad = False
		for m in self._modules():
			m.eval()

	def encode(self, data, is_packed=False, to_numpy=True, output='prob', return_argmax=False):
		"""Per-segment class probabilities (or logits / features) of a (packed) batch."""
		if not is_packed:
			if not isinstance(data, list):
				data = [data]
			data = torch.nn.utils.rnn.pack_sequence(data)
		with torch.no_grad():
			data = data.to(self.device)
			last_hidden = self.encoder(data)
			if output == 'feature':
				out = self.feature_sampler.to_code_like(last_hidden)
				amax = None
			else:
				logits = self.feature_sampler(last_hidden)
				out = logits
				amax = None
				if output == 'prob' or return_argmax:
					B, K = logits.shape
					probs = torch.empty_like(logits)
					amax = torch.empty(B, dtype=torch.int32, device=logits.device)
					L.call('abcd_softmax_argmax', B, K, L.ptr(logits), L.ptr(probs), L.ptr(amax), L.stream_ptr())
					if output == 'prob':
						out = probs
		if return_argmax:
			return out, amax
		if to_numpy:
			out = (p.data.cpu().numpy() for p in out)
		return out

	def encode_dataset(self, dataset, save_path, to_numpy=True, batch_size=1, output='prob', writer_threads=4):
		"""Long-format CSV of the reference (encode.py:38-55): per batch, rows `data_ix, category_ix, prob [+ annotation
		columns]`, category-major like DataFrame.melt, values printed as pandas prints a float32 column.  The GPU encodes batch
		i+1 while host threads format batch i with the native row formatter (abcd_csv_format_long); the file is written in
		batch order.  Returns the number of rows written."""
		import collections
		import concurrent.futures
		dataloader = data_utils.DataLoader(dataset, batch_size=batch_size, device=self.device)
		rename_existing_file(save_path)
		var_name, value_name = OUTPUT_KINDS[output]
		suffix_of, ann_cols = annotation_suffixes(dataset.df_annotation)
		header = ['data_ix', var_name, value_name] + ann_cols
		rows = 0

		def format_batch(vals, ixs):
			return format_long_rows(vals, ixs, suffix_of)

		pending = collections.deque()
		with open(save_path, 'wb') as f, concurrent.futures.ThreadPoolExecutor(max(1, writer_threads)) as pool:
			f.write((','.join(_csv_cell(c) for c in header) + '\n').encode())
			for data, _, _, ix_in_list in dataloader:
				vals = self.encode(data, is_packed=True, to_numpy=False, output=output)
				vals = np.ascontiguousarray(vals.cpu().numpy(), dtype=np.float32)
				ixs = np.ascontiguousarray(np.asarray(ix_in_list), dtype=np.int64)
				rows += vals.size
				pending.append(pool.submit(format_batch, vals, ixs))
				while len(pending) > 2 * writer_threads:
					f.write(memoryview(pending.popleft().result()))
			while pending:
				f.write(memoryview(pending.popleft().result()))
		return rows


def format_long_rows(vals, ixs, suffix_of=None):
	"""Rows of one batch as a uint8 array: for k, for b: `ixs[b],k,vals[b,k]<suffix_of[ixs[b]]>\\n` (host-only native call)."""
	lib = L.load()
	vals = np.ascontiguousarray(vals, dtype=np.float32)
	ixs = np.ascontiguousarray(ixs, dtype=np.int64)
	B, K = vals.shape
	blob, off = None, None
	if suffix_of is not None:
		parts = [suffix_of[ix] for ix in ixs.tolist()]
		off = np.zeros(B + 1, dtype=np.int64)
		np.cumsum([len(p) for p in parts], out=off[1:])
		blob = b''.join(parts)
	cap = B * K * lib.abcd_csv_row_bound(int(np.diff(off).max()) if off is not None and B else 0)
	out = np.empty(max(cap, 1), dtype=np.uint8)
	n = lib.abcd_csv_format_long(vals.ctypes.data, B, K, K, ixs.ctypes.data, blob, off.ctypes.data if off is not None else None,
								out.ctypes.data, cap)
	if n < 0:
		raise L.AbcdError('abcd_csv_format_long failed ({}): {}'.format(n, lib.abcd_last_error().decode()))
	return out[:n]


def annotation_suffixes(df_annotation):
	"""The annotation cells the reference appends to every row of a segment (encode.py:41-44, 50-51: left merge on data_ix
	when the annotation has a `label` column), formatted ONCE per segment by pandas itself: {data_ix: b',cell,cell,...'}."""
	if 'label' not in df_annotation.columns:
		return None, []
	df_ann = df_annotation.drop(columns=['onset_ix', 'offset_ix', 'length'])
	lines = df_ann.to_csv(header=False, index=False).split('\n')
	if lines and lines[-1] == '':
		lines.pop()
	if len(lines) != len(df_ann):   # a cell contains a line break: fall back to per-row formatting
		lines = [df_ann.iloc[[i]].to_csv(header=False, index=False).rstrip('\n') for i in range(len(df_ann))]
	return {ix: (',' + line).encode() for ix, line in zip(df_ann.index.tolist(), lines)}, list(df_ann.columns)


def _csv_cell(v):
	s = str(v)
	if any(c in s for c in ',"\n'):
		s = '"' + s.replace('"', '""') + '"'
	return s


def rename_existing_file(filepath):
	if os.path.isfile(filepath):
		new_path = filepath + '.prev'
		rename_existing_file(new_path)
		os.rename(filepath, new_path)


def get_parameters(argv=None):
	p = argparse.ArgumentParser(description='Encode segments with a trained ABCD-VAE on a B200.')
	p.add_argument('model_path', type=str, help='Path to the checkpoint of a trained model.')
	p.add_argument('input_root', type=str, help='Path to the root directory under which inputs are located.')
	p.add_argument('annotation_file', type=str, help='Path to the annotation csv file.')
	p.add_argument('data_normalizer', type=float, help='Normalizing constant to divide the data.')
	p.add_argument('--annotation_sep', type=str, default=',')
	p.add_argument('-d', '--device', type=str, default='cuda')
	p.add_argument('-S', '--save_path', type=str, default=None)
	p.add_argument('--fft_frame_length', type=float, default=0.008)
	p.add_argument('--fft_step_size', type=float, default=0.004)
	p.add_argument('--fft_window_type', type=str, default='hann_window')
	p.add_argument('--fft_no_centering', action='store_true')
	p.add_argument('--channel', type=int, default=0)
	p.add_argument('-E', '--epsilon', type=float, default=2 ** (-15))
	p.add_argument('-b', '--batch_size', type=int, default=1)
	p.add_argument('--output', type=str, default='prob', choices=sorted(OUTPUT_KINDS),
				help='prob = encode.py, logit = encode_logit.py, feature = encode_features.py')
	return p.parse_args(argv)


def main(argv=None, output=None):
	"""encode.py / encode_logit.py / encode_features.py of the reference share everything but the encoded quantity:
	`output` is fixed by the entry-point module (encode_logit.py -> 'logit', encode_features.py -> 'feature')."""
	a = get_parameters(argv)
	save_path = a.save_path or os.path.join(a.input_root, 'autoencoded.csv')
	os.makedirs(os.path.dirname(os.path.abspath(save_path)), exist_ok=True)
	parser = data_utils.Data_Parser(a.input_root, a.annotation_file, annotation_sep=a.annotation_sep)
	fs = parser.get_sample_freq()
	frame = int(np.floor(a.fft_frame_length * fs))
	step = int(np.floor(a.fft_step_size * fs))
	encoder = Encoder(a.model_path, device=a.device)
	stft = data_utils.STFT(frame, step, window=a.fft_window_type, centering=not a.fft_no_centering, device=a.device)
	transform = data_utils.LogSTFT(stft, epsilon=a.epsilon, normalizer=a.data_normalizer)
	dataset = parser.get_data(transform=transform, channel=a.channel, device=a.device)
	encoder.encode_dataset(dataset, save_path, batch_size=a.batch_size, output=output or a.output)


if __name__ == '__main__':
	main()
