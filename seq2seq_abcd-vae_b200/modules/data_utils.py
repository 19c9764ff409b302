# coding: utf-8
"""Segment dataset, log-STFT front end and length-sorted packed batcher.

API mirror of /root/reference ABCD-VAE/modules/data_utils.py (Data_Parser, Dataset, ToTensor, Transform, STFT,
Compose, DataLoader) with the data path re-laid for the GPU:
  * every wav file is decoded ONCE and cached (int16 on the device) instead of being re-read per item
    (reference: data_utils.py:90-94);
  * a batch's frames are produced by one STFT kernel launch that writes log-amplitudes directly in
    PackedSequence order (reference: per-item torch.stft on the CPU, data_utils.py:131-139, 173-180);
  * the length sort is a stable descending sort on the sample count (the reference's quicksort leaves tie order
    implementation-defined, data_utils.py:81-82).
"""
import ctypes as C
import os.path

import numpy as np
import pandas as pd
import scipy.io.wavfile as spw
import torch
import torch.utils.data

from .. import _lib as L


class Data_Parser(object):
	def __init__(self, input_root, annotation_file, data_type_col_name='data_type', annotation_sep=',', speaker2ix=None):
		self.df_annotation = pd.read_csv(annotation_file, sep=annotation_sep)
		self.input_root = input_root
		self.data_type_col_name = data_type_col_name
		self.index_speakers(speaker2ix)

	def index_speakers(self, speaker2ix):
		if 'speaker' in self.df_annotation.columns:
			self.df_annotation['speaker'] = self.df_annotation.speaker.astype(str)
			if speaker2ix is None:
				self.speaker2ix = {spk: ix for ix, spk in enumerate(self.df_annotation.speaker.unique())}
			else:
				self.speaker2ix = speaker2ix
		else:
			self.speaker2ix = None

	def get_num_speakers(self):
		return 0 if self.speaker2ix is None else len(self.speaker2ix)

	def get_data(self, data_type=None, transform=None, channel=0, device=None):
		if data_type is None:
			sub_df = self.df_annotation.copy()
		else:
			sub_df = self.df_annotation[self.df_annotation[self.data_type_col_name] == data_type].copy()
		return Dataset(sub_df, self.input_root, transform=transform, channel=channel, speaker2ix=self.speaker2ix, device=device)

	def get_sample_freq(self, input_path=None):
		if input_path is None:
			input_path = self.df_annotation.loc[0, 'input_path']
		fs, _ = spw.read(os.path.join(self.input_root, input_path))
		return fs


class Dataset(torch.utils.data.Dataset):
	def __init__(self, df_annotation, input_root, transform=None, channel=0, speaker2ix=None, device=None):
		self.df_annotation = df_annotation
		self.input_root = input_root
		self.transform = transform
		self.channel = channel
		self.speaker2ix = speaker2ix
		self.device = torch.device(device) if device is not None else None
		self._wav_host = {}
		self._wav_dev = {}
		self._corpus = {}   # device -> (one buffer holding every referenced wav back to back, {path: first sample})
		self.get_discrete_bounds()

	def _wav(self, input_path):
		"""Decoded channel of a wav file (host copy), read once."""
		if input_path not in self._wav_host:
			_, data = spw.read(os.path.join(self.input_root, input_path))
			if data.ndim > 1:
				data = data[:, self.channel]
			self._wav_host[input_path] = np.ascontiguousarray(data)
		return self._wav_host[input_path]

	def wav_on_device(self, input_path, device):
		key = (input_path, str(device))
		if key not in self._wav_dev:
			data = self._wav(input_path)
			if data.dtype == np.int16:
				t = torch.from_numpy(data)
			else:
				t = torch.from_numpy(data.astype(np.float32))
			self._wav_dev[key] = t.to(device)
		return self._wav_dev[key]

	def corpus_on_device(self, device):
		"""All wav files this dataset references, decoded once and stored back to back in ONE device buffer (int16 when every
		file is int16, else float32), with the first-sample index of each file.  Batches that touch several files then need no
		per-batch concatenation: segment onsets are rebased with int64 arithmetic and the STFT kernel reads the corpus directly.
		Host copies are dropped once uploaded (the device buffer is the cache; 180 GB of HBM holds ~780 h of 16 kHz int16)."""
		key = str(device)
		if key not in self._corpus:
			paths = list(dict.fromkeys(self.df_annotation['input_path'].tolist()))
			arrays = [self._wav(p) for p in paths]
			all_i16 = all(a.dtype == np.int16 for a in arrays)
			base, off = {}, 0
			for p, a in zip(paths, arrays):
				base[p] = off
				off += int(a.shape[0])
			buf = torch.empty(off, dtype=torch.int16 if all_i16 else torch.float32, device=device)
			for p, a in zip(paths, arrays):
				t = torch.from_numpy(a if all_i16 else a.astype(np.float32))
				buf[base[p]:base[p] + t.numel()].copy_(t)
			self._corpus[key] = (buf, base)
			self._file_len = {p: int(a.shape[0]) for p, a in zip(paths, arrays)}
			self._wav_host.clear()
		return self._corpus[key]

	def get_discrete_bounds(self):
		"""onset/offset in samples = round-half-even(seconds * fs); length = offset_ix - onset_ix (in samples)."""
		onset_ix = np.zeros(len(self.df_annotation), dtype=np.int64)
		offset_ix = np.zeros(len(self.df_annotation), dtype=np.int64)
		paths = self.df_annotation['input_path'].to_numpy()
		onsets = self.df_annotation['onset'].to_numpy(dtype=np.float64)
		offsets = self.df_annotation['offset'].to_numpy(dtype=np.float64)
		fs_of = {}
		for path in pd.unique(paths):
			fs_of[path], _ = spw.read(os.path.join(self.input_root, path))
		for i, path in enumerate(paths):
			onset_ix[i] = int(np.round(onsets[i] * fs_of[path]))
			offset_ix[i] = int(np.round(offsets[i] * fs_of[path]))
		self.df_annotation['onset_ix'] = onset_ix
		self.df_annotation['offset_ix'] = offset_ix
		self.df_annotation['length'] = offset_ix - onset_ix

	def sort_indices_by_length(self, ixs):
		sub = self.df_annotation.iloc[list(ixs), :]
		order = np.argsort(-sub['length'].to_numpy(), kind='stable')
		return sub.index[order]

	def __len__(self):
		return self.df_annotation.shape[0]

	def speaker_of(self, ix):
		if self.speaker2ix is None:
			return float('nan')
		return self.speaker2ix[self.df_annotation.loc[ix, 'speaker']]

	def __getitem__(self, ix):
		row = self.df_annotation.loc[ix]
		data = self._wav(row['input_path'])[int(row['onset_ix']):int(row['offset_ix'])].astype(np.float32)
		speaker = self.speaker_of(ix)
		if self.transform:
			data = self.transform(data)
		return data, speaker


class ToTensor(object):
	"""Convert ndarrays to Tensors."""

	def __call__(self, input_data):
		return torch.from_numpy(input_data)


class Transform(object):
	def __init__(self, in_trans):
		self.in_trans = in_trans

	def __call__(self, input_data):
		return self.in_trans(input_data)


def _stft_launch(wav, onset, nsamp, desc, frame_length, step_size, centering, window, eps, normalizer, out):
	"""onset: int64 [B] first sample of each segment in `wav`; nsamp: int32 [B] samples per segment (both on the device)."""
	L.ensure_device(out.device)
	assert onset.dtype == torch.int64 and nsamp.dtype == torch.int32
	L.call('abcd_stft_logamp', L.ptr(wav), int(wav.dtype == torch.int16), wav.numel(), L.ptr(onset), L.ptr(nsamp), desc.ref(),
		int(frame_length), int(step_size), int(bool(centering)), L.ptr(window), float(eps), float(normalizer),
		L.ptr(out), out.shape[1], L.stream_ptr())


class STFT(object):
	"""Amplitude spectrogram [T, F] of one segment: n_fft = frame_length, Hann (periodic) window by default,
	reflect-padded centring, onesided; T = 1 + n // step_size, F = frame_length // 2 + 1."""

	def __init__(self, frame_length, step_size, window='hann_window', centering=True, device=None):
		self.frame_length = frame_length
		self.step_size = step_size
		self.window = getattr(torch, window)(frame_length)
		self.centering = centering
		self.device = torch.device(device) if device is not None else None
		self._win_dev = {}

	def window_on(self, device):
		key = str(device)
		if key not in self._win_dev:
			self._win_dev[key] = self.window.to(device=device, dtype=torch.float32).contiguous()
		return self._win_dev[key]

	def num_frames(self, n):
		if self.centering:
			return 1 + n // self.step_size
		return 1 + (n - self.frame_length) // self.step_size

	def __call__(self, input_data, eps=0.0, normalizer=0.0):
		dev = self.device or (input_data.device if input_data.is_cuda else torch.device('cuda'))
		wav = input_data.to(dev)
		if wav.dtype not in (torch.int16, torch.float32):
			wav = wav.float()
		wav = wav.contiguous()
		n = wav.numel()
		T = self.num_frames(n)
		desc = L.Packed(torch.ones(T, dtype=torch.int64), dev)
		onset = torch.zeros(1, dtype=torch.int64, device=dev)
		nsamp = torch.full((1,), n, dtype=torch.int32, device=dev)
		out = torch.empty(T, self.frame_length // 2 + 1, dtype=torch.float32, device=dev)
		# normalizer == 0 selects the plain-amplitude output of the kernel
		_stft_launch(wav, onset, nsamp, desc, self.frame_length, self.step_size, self.centering, self.window_on(dev), eps, normalizer, out)
		return out


class LogSTFT(object):
	"""STFT followed by log(x + epsilon) / normalizer, i.e. the reference's Compose([ToTensor, STFT,
	log_and_normalize]) (learning.py:464-466) as one object, so the DataLoader can run a whole batch in one launch."""

	def __init__(self, stft, epsilon=2 ** (-15), normalizer=1.0):
		self.stft = stft
		self.epsilon = epsilon
		self.normalizer = normalizer

	def __call__(self, input_data):
		if isinstance(input_data, np.ndarray):
			input_data = torch.from_numpy(input_data)
		return self.stft(input_data, eps=self.epsilon, normalizer=self.normalizer)


class Compose(object):
	def __init__(self, transforms):
		self.transforms = transforms

	def __call__(self, data):
		for trans in self.transforms:
			data = trans(data)
		return data


def pack_time_major(lengths):
	"""batch_sizes (int64 CPU tensor) of sequences with the given descending frame counts."""
	lengths = np.asarray(lengths, dtype=np.int64)
	assert (np.diff(lengths) <= 0).all(), 'lengths must be sorted in descending order'
	T_max = int(lengths[0])
	# batch_sizes[t] = #{b : lengths[b] > t}
	counts = np.bincount(lengths, minlength=T_max + 1)
	bs = len(lengths) - np.cumsum(counts)[:T_max]
	return torch.from_numpy(bs.astype(np.int64))


class DataLoader(object):
	"""Yields (packed log-STFT, packed is_offset, speakers, indices) per batch, like the reference's DataLoader
	(data_utils.py:150-185): batches come from a BatchSampler (drop_last=False) and are consumed from the end."""

	def __init__(self, dataset, batch_size=1, shuffle=False, device=None, rank=0, world_size=1):
		"""rank / world_size (not in the reference): under data parallelism every rank draws the SAME batches (same torch
		seed) and keeps the strided shard rank, rank + world_size, ... of each length-sorted batch (dist.shard_sorted_batch),
		so only its own segments go through the STFT; `last_global_batch_size` is the divisor of the loss."""
		self.dataset = dataset
		self.shuffle = shuffle
		self.rank, self.world_size = int(rank), int(world_size)
		self.last_global_batch_size = None
		self.last_replicated = False
		if shuffle:
			sampler = torch.utils.data.RandomSampler(self.dataset, replacement=False)
		else:
			sampler = torch.utils.data.SequentialSampler(self.dataset)
		self.batch_sampler = torch.utils.data.BatchSampler(sampler, batch_size, drop_last=False)
		self.device = torch.device(device) if device is not None else (dataset.device or torch.device('cuda'))

	def __iter__(self):
		self.batches = list(self.batch_sampler)
		return self

	def _fused_transform(self):
		tr = self.dataset.transform
		return tr if isinstance(tr, LogSTFT) else None

	def __next__(self):
		if not self.batches:
			raise StopIteration
		ixs = self.batches.pop()
		ixs = self.dataset.sort_indices_by_length(ixs)
		self.last_global_batch_size = len(ixs)
		self.last_replicated = False
		if self.world_size > 1:
			if len(ixs) >= self.world_size:
				ixs = ixs[self.rank::self.world_size]
			else:
				# fewer segments than ranks (tail batch): every rank takes all of them and the divisor grows accordingly
				self.last_global_batch_size = len(ixs) * self.world_size
				self.last_replicated = True
		fused = self._fused_transform()
		if fused is not None:
			batched_input, is_offset = self._next_fused(ixs, fused)
		else:
			seqs = [self.dataset[ix][0] for ix in ixs]
			offs = [torch.tensor([0.0] * (s.size(0) - 1) + [1.0]) for s in seqs]
			batched_input = torch.nn.utils.rnn.pack_sequence(seqs)
			is_offset = torch.nn.utils.rnn.pack_sequence(offs)
		speakers = torch.tensor([self.dataset.speaker_of(ix) for ix in ixs])
		return batched_input, is_offset, speakers, ixs

	def _next_fused(self, ixs, fused):
		df = self.dataset.df_annotation.loc[ixs]
		paths = df['input_path'].to_numpy()
		stft = fused.stft
		dev = self.device
		nsamp = df['length'].to_numpy(dtype=np.int64)
		onset = df['onset_ix'].to_numpy(dtype=np.int64)
		frames = np.array([stft.num_frames(int(n)) for n in nsamp], dtype=np.int64)
		assert (np.diff(frames) <= 0).all()
		batch_sizes = pack_time_major(frames)
		desc = L.Packed(batch_sizes, dev)
		F = stft.frame_length // 2 + 1
		out = torch.empty(desc.sumT, F, dtype=torch.float32, device=dev)
		# every wav file of the dataset lives back to back in one device buffer: rebase the onsets (int64) and launch once
		wav, base = self.dataset.corpus_on_device(dev)
		file_len = np.array([self.dataset._file_len[p] for p in paths], dtype=np.int64)
		if (onset < 0).any() or (onset + nsamp > file_len).any() or int(nsamp.max()) > np.iinfo(np.int32).max:
			raise ValueError('annotation bounds lie outside the decoded audio of {}'.format(sorted(set(paths[(onset + nsamp > file_len) | (onset < 0)]))))
		onset = onset + np.array([base[p] for p in paths], dtype=np.int64)
		onset_d = torch.from_numpy(onset).to(dev)
		nsamp_d = torch.from_numpy(nsamp.astype(np.int32)).to(dev)
		_stft_launch(wav, onset_d, nsamp_d, desc, stft.frame_length, stft.step_size, stft.centering, stft.window_on(dev),
					fused.epsilon, fused.normalizer, out)
		# is_offset: 1.0 at each segment's last frame = packed index offsets[len-1] + b
		is_off = torch.zeros(desc.sumT, dtype=torch.float32, device=dev)
		last = desc.offsets_cpu[desc.lengths_cpu - 1] + torch.arange(desc.B)
		is_off[last.to(dev)] = 1.0
		packed = torch.nn.utils.rnn.PackedSequence(out, batch_sizes)
		packed_off = torch.nn.utils.rnn.PackedSequence(is_off, batch_sizes)
		return packed, packed_off

	def get_num_batches(self):
		return len(self.batch_sampler)
