# coding: utf-8
"""GPU suite: data path + CLI surface end to end on a synthetic wav/annotation pair (the reference's toy_data is
not on the GPU box): batching / packing bit-exactness, learning.py (train, validate, LR schedule, checkpoint,
resume) and encode.py (CSV in the reference's long format)."""
import importlib
import os

import numpy as np
import pytest
import scipy.io.wavfile as spw
import torch

from oracle import abcd_oracle as O

pytestmark = pytest.mark.gpu


def _make_corpus(root, n_seg=14, fs=16000, seed=0):
	rng = np.random.RandomState(seed)
	dur = 6.0
	t = np.arange(int(dur * fs)) / fs
	wav = (2500 * np.sin(2 * np.pi * (600 + 900 * t) * t) + 1200 * np.sin(2 * np.pi * 2300 * t) + 400 * rng.randn(t.size)).astype(np.int16)
	spw.write(os.path.join(root, 'a.wav'), fs, wav)
	lens = rng.uniform(0.03, 0.25, n_seg)
	onsets = np.cumsum(np.concatenate([[0.05], lens[:-1] + 0.02]))
	rows = ['onset,offset,input_path,data_type,speaker,label']
	for i, (on, l) in enumerate(zip(onsets, lens)):
		rows.append('{:.6f},{:.6f},a.wav,{},{},{}'.format(on, on + l, 'valid' if i % 5 == 4 else 'train', 'ab'[i % 2], 'XYZ'[i % 3]))
	ann = os.path.join(root, 'ann.csv')
	with open(ann, 'w') as f:
		f.write('\n'.join(rows) + '\n')
	return ann, wav, fs


def test_dataloader_packs_like_reference(abcd, tmp_path):
	root = str(tmp_path)
	ann, wav, fs = _make_corpus(root)
	du = abcd.data_utils
	parser = du.Data_Parser(root, ann)
	stft = du.STFT(128, 64, device='cuda')
	ds = parser.get_data(data_type='train', transform=du.LogSTFT(stft, 2.0 ** -15, 11.0), device='cuda')
	loader = du.DataLoader(ds, batch_size=64, device='cuda')
	packed, is_off, speakers, ixs = next(iter(loader))
	df = ds.df_annotation.loc[ixs]
	# length sort: descending sample counts, stable
	assert (np.diff(df['length'].to_numpy()) <= 0).all()
	# oracle: per-item STFT + log, then time-major packing (bit-exact batch_sizes / is_offset, values within fp32 STFT noise)
	seqs = []
	for _, row in df.iterrows():
		on, off, n = O.discrete_bounds(row['onset'], row['offset'], fs)
		assert (on, off) == (row['onset_ix'], row['offset_ix'])
		seqs.append(torch.from_numpy(O.log_normalize(O.stft_amplitude(wav[on:off], 128, 64), 2.0 ** -15, 11.0)))
	data, bs, _ = O.pack(seqs)
	assert packed.batch_sizes.tolist() == bs
	assert torch.equal(is_off.data.cpu(), O.is_offset_packed(bs))
	assert float((packed.data.cpu().double() - data).abs().max()) < 2e-3
	assert speakers.tolist() == [parser.speaker2ix[s] for s in df['speaker']]


def test_learning_and_encode_cli(abcd, tmp_path):
	root = str(tmp_path)
	ann, wav, fs = _make_corpus(root)
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	encode = importlib.import_module('seq2seq_abcd-vae_b200.encode')
	out = os.path.join(root, 'out')
	argv = [root, ann, '-S', out, '-j', 'job', '-e', '3', '--pretrain_epochs', '1', '-b', '8', '-N', '11', '--speaker_embed_dim', '16']
	learner = learning.main(argv)
	log = open(os.path.join(out, 'job', 'history.log')).read()
	assert 'END OF TRAINING' in log and 'END OF PRETRAINING.' in log
	assert 'mean training total loss (per string)' in log and 'Perplexity of the posterior Dirichlet shape' in log
	losses = [float(l.rsplit(' ', 1)[1]) for l in log.splitlines() if 'mean training total loss' in l]
	assert len(losses) == 3 and all(np.isfinite(losses)) and losses[-1] < losses[0]
	ckpt = torch.load(os.path.join(out, 'job', 'checkpoint.pt'), map_location='cpu', weights_only=False)
	for k in ('epoch', 'encoder', 'encoder_init_parameters', 'feature_sampler', 'feature_sampler_init_parameters', 'decoder',
			'decoder_init_parameters', 'optimizer', 'lr_scheduler', 'gradient_clip', 'random_state'):
		assert k in ckpt
	assert ckpt['epoch'] == 3 and os.path.isfile(os.path.join(out, 'job', 'speaker_coding.json'))
	# resume: an existing history.log switches to retrieval and continues at epoch 4
	learning.main(argv[:-6] + ['-e', '4', '--pretrain_epochs', '1', '-b', '8', '-N', '11', '--speaker_embed_dim', '16'])
	log2 = open(os.path.join(out, 'job', 'history.log')).read()
	assert 'LEARNING RETRIEVED.' in log2 and 'START OF EPOCH:   4' in log2
	# encode
	csv_path = os.path.join(out, 'enc.csv')
	encode.main([os.path.join(out, 'job', 'checkpoint.pt'), root, ann, '11', '-S', csv_path, '-b', '5'])
	import pandas as pd
	df = pd.read_csv(csv_path)
	assert list(df.columns[:3]) == ['data_ix', 'category_ix', 'prob']
	assert len(df) == 14 * 128
	sums = df.groupby('data_ix')['prob'].sum()
	assert np.allclose(sums.to_numpy(), 1.0, atol=1e-4)
