# coding: utf-8
"""CPU suite, part 3: fixtures written by the reference's own driver classes (oracle/make_golden_runs.py) against the
host-side code of the package -- no GPU needed:
  * the native long-format CSV row formatter + annotation merge against the files written by the reference's
    Encoder.encode_dataset (pandas melt / merge / to_csv) of encode.py, encode_logit.py and encode_features.py, byte for byte;
  * seeded initialisation: the package's modules created in the reference's order under torch.manual_seed(1111) carry the
    same initial weights as the reference's (checksums recorded from the reference Learner), for ABCD-VAE and plain/;
  * the float32 text format against numpy's own on adversarial values."""
import importlib
import io
import os

import numpy as np
import pandas as pd
import pytest
import scipy.io.wavfile as spw
import torch


def _toy_corpus(golden_dir, root):
	g = np.load(os.path.join(golden_dir, 'toy_run.npz'))
	ann_text = str(g['annotation'])
	wav_name = pd.read_csv(io.StringIO(ann_text))['input_path'][0]
	spw.write(os.path.join(root, wav_name), int(g['fs']), g['wav'])
	ann = os.path.join(root, 'annotation.csv')
	with open(ann, 'w') as f:
		f.write(ann_text)
	return ann


def split_reference_csv(text):
	"""-> header, [(data_ix order, values [B, K], raw bytes of the batch)] of a CSV written batch by batch in melt order."""
	lines = text.split('\n')
	header, rows = lines[0], [l for l in lines[1:] if l]
	cats = [int(r.split(',', 3)[1]) for r in rows]
	K = max(cats) + 1
	batches, start = [], 0
	for i in range(1, len(rows) + 1):
		if i == len(rows) or (cats[i] == 0 and cats[i - 1] == K - 1):
			chunk = rows[start:i]
			B = len(chunk) // K
			assert B * K == len(chunk)
			ixs = [int(r.split(',', 1)[0]) for r in chunk[:B]]
			vals = np.array([np.float32(r.split(',', 3)[2]) for r in chunk], dtype=np.float32).reshape(K, B).T
			batches.append((ixs, np.ascontiguousarray(vals), ('\n'.join(chunk) + '\n').encode()))
			start = i
	return header, batches


@pytest.mark.parametrize('script,names', [('encode', ('category_ix', 'prob')), ('encode_logit', ('dimension', 'logit')),
										('encode_features', ('dimension', 'feature_value'))])
def test_csv_rows_match_reference_writer_bytes(abcd, golden_dir, tmp_path, script, names):
	encode = importlib.import_module('seq2seq_abcd-vae_b200.encode')
	g = np.load(os.path.join(golden_dir, 'encode_csv.npz'))
	ann = _toy_corpus(golden_dir, str(tmp_path))
	dataset = abcd.data_utils.Data_Parser(str(tmp_path), ann).get_data()
	suffix_of, ann_cols = encode.annotation_suffixes(dataset.df_annotation)
	header, batches = split_reference_csv(str(g[script + '_csv']))
	assert header == ','.join(['data_ix', names[0], names[1]] + ann_cols)
	assert [len(b[0]) for b in batches] == [2, 3, 3]          # BatchSampler(3) over 8 segments, consumed from the end
	for ixs, vals, raw in batches:
		got = encode.format_long_rows(vals, np.array(ixs), suffix_of).tobytes()
		assert got == raw
		# the batch order is the package's stable length sort of those indices
		assert list(dataset.sort_indices_by_length(sorted(ixs))) == ixs


def test_float32_text_matches_numpy(abcd):
	encode = importlib.import_module('seq2seq_abcd-vae_b200.encode')
	rng = np.random.RandomState(3)
	vals = np.concatenate([rng.rand(500, 8).astype(np.float32) ** 9, (10.0 ** rng.uniform(-40, 38, (500, 8))).astype(np.float32),
							(10.0 ** rng.uniform(-5, -3, (500, 8))).astype(np.float32), (10.0 ** rng.uniform(5, 7, (500, 8))).astype(np.float32)], 1)
	vals[0, :14] = [0, 1, -1, 1e-4, 1e-5, 0.0001000001, 1e16, 9.9999e15, 123456789.0, np.nan, np.inf, -np.inf, 999999.94, 1e6]
	vals[1, :4] = [-0.0, 5e-45, 3.4028235e38, 0.5]
	vals[2] = -vals[3]
	ixs = np.arange(500) * 3 - 7
	text = encode.format_long_rows(vals, ixs).tobytes().decode().split('\n')
	i = 0
	for k in range(vals.shape[1]):
		for b in range(vals.shape[0]):
			assert text[i] == '{},{},{}'.format(ixs[b], k, str(vals[b, k])), (b, k)
			i += 1
	assert text[i] == '' and i == vals.size


@pytest.mark.parametrize('fixture,plain', [('toy_run.npz', False), ('toy_run_plain.npz', True)])
def test_seeded_initial_weights_equal_reference(abcd, golden_dir, fixture, plain):
	"""Same seed + same creation order => same initial weights as the reference Learner (learning.py:84-92)."""
	g = np.load(os.path.join(golden_dir, fixture))
	torch.manual_seed(1111)
	enc = abcd.model.RNN_Variational_Encoder(65, 256, rnn_type='LSTM')
	if plain:
		smp = abcd.model.Sampler(enc.hidden_size_total, 256, 16)
		dec = abcd.model.RNN_Variational_Decoder(65, 256, 256, 16, num_speakers=2, speaker_embed_dim=None)
	else:
		smp = abcd.model.ABCDSampler(enc.hidden_size_total, 256, 128, 256)
		dec = abcd.model.RNN_Variational_Decoder(65, 256, 256, 256, num_speakers=2, speaker_embed_dim=None)
	n = 0
	for name, m in (('encoder', enc), ('feature_sampler', smp), ('decoder', dec)):
		for k, v in m.state_dict().items():
			if not v.dtype.is_floating_point:
				continue
			v = v.double()
			want = g['init/{}/{}'.format(name, k)]
			got = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
			np.testing.assert_allclose(got, want, rtol=1e-12, atol=0, err_msg='{}/{}'.format(name, k))
			n += 1
	assert n == sum(1 for k in g.files if k.startswith('init/'))
