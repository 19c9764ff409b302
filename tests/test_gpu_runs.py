# coding: utf-8
"""GPU suite, part 4: whole jobs against fixtures produced by the reference's own driver classes
(oracle/make_golden_runs.py; the reference is absent on the GPU box):

  * BASELINE configs[0]: `learning.py toy_data annotation -e 20` (defaults, seed 1111) -- the 20-epoch trajectory of training /
    validation losses, the ReduceLROnPlateau schedule and the final parameters against the reference Learner.learn run,
    with the same seeded initial weights and the same (hash-generated) emission / Gumbel noise on both sides;
  * BASELINE configs[1]: the same for plain/learning.py (Gaussian latent);
  * a checkpoint WRITTEN BY the reference Learner.save_model loaded into this Learner / Encoder (parameters, momentum
    buffers, learning rate, scheduler, temperature counters), and the checkpoint this Learner writes compared key by key;
  * encode.py / encode_logit.py / encode_features.py: the CSV this Encoder writes for the toy data against the bytes the
    reference's Encoder.encode_dataset wrote (pandas melt / merge / to_csv);
  * a DataLoader batch that spans several wav files against the per-item oracle front end.
"""
import importlib
import io
import os
import shutil

import numpy as np
import pandas as pd
import pytest
import scipy.io.wavfile as spw
import torch

from oracle import abcd_oracle as O
from oracle.make_golden_runs import EPS_SEED, GUMBEL_SEED, parse_history

pytestmark = pytest.mark.gpu


def _toy_corpus(golden_dir, root):
	g = np.load(os.path.join(golden_dir, 'toy_run.npz'))
	ann_text = str(g['annotation'])
	wav_name = pd.read_csv(io.StringIO(ann_text))['input_path'][0]
	spw.write(os.path.join(root, wav_name), int(g['fs']), g['wav'])
	ann = os.path.join(root, 'annotation.csv')
	with open(ann, 'w') as f:
		f.write(ann_text)
	return ann


def _hook_noise(learner, plain):
	"""Same noise as oracle/make_golden_runs.py: forward call c uses eps = det_tensor((sumT, F), EPS_SEED + c) and Gumbel
	(or latent-Gaussian) noise det_tensor((B, .), GUMBEL_SEED + c)."""
	state = {'call': -1}
	dec, smp = learner.decoder, learner.feature_sampler
	orig_dec = dec.forward

	def dec_forward(features, lengths=None, batch_sizes=None, **kw):
		state['call'] += 1
		dec.injected_emission_noise = O.det_tensor((int(batch_sizes.sum()), dec.rnn_cell.cell.input_size), EPS_SEED + state['call'], 'normal', 1.0).cuda()
		return orig_dec(features, lengths=lengths, batch_sizes=batch_sizes, **kw)
	dec.forward = dec_forward
	if plain:
		orig = smp.sample_and_kl

		def sample_and_kl(params):
			smp.injected_noise = O.det_tensor(tuple(params[0].shape), GUMBEL_SEED + state['call'] + 1, 'normal', 1.0).cuda()
			return orig(params)
		smp.sample_and_kl = sample_and_kl
	else:
		orig = smp.sample

		def sample(logits, no_sample=False):
			if not no_sample:
				smp.injected_gumbel_noise = O.det_tensor(tuple(logits.shape), GUMBEL_SEED + state['call'] + 1, 'gumbel', 1.0).cuda()
			return orig(logits, no_sample=no_sample)
		smp.sample = sample
		orig_fused = smp.forward_sample_kl

		def forward_sample_kl(x, entire_data_size, no_sample=False):   # the Learner's fused path
			if not no_sample:
				smp.injected_gumbel_noise = O.det_tensor((x.shape[0], smp.num_categories), GUMBEL_SEED + state['call'] + 1, 'gumbel', 1.0).cuda()
			return orig_fused(x, entire_data_size, no_sample=no_sample)
		smp.forward_sample_kl = forward_sample_kl
	return state


@pytest.mark.parametrize('fixture,plain', [('toy_run.npz', False), ('toy_run_plain.npz', True)])
def test_toy_20_epoch_trajectory_matches_reference_run(abcd, golden_dir, tmp_path, fixture, plain):
	g = np.load(os.path.join(golden_dir, fixture))
	root = str(tmp_path)
	ann = _toy_corpus(golden_dir, root)
	du = abcd.data_utils
	if plain:
		learning = importlib.import_module('seq2seq_abcd-vae_b200.plain.learning')
		parser = learning.Data_Parser(root, ann)
	else:
		learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
		parser = du.Data_Parser(root, ann)
	fs = parser.get_sample_freq()
	frame, hop = int(np.floor(0.008 * fs)), int(np.floor(0.004 * fs))
	save_dir = learning.get_save_dir(os.path.join(root, 'out'), 'run')
	if plain:
		learner = learning.Learner(int(frame / 2 + 1), 256, 256, 256, 16, save_dir, device='cuda', seed=1111,
									num_speakers=parser.get_num_speakers(), speaker_embed_dim=None)
	else:
		learner = learning.Learner(int(frame / 2 + 1), 256, 256, 256, 128, 256, save_dir, device='cuda', seed=1111,
									num_speakers=parser.get_num_speakers(), speaker_embed_dim=None)
	state = _hook_noise(learner, plain)
	transform = du.LogSTFT(du.STFT(frame, hop, device='cuda'), epsilon=2 ** (-15), normalizer=1.0)
	train = parser.get_data(data_type='train', transform=transform, device='cuda')
	valid = parser.get_data(data_type='valid', transform=transform, device='cuda')
	epochs = int(g['epochs'])
	kw = {} if plain else {'pretrain_epochs': 5}
	learner.learn(train, valid, epochs, 512, 512, learning_rate=1.0, momentum=0.0, gradient_clip=1.0, patience=0, **kw)
	assert state['call'] + 1 == int(g['forward_calls'])
	hist = parse_history(open(os.path.join(save_dir, 'history.log')).read())
	report = {}
	for k in ('training_total', 'validation_total', 'training_emission', 'training_offset', 'training_kl', 'validation_kl', 'batch_loss'):
		want, got = g['hist/' + k], hist[k]
		assert got.shape == want.shape == (epochs,), k
		report[k] = np.abs(got - want) / np.abs(want)
	print(fixture, 'per-epoch relative deviation of the training total loss:', ' '.join('%.1e' % v for v in report['training_total']))
	print(fixture, 'validation total:', ' '.join('%.1e' % v for v in report['validation_total']))
	print(fixture, 'worst', {k: '%.1e' % v.max() for k, v in report.items()})
	np.testing.assert_allclose(hist['lr'], g['hist/lr'], rtol=1e-9)          # identical ReduceLROnPlateau decisions
	# losses are printed with 4 decimals; inputs differ by the STFT round-off (fp32 DFT here, fp32 FFT there)
	for k, v in report.items():
		assert v.max() < 1e-4, (k, v)   # measured on a B200: 4.3e-5 (ABCD-VAE), 1.8e-6 (plain)
	worst = 0.0
	for name, m in (('encoder', learner.encoder), ('feature_sampler', learner.feature_sampler), ('decoder', learner.decoder)):
		for key, val in m.state_dict().items():
			if not val.dtype.is_floating_point:
				continue
			v = val.double()
			got = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
			want = g['final/{}/{}'.format(name, key)]
			worst = max(worst, abs(got[1] - want[1]) / want[1], abs(got[2] - want[2]) / max(want[2], 1e-30))
	print(fixture, 'final parameters: worst relative deviation of sum|w| and sum w^2 per tensor %.1e' % worst)
	assert worst < 1e-3


def test_reference_written_checkpoint_loads(abcd, golden_dir, tmp_path):
	"""learning.py:293-347: a checkpoint.pt saved by the reference Learner (momentum 0.9, 2 epochs, tiny sizes) resumes here."""
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	encode = importlib.import_module('seq2seq_abcd-vae_b200.encode')
	want = np.load(os.path.join(golden_dir, 'ref_checkpoint_tiny.npz'))
	save_dir = str(tmp_path / 'job')
	os.makedirs(save_dir)
	shutil.copy(os.path.join(golden_dir, 'ref_checkpoint_tiny.pt'), os.path.join(save_dir, 'checkpoint.pt'))
	open(os.path.join(save_dir, 'history.log'), 'w').write('previous run\n')
	learner = learning.Learner(65, 8, 8, 8, 5, 6, save_dir, device='cuda')       # an existing history.log => retrieval
	assert learner.retrieval and learner.last_epoch == int(want['epoch'])
	learner.setup_optimizer(0.123, 0.0)                                           # values come from the checkpoint, not from here
	assert learner.gradient_clip == float(want['gradient_clip'])
	assert learner.optimizer.param_groups[0]['lr'] == float(want['lr'])
	assert learner.optimizer.param_groups[0]['momentum'] == float(want['momentum'])
	assert learner.lr_scheduler.best == float(want['scheduler_best'])
	assert learner.lr_scheduler.num_bad_epochs == int(want['scheduler_num_bad_epochs'])
	assert learner.feature_sampler.iter_counts == int(want['iter_counts'])
	names = []
	for mod_name, m in (('encoder', learner.encoder), ('feature_sampler', learner.feature_sampler), ('decoder', learner.decoder)):
		names += ['{}/{}'.format(mod_name, k) for k, _ in m.named_parameters()]
	assert names == list(want['param_order'])
	assert learner.arena.momentum_buf is not None
	for n, p, o in zip(names, learner.arena.params, learner.arena.offsets):
		assert torch.equal(p.detach().cpu(), torch.from_numpy(want['param/' + n])), n
		mom = learner.arena.momentum_buf[o:o + p.numel()].view(p.shape).cpu()
		assert torch.equal(mom, torch.from_numpy(want['momentum/' + n])), n
	# and it writes back a checkpoint with the reference's structure (momentum as optimizer.state[i]['momentum_buffer'])
	ref = torch.load(os.path.join(golden_dir, 'ref_checkpoint_tiny.pt'), map_location='cpu', weights_only=False)
	learner.save_model(int(want['epoch']))
	mine = torch.load(os.path.join(save_dir, 'checkpoint.pt'), map_location='cpu', weights_only=False)
	assert set(ref.keys()) <= set(mine.keys())
	assert set(mine.keys()) - set(ref.keys()) <= {'random_state_cuda', 'noise_state'}   # (the reference adds random_state_cuda on GPUs too)
	for k in ('encoder', 'feature_sampler', 'decoder'):
		assert list(ref[k].keys()) == list(mine[k].keys())
		for kk in ref[k]:
			assert torch.equal(ref[k][kk], mine[k][kk]), (k, kk)
		assert ref[k + '_init_parameters'] == mine[k + '_init_parameters'], k
	assert ref['optimizer']['param_groups'] == mine['optimizer']['param_groups']
	assert sorted(ref['optimizer']['state'].keys()) == sorted(mine['optimizer']['state'].keys())
	for i in ref['optimizer']['state']:
		assert set(mine['optimizer']['state'][i].keys()) == {'momentum_buffer'}
		assert torch.equal(ref['optimizer']['state'][i]['momentum_buffer'], mine['optimizer']['state'][i]['momentum_buffer']), i
	assert ref['lr_scheduler'] == mine['lr_scheduler'] and ref['epoch'] == mine['epoch']
	# the encode-side loader takes the same file
	enc = encode.Encoder(os.path.join(golden_dir, 'ref_checkpoint_tiny.pt'), device='cuda')
	assert not any(p.requires_grad for p in enc.parameters())


@pytest.mark.parametrize('script,kind', [('encode', 'prob'), ('encode_logit', 'logit'), ('encode_features', 'feature')])
def test_encode_scripts_write_the_reference_csv(abcd, golden_dir, tmp_path, script, kind):
	"""The whole encode_dataset path (DataLoader batching / length sort / GPU STFT / writer threads) with the per-segment values
	taken from the reference's CSV (its tiny model cannot run on the specialised kernels): the file must be byte-identical."""
	from test_cpu_fixtures import split_reference_csv
	mod = importlib.import_module('seq2seq_abcd-vae_b200.' + script)
	g = np.load(os.path.join(golden_dir, 'encode_csv.npz'))
	text = str(g[script + '_csv'])
	_, batches = split_reference_csv(text)
	lookup = {}
	for ixs, vals, _ in batches:
		for b, ix in enumerate(ixs):
			lookup[ix] = vals[b]
	root = str(tmp_path)
	ann = _toy_corpus(golden_dir, root)
	du = abcd.data_utils
	parser = du.Data_Parser(root, ann)
	dataset = parser.get_data(transform=du.LogSTFT(du.STFT(128, 64, device='cuda'), 2 ** (-15), 11.0), device='cuda')
	enc = mod.Encoder.__new__(mod.Encoder)
	enc.device = torch.device('cuda')
	seen = []

	def fake_encode(data, is_packed=False, to_numpy=True, output=kind, return_argmax=False):
		assert output == kind and is_packed and data.data.is_cuda
		seen.append(data.batch_sizes.tolist())
		return torch.from_numpy(np.stack([lookup[ix] for ix in fake_encode.ixs]))
	enc.encode = fake_encode
	orig_next = du.DataLoader.__next__

	def spy_next(self):
		out = orig_next(self)
		fake_encode.ixs = list(out[3])
		return out
	du.DataLoader.__next__ = spy_next
	try:
		path = os.path.join(root, script + '.csv')
		rows = enc.encode_dataset(dataset, path, batch_size=3)
	finally:
		du.DataLoader.__next__ = orig_next
	assert open(path).read() == text
	assert rows == text.count('\n') - 1 and len(seen) == 3


def test_multi_file_batch_matches_oracle_front_end(abcd, tmp_path):
	"""A batch whose segments come from three wav files (one of them float32) through the single corpus buffer + int64
	onsets against the per-item oracle STFT."""
	root = str(tmp_path)
	rng = np.random.RandomState(4)
	fs = 16000
	wavs = {}
	for i, name in enumerate(('a.wav', 'b.wav', 'c.wav')):
		t = np.arange(int((1.5 + i) * fs)) / fs
		w = 3000 * np.sin(2 * np.pi * (500 + 400 * i + 700 * t) * t) + 300 * rng.randn(t.size)
		wavs[name] = w.astype(np.int16)
		spw.write(os.path.join(root, name), fs, wavs[name])
	rows = ['onset,offset,input_path,data_type']
	segs = []
	for j in range(12):
		name = ('a.wav', 'b.wav', 'c.wav')[j % 3]
		on = 0.05 + 0.11 * (j // 3)
		off = on + 0.03 + 0.013 * j
		rows.append('{:.6f},{:.6f},{},train'.format(on, off, name))
		segs.append((name, on, off))
	ann = os.path.join(root, 'ann.csv')
	open(ann, 'w').write('\n'.join(rows) + '\n')
	du = abcd.data_utils
	ds = du.Data_Parser(root, ann).get_data(transform=du.LogSTFT(du.STFT(128, 64, device='cuda'), 2.0 ** -15, 11.0), device='cuda')
	packed, is_off, _, ixs = next(iter(du.DataLoader(ds, batch_size=64, device='cuda')))
	seqs = []
	for ix in ixs:
		name, on, off = segs[ix]
		a, b, _ = O.discrete_bounds(on, off, fs)
		seqs.append(torch.from_numpy(O.log_normalize(O.stft_amplitude(wavs[name][a:b], 128, 64), 2.0 ** -15, 11.0)))
	data, bs, _ = O.pack(seqs)
	assert packed.batch_sizes.tolist() == bs
	assert torch.equal(is_off.data.cpu(), O.is_offset_packed(bs))
	assert float((packed.data.cpu().double() - data).abs().max()) < 2e-3
	buf, base = ds.corpus_on_device(torch.device('cuda'))
	assert buf.dtype == torch.int16 and buf.numel() == sum(len(w) for w in wavs.values()) and len(base) == 3
