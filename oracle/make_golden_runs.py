# coding: utf-8
"""Golden fixtures that need the reference's own DRIVER classes (Learner / Encoder / DataLoader), not just its modules --
TEST INFRASTRUCTURE ONLY; runs in the build container where /root/reference exists:

    python oracle/make_golden_runs.py

  tests/golden/toy_run.npz            BASELINE configs[0]: `ABCD-VAE/learning.py toy_data annotation -e 20` (all defaults, seed
                                      1111, CPU) through the UNMODIFIED reference Learner.learn: per-epoch training / validation
                                      losses, learning rates, seeded initial-weight checksums, final-parameter checksums, and the
                                      toy recording itself (int16) so that the GPU box can rerun the same job.
  tests/golden/toy_run_plain.npz      BASELINE configs[1]: the same for plain/learning.py (Gaussian latent, -f 16).
  tests/golden/ref_checkpoint_tiny.pt a checkpoint WRITTEN BY the reference Learner.save_model (tiny hidden sizes, momentum 0.9,
                                      2 epochs on the toy data) + ref_checkpoint_tiny.npz with what a loader must recover.
  tests/golden/encode_csv.npz         the CSVs written by the reference Encoder.encode_dataset of encode.py / encode_logit.py /
                                      encode_features.py (pandas melt + merge + to_csv) for that checkpoint, batch size 3.

Compat shims (applied by monkey-patching the imported reference modules, /root/reference is never edited):
STFT return_complex; integer onset/offset columns; torch._six.inf; plain/ speaker column assignment.  Noise: the reference draws emission / Gumbel noise
from torch's global generator; to make the trajectory reproducible by the CUDA path (whose noise is Philox) both sides use
the oracle's hash-based generator keyed by the forward-call counter: call c uses eps = det_tensor((sumT, F), 50000 + c),
gumbel = det_tensor((B, K), 60000 + c).
"""
import importlib
import io
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import abcd_oracle as O  # noqa: E402

REF = os.environ.get('ABCD_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')
TOY_WAV = '20170806-080002_89.2-94.22.1ch.wav'
TOY_ANN = 'annotation_20170806-080002_89.2-94.22.csv'
EPS_SEED, GUMBEL_SEED = 50000, 60000


def import_reference(package):
	"""Import <REF>/<package>/learning.py (+ encode*.py) with their `modules` package, isolated from other variants."""
	for name in [n for n in sys.modules if n == 'modules' or n.startswith('modules.') or n in ('learning', 'encode', 'encode_logit', 'encode_features')]:
		del sys.modules[name]
	path = os.path.join(REF, package)
	sys.path.insert(0, path)
	try:
		if not hasattr(torch, '_six'):
			torch._six = types.SimpleNamespace(inf=float('inf'))
		learning = importlib.import_module('learning')
		mods = {'learning': learning, 'data_utils': sys.modules['modules.data_utils'], 'model': sys.modules['modules.model']}
		for extra in ('encode', 'encode_logit', 'encode_features'):
			if os.path.isfile(os.path.join(path, extra + '.py')):
				mods[extra] = importlib.import_module(extra)
	finally:
		sys.path.remove(path)
	du = mods['data_utils']

	def stft_call(self, input_data):   # data_utils.py:131-139 with return_complex=True (torch >= 2)
		tr = torch.view_as_real(input_data.stft(self.frame_length, hop_length=self.step_size, window=self.window,
												center=self.centering, return_complex=True))
		return tr.pow(2).sum(-1).sqrt().transpose(0, 1).contiguous()
	du.STFT.__call__ = stft_call

	def get_discrete_bounds(self):     # data_utils.py:69-79 with integer columns (pandas >= 2 refuses the in-place cast)
		import scipy.io.wavfile as spw
		on = np.zeros(len(self.df_annotation), dtype=np.int64)
		off = np.zeros(len(self.df_annotation), dtype=np.int64)
		pos = {ix: i for i, ix in enumerate(self.df_annotation.index)}
		for input_path, sub_df in self.df_annotation.groupby('input_path'):
			fs, _ = spw.read(os.path.join(self.input_root, input_path))
			o1 = (sub_df.onset * fs).round()
			o2 = (sub_df.offset * fs).round()
			for ix in sub_df.index:
				on[pos[ix]] = int(o1[ix]); off[pos[ix]] = int(o2[ix])
		self.df_annotation['onset_ix'] = on
		self.df_annotation['offset_ix'] = off
		self.df_annotation['length'] = off - on
	du.Dataset.get_discrete_bounds = get_discrete_bounds
	if package == 'plain':
		def index_speakers(self):          # plain/modules/data_utils.py:17-22 (pandas >= 3 refuses ints in a str column via .loc)
			if 'speaker' in self.df_annotation.columns:
				speaker2ix = {spk: ix for ix, spk in enumerate(self.df_annotation.speaker.unique())}
				self.df_annotation['speaker'] = self.df_annotation.speaker.map(speaker2ix)
			else:
				self.df_annotation['speaker'] = float('nan')
		du.Data_Parser.index_speakers = index_speakers
	return mods


class NoiseClock(object):
	"""Deterministic noise keyed by the forward-call counter (see the module docstring)."""

	def __init__(self):
		self.call = -1
		self.eps = None
		self.cursor = 0

	def begin_forward(self, sumT, F):
		self.call += 1
		self.eps = O.det_tensor((sumT, F), EPS_SEED + self.call, 'normal', 1.0)
		self.cursor = 0

	def emission(self, mean, log_variance):
		n = mean.shape[0]
		e = self.eps[self.cursor:self.cursor + n]
		self.cursor += n
		return mean + (0.5 * log_variance).exp() * e


def hook_noise(learner, clock, gaussian_latent=False):
	dec = learner.decoder
	dec.emission_sampler._sampler = clock.emission
	orig_forward = dec.forward

	def forward(features, lengths=None, batch_sizes=None, **kw):
		clock.begin_forward(int(batch_sizes.sum()), dec.rnn_cell.cell.input_size)
		return orig_forward(features, lengths=lengths, batch_sizes=batch_sizes, **kw)
	dec.forward = forward
	if gaussian_latent:
		# plain: z = mean + exp(0.5 lv) * eps_z with eps_z keyed like the Gumbel noise of the ABCD variant.  The latent is drawn
		# BEFORE the decoder forward of the same step, i.e. when clock.call still counts the previous forward: use call + 1.
		def latent(mean, log_variance):
			return mean + (0.5 * log_variance).exp() * O.det_tensor(tuple(mean.shape), GUMBEL_SEED + clock.call + 1, 'normal', 1.0)
		learner.feature_sampler._sampler = latent


def patched_gumbel(clock):
	def gs(logits, tau=1, hard=False, eps=1e-10, dim=-1):
		g = O.det_tensor(tuple(logits.shape), GUMBEL_SEED + clock.call + 1, 'gumbel', 1.0)
		return ((logits + g) / tau).softmax(dim)
	return gs


def tensor_checksums(modules):
	out = {}
	for name, m in modules:
		for k, v in m.state_dict().items():
			if v.dtype.is_floating_point:
				v = v.double()
				out['{}/{}'.format(name, k)] = np.array([float(v.sum()), float(v.abs().sum()), float((v * v).sum())])
	return out


def parse_history(text):
	out = {}
	for phase in ('training', 'validation'):
		for what, key in (('emission negative pdf loss', 'emission'), ('end-prediction loss', 'offset'), ('KL', 'kl'), ('total loss', 'total')):
			pat = r'mean {} {} \(per string\): ([-0-9.eE+naif]+)'.format(phase, re.escape(what))
			out['{}_{}'.format(phase, key)] = np.array([float(v) for v in re.findall(pat, text)])
	out['lr'] = np.array([float(v) for v in re.findall(r'current learning rate: ([-0-9.eE+]+)', text)])
	out['batch_loss'] = np.array([float(v) for v in re.findall(r'training batches complete. mean loss: (-?[0-9]+\.[0-9]+|nan|inf)', text)])
	return out


def run_toy(package, out_name, extra_kwargs, learn_kwargs, epochs=20, gaussian=False):
	mods = import_reference(package)
	learning, du, model = mods['learning'], mods['data_utils'], mods['model']
	import scipy.io.wavfile as spw
	toy = os.path.join(REF, 'toy_data')
	fs, wav = spw.read(os.path.join(toy, TOY_WAV))
	tmp = tempfile.mkdtemp()
	try:
		save_dir = learning.get_save_dir(tmp, 'run')
		parser = du.Data_Parser(toy, os.path.join(toy, TOY_ANN))
		frame, hop = int(np.floor(0.008 * fs)), int(np.floor(0.004 * fs))
		args = [int(frame / 2 + 1), 256, 256, 256] + extra_kwargs.pop('positional') + [save_dir]
		learner = learning.Learner(*args, device='cpu', seed=1111, num_speakers=parser.get_num_speakers(), speaker_embed_dim=None, **extra_kwargs)
		init_sums = tensor_checksums([('encoder', learner.encoder), ('feature_sampler', learner.feature_sampler), ('decoder', learner.decoder)])
		clock = NoiseClock()
		hook_noise(learner, clock, gaussian_latent=gaussian)
		orig_gs = torch.nn.functional.gumbel_softmax
		torch.nn.functional.gumbel_softmax = patched_gumbel(clock)
		try:
			eps, N = 2 ** (-15), 1.0
			tr = du.Compose([du.ToTensor(), du.STFT(frame, hop), du.Transform(lambda x: (x + eps).log() / N)])
			train = parser.get_data(data_type='train', transform=tr)
			valid = parser.get_data(data_type='valid', transform=tr)
			learner.learn(train, valid, epochs, 512, 512, learning_rate=1.0, momentum=0.0, gradient_clip=1.0, patience=0, **learn_kwargs)
		finally:
			torch.nn.functional.gumbel_softmax = orig_gs
		text = open(os.path.join(save_dir, 'history.log')).read()
		hist = parse_history(text)
		final_sums = tensor_checksums([('encoder', learner.encoder), ('feature_sampler', learner.feature_sampler), ('decoder', learner.decoder)])
		out = {'wav': wav.astype(np.int16), 'fs': np.int64(fs), 'annotation': np.array(open(os.path.join(toy, TOY_ANN)).read()),
				'epochs': np.int64(epochs), 'forward_calls': np.int64(clock.call + 1)}
		out.update({'hist/' + k: v for k, v in hist.items()})
		out.update({'init/' + k: v for k, v in init_sums.items()})
		out.update({'final/' + k: v for k, v in final_sums.items()})
		np.savez_compressed(os.path.join(GOLD, out_name), **out)
		print(out_name, 'epochs', epochs, 'train total', hist['training_total'][[0, -1]], 'valid total', hist['validation_total'][[0, -1]],
			'lr', hist['lr'][[0, -1]], 'forward calls', clock.call + 1)
	finally:
		shutil.rmtree(tmp)
		for h in list(learning.logger.handlers):
			learning.logger.removeHandler(h)


def gen_checkpoint_and_csv():
	"""A real reference checkpoint (tiny sizes) and the CSVs of the three reference encode scripts for it."""
	mods = import_reference('ABCD-VAE')
	learning, du = mods['learning'], mods['data_utils']
	import scipy.io.wavfile as spw
	toy = os.path.join(REF, 'toy_data')
	fs, wav = spw.read(os.path.join(toy, TOY_WAV))
	tmp = tempfile.mkdtemp()
	try:
		save_dir = learning.get_save_dir(tmp, 'tiny')
		parser = du.Data_Parser(toy, os.path.join(toy, TOY_ANN))
		frame, hop = int(np.floor(0.008 * fs)), int(np.floor(0.004 * fs))
		learner = learning.Learner(int(frame / 2 + 1), 8, 8, 8, 5, 6, save_dir, device='cpu', seed=7,
									num_speakers=parser.get_num_speakers(), speaker_embed_dim=None)
		tr = du.Compose([du.ToTensor(), du.STFT(frame, hop), du.Transform(lambda x: (x + 2 ** (-15)).log() / 11.0)])
		train = parser.get_data(data_type='train', transform=tr)
		valid = parser.get_data(data_type='valid', transform=tr)
		learner.learn(train, valid, 2, 4, 4, pretrain_epochs=1, learning_rate=0.5, momentum=0.9, gradient_clip=1.0, patience=0)
		ckpt_path = os.path.join(save_dir, 'checkpoint.pt')
		shutil.copy(ckpt_path, os.path.join(GOLD, 'ref_checkpoint_tiny.pt'))
		ck = torch.load(ckpt_path, map_location='cpu', weights_only=False)
		expect = {'epoch': np.int64(ck['epoch']), 'gradient_clip': np.float64(ck['gradient_clip']),
				'lr': np.float64(ck['optimizer']['param_groups'][0]['lr']), 'momentum': np.float64(ck['optimizer']['param_groups'][0]['momentum']),
				'scheduler_best': np.float64(ck['lr_scheduler']['best']), 'scheduler_num_bad_epochs': np.int64(ck['lr_scheduler']['num_bad_epochs']),
				'iter_counts': np.int64(ck['feature_sampler_init_parameters']['epoch_init_iter_counts'])}
		params = list(learner.parameters())
		names = []
		for mod_name, m in (('encoder', learner.encoder), ('feature_sampler', learner.feature_sampler), ('decoder', learner.decoder)):
			names += ['{}/{}'.format(mod_name, k) for k, _ in m.named_parameters()]
		assert len(names) == len(params)
		for i, (n, p) in enumerate(zip(names, params)):
			expect['param/' + n] = p.detach().numpy().copy()
			expect['momentum/' + n] = ck['optimizer']['state'][i]['momentum_buffer'].numpy().copy()
		expect['param_order'] = np.array(names)
		np.savez_compressed(os.path.join(GOLD, 'ref_checkpoint_tiny.npz'), **expect)
		print('ref_checkpoint_tiny.pt', os.path.getsize(ckpt_path), 'bytes; keys', sorted(ck.keys()))
		for h in list(learning.logger.handlers):
			learning.logger.removeHandler(h)

		# the three encode scripts of the reference on that checkpoint (batch size 3: three batches, consumed from the end)
		out = {'annotation': np.array(open(os.path.join(toy, TOY_ANN)).read())}
		for script, value_name in (('encode', 'prob'), ('encode_logit', 'logit'), ('encode_features', 'feature_value')):
			enc_mod = mods[script]
			encoder = enc_mod.Encoder(ckpt_path, device='cpu')
			dataset = parser.get_data(transform=tr)
			csv_path = os.path.join(tmp, script + '.csv')
			encoder.encode_dataset(dataset, csv_path, batch_size=3)
			text = open(csv_path).read()
			out[script + '_csv'] = np.array(text)
			print(script, 'rows', text.count('\n') - 1, 'header', text.split('\n', 1)[0])
		np.savez_compressed(os.path.join(GOLD, 'encode_csv.npz'), **out)
	finally:
		shutil.rmtree(tmp)


if __name__ == '__main__':
	assert os.path.isdir(REF), 'the reference checkout is needed to (re)generate golden vectors'
	torch.set_num_threads(8)
	which = sys.argv[1:] or ['toy', 'plain', 'ckpt']
	if 'toy' in which:
		run_toy('ABCD-VAE', 'toy_run.npz', {'positional': [128, 256]}, {'pretrain_epochs': 5})
	if 'plain' in which:
		run_toy('plain', 'toy_run_plain.npz', {'positional': [16]}, {}, gaussian=True)
	if 'ckpt' in which:
		gen_checkpoint_and_csv()
