# coding: utf-8
"""plain/learning.py of the reference on the B200 hot path: the VAE with an isotropic Gaussian latent
(/root/reference plain/learning.py:36-310, CLI flags :327-368).  Everything but the latent is shared with the ABCD-VAE
Learner (..learning): flat parameter / gradient arenas, encoder and decoder kernels, NCCL all-reduce, fused clip + SGD,
checkpoint keys, log lines.  Differences that follow the reference: positional `feature_size` in place of
(num_feature_categories, feature_dim), no pretraining epochs, the short per-batch log line, speakers indexed by
first appearance without a speaker_coding.json.
"""
import argparse
import os

import numpy as np
import torch

from .. import learning as base
from ..learning import logger, get_save_dir   # noqa: F401  (same logger: history.log format is shared)
from ..modules import data_utils, model


class Learner(base.Learner):
	has_pretraining = False

	def __init__(self, input_size, encoder_rnn_hidden_size, decoder_rnn_hidden_size, mlp_hidden_size, feature_size, save_dir,
				encoder_rnn_type='LSTM', decoder_rnn_type='LSTM', encoder_rnn_layers=1, bidirectional_encoder=True,
				bidirectional_decoder=False, right2left_decoder_weight=0.5, encoder_hidden_dropout=0.0, decoder_input_dropout=0.0,
				device='cuda', seed=1111, feature_distribution='isotropic_gaussian', emission_distribution='isotropic_gaussian',
				decoder_self_feedback=True, esn_leak=1.0, num_speakers=None, speaker_embed_dim=None, process_group=None):
		self.feature_distribution = feature_distribution
		super(Learner, self).__init__(
			input_size, encoder_rnn_hidden_size, decoder_rnn_hidden_size, mlp_hidden_size, None, feature_size, save_dir,
			encoder_rnn_type=encoder_rnn_type, decoder_rnn_type=decoder_rnn_type, encoder_rnn_layers=encoder_rnn_layers,
			bidirectional_encoder=bidirectional_encoder, bidirectional_decoder=bidirectional_decoder,
			right2left_decoder_weight=right2left_decoder_weight, encoder_hidden_dropout=encoder_hidden_dropout,
			decoder_input_dropout=decoder_input_dropout, device=device, seed=seed, emission_distribution=emission_distribution,
			decoder_self_feedback=decoder_self_feedback, esn_leak=esn_leak, num_speakers=num_speakers,
			speaker_embed_dim=speaker_embed_dim, process_group=process_group)

	# -- latent-variable hooks ---------------------------------------------------------------------------
	def _build_sampler(self, hidden_size_total, mlp_hidden_size, _num_categories, feature_size, _prior_concentration):
		logger.info('Data to be encoded into {feature_size}-dim features.'.format(feature_size=feature_size))
		logger.info('Features are assumed to be distributed according to {feature_distribution}.'.format(feature_distribution=self.feature_distribution))
		return model.Sampler(hidden_size_total, mlp_hidden_size, feature_size, distribution_name=self.feature_distribution)

	def _sampler_from_checkpoint(self, init_parameters):
		return model.Sampler(**init_parameters)

	def _latent(self, last_hidden, num_strings, is_pretraining):
		"""plain/learning.py:147-149: feature_params -> sample, KL (one fused kernel here)."""
		feature_params = self.feature_sampler(last_hidden)
		features, kl_loss = self.feature_sampler.sample_and_kl(feature_params)
		return features, kl_loss, None

	def _log_batch(self, batch_ix, num_batches, loss, logits):
		logger.info('{batch_ix}/{num_batches} training batches complete. mean loss: {loss:5.4f}'.format(
			batch_ix=batch_ix, num_batches=num_batches, loss=loss.item()))

	# -- the reference's plain signatures (no is_pretraining / pretrain_epochs) ----------------------------
	def train(self, dataloader, is_pretraining=False):
		return super(Learner, self).train(dataloader, False)

	def test_or_validate(self, dataloader, is_pretraining=False):
		return super(Learner, self).test_or_validate(dataloader, False)

	def learn(self, train_dataset, valid_dataset, num_epochs, batch_size_train, batch_size_valid, learning_rate=0.1,
			momentum=0.9, gradient_clip=0.25, patience=0):
		return super(Learner, self).learn(train_dataset, valid_dataset, num_epochs, batch_size_train, batch_size_valid,
										pretrain_epochs=0, learning_rate=learning_rate, momentum=momentum,
										gradient_clip=gradient_clip, patience=patience)


class Data_Parser(data_utils.Data_Parser):
	"""plain/modules/data_utils.py:11-25: speakers are numbered by first appearance; without a speaker column there is one
	(nan) speaker, so get_num_speakers() is 1."""

	def __init__(self, input_root, annotation_file, data_type_col_name='data_type', annotation_sep=','):
		super(Data_Parser, self).__init__(input_root, annotation_file, data_type_col_name=data_type_col_name,
										annotation_sep=annotation_sep, speaker2ix=None)

	def get_num_speakers(self):
		return 1 if self.speaker2ix is None else len(self.speaker2ix)


def get_parameters(argv=None):
	p = argparse.ArgumentParser(description='Train the plain (Gaussian-latent) seq2seq VAE on a B200.')
	p.add_argument('input_root', type=str, help='Path to the root directory under which inputs are located.')
	p.add_argument('annotation_file', type=str, help='Path to the annotation csv file.')
	p.add_argument('--annotation_sep', type=str, default=',', help='Separator symbol of the annotation file.')
	p.add_argument('-S', '--save_root', type=str, default=None, help='Path to the directory where results are saved.')
	p.add_argument('-j', '--job_id', type=str, default='NO_JOB_ID', help='Job ID.')
	p.add_argument('-s', '--seed', type=int, default=1111, help='random seed')
	p.add_argument('-d', '--device', type=str, default='cuda', help='Computing device (CUDA only).')
	p.add_argument('-e', '--epochs', type=int, default=20, help='# of epochs to train the model.')
	p.add_argument('-b', '--batch_size', type=int, default=512, help='Batch size for training.')
	p.add_argument('--validation_batch_size', type=int, default=None, help='Batch size for validation (default: same).')
	p.add_argument('-l', '--learning_rate', type=float, default=1.0, help='Initial learning rate.')
	p.add_argument('-M', '--momentum', type=float, default=0.0, help='Momentum for SGD.')
	p.add_argument('-c', '--clip', type=float, default=1.0, help='Gradient clipping.')
	p.add_argument('-p', '--patience', type=int, default=0, help='# of epochs before updating the learning rate.')
	p.add_argument('-R', '--encoder_rnn_type', type=str, default='LSTM', help='LSTM or GRU.')
	p.add_argument('--decoder_rnn_type', type=str, default=None, help='Decoder RNN (default: same as the encoder).')
	p.add_argument('-f', '--feature_size', type=int, default=16, help='# of dimensions of features into which data are encoded.')
	p.add_argument('--encoder_rnn_layers', type=int, default=1)
	p.add_argument('--encoder_rnn_hidden_size', type=int, default=256)
	p.add_argument('--decoder_rnn_hidden_size', type=int, default=256)
	p.add_argument('--mlp_hidden_size', type=int, default=256)
	p.add_argument('--speaker_embed_dim', type=int, default=None)
	p.add_argument('--encoder_hidden_dropout', type=float, default=0.0)
	p.add_argument('--decoder_input_dropout', type=float, default=0.0)
	p.add_argument('--greedy_decoder', action='store_true')
	p.add_argument('--esn_leak', type=float, default=1.0)
	p.add_argument('--unidirectional_encoder', action='store_true')
	p.add_argument('--bidirectional_decoder', action='store_true')
	p.add_argument('--right2left_decoder_weight', type=float, default=0.5)
	p.add_argument('--fft_frame_length', type=float, default=0.008, help='FFT frame length in sec.')
	p.add_argument('--fft_step_size', type=float, default=0.004, help='FFT step size in sec.')
	p.add_argument('--fft_window_type', type=str, default='hann_window')
	p.add_argument('--fft_no_centering', action='store_true')
	p.add_argument('--channel', type=int, default=0)
	p.add_argument('-N', '--data_normalizer', type=float, default=1.0, help='Normalizing constant to divide the data.')
	p.add_argument('-E', '--epsilon', type=float, default=2 ** (-15), help='Small positive number added before log.')
	return p.parse_args(argv)


def main(argv=None):
	a = get_parameters(argv)
	save_dir = get_save_dir(a.save_root if a.save_root is not None else a.input_root, a.job_id)
	parser = Data_Parser(a.input_root, a.annotation_file, annotation_sep=a.annotation_sep)
	fs = parser.get_sample_freq()
	num_speakers = parser.get_num_speakers()
	fft_frame_length = int(np.floor(a.fft_frame_length * fs))
	fft_step_size = int(np.floor(a.fft_step_size * fs))
	if a.decoder_rnn_type is None:
		a.decoder_rnn_type = a.encoder_rnn_type
	learner = Learner(
		int(fft_frame_length / 2 + 1), a.encoder_rnn_hidden_size, a.decoder_rnn_hidden_size, a.mlp_hidden_size, a.feature_size,
		save_dir, encoder_rnn_type=a.encoder_rnn_type, decoder_rnn_type=a.decoder_rnn_type,
		encoder_rnn_layers=a.encoder_rnn_layers, encoder_hidden_dropout=a.encoder_hidden_dropout,
		decoder_input_dropout=a.decoder_input_dropout, device=a.device, seed=a.seed,
		decoder_self_feedback=not a.greedy_decoder, bidirectional_encoder=not a.unidirectional_encoder,
		bidirectional_decoder=a.bidirectional_decoder, right2left_decoder_weight=a.right2left_decoder_weight,
		num_speakers=num_speakers, speaker_embed_dim=a.speaker_embed_dim)
	stft = data_utils.STFT(fft_frame_length, fft_step_size, window=a.fft_window_type, centering=not a.fft_no_centering, device=a.device)
	transform = data_utils.LogSTFT(stft, epsilon=a.epsilon, normalizer=a.data_normalizer)
	logger.info("log(abs(STFT(wav))) + {eps}) / {normalizer} will be the input.".format(eps=a.epsilon, normalizer=a.data_normalizer))
	logger.info("Sampling frequency of data: {fs}".format(fs=fs))
	logger.info("STFT window type: {fft_window}".format(fft_window=a.fft_window_type))
	logger.info("STFT frame lengths: {fft_frame_length_in_sec} sec".format(fft_frame_length_in_sec=a.fft_frame_length))
	logger.info("STFT step size: {fft_step_size_in_sec} sec".format(fft_step_size_in_sec=a.fft_step_size))
	train_dataset = parser.get_data(data_type='train', transform=transform, channel=a.channel, device=a.device)
	valid_dataset = parser.get_data(data_type='valid', transform=transform, channel=a.channel, device=a.device)
	if a.validation_batch_size is None:
		a.validation_batch_size = a.batch_size
	learner.learn(train_dataset, valid_dataset, a.epochs, a.batch_size, a.validation_batch_size, learning_rate=a.learning_rate,
				momentum=a.momentum, gradient_clip=a.clip, patience=a.patience)
	return learner


if __name__ == '__main__':
	main()
