# coding: utf-8
"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/abcd_b200.h declares; host-side
logic (packed descriptors, batching arithmetic).  No compute calls -- there is no GPU here."""
import os
import re

import numpy as np
import pytest
import torch


def test_library_exports_every_declared_symbol(abcd):
	lib = abcd._lib.load()
	header = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'include', 'abcd_b200.h')).read()
	declared = set(re.findall(r'\b(abcd_[a-z0-9_]+)\s*\(', header))
	assert declared, 'no declarations parsed'
	bound = set(abcd._lib.SIGNATURES)
	assert declared == bound, (declared ^ bound)
	for name in declared:
		assert hasattr(lib, name), name
	assert b'sm_100a' in lib.abcd_version()


def test_workspace_size_queries_are_pure_host_calls(abcd):
	lib = abcd._lib.load()
	small = lib.abcd_enc_workspace_bytes(0, 65, 8, 100, 1, 0)
	train = lib.abcd_enc_workspace_bytes(0, 65, 8, 100, 1, 1)
	assert 0 < small < train
	assert lib.abcd_dec_workspace_bytes(0, 65, 256, 8, 100) > 0
	assert lib.abcd_sampler_kl_ws_floats(128) >= 3 * 128
	assert lib.abcd_optim_partial_count() > 0


def test_missing_gpu_fails_loudly(abcd):
	if torch.cuda.is_available():
		pytest.skip('GPU present')
	enc = abcd.model.RNN_Variational_Encoder(65, 256)
	x = torch.nn.utils.rnn.pack_sequence([torch.randn(4, 65), torch.randn(2, 65)])
	with pytest.raises((abcd._lib.AbcdError, AssertionError, RuntimeError)):
		enc(x)


def test_pack_time_major_matches_torch(abcd):
	lengths = [9, 9, 4, 4, 4, 1]
	bs = abcd.data_utils.pack_time_major(lengths)
	ref = torch.nn.utils.rnn.pack_sequence([torch.zeros(l) for l in lengths]).batch_sizes
	assert torch.equal(bs, ref)


def test_state_dict_keys_match_reference_names(abcd):
	enc = abcd.model.RNN_Variational_Encoder(65, 256)
	assert set(enc.state_dict()) == {'rnn.' + n + s for n in ('weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0') for s in ('', '_reverse')}
	assert enc.hidden_size_total == 1024
	smp = abcd.model.ABCDSampler(1024, 256, 128, 256)
	assert set(smp.state_dict()) == {'prior_concentration', 'posterior_shape_logits', 'codebook',
									'to_code_like.whole_network.0.weight', 'to_code_like.whole_network.0.bias',
									'to_code_like.whole_network.2.weight', 'to_code_like.whole_network.2.bias'}
	dec = abcd.model.RNN_Variational_Decoder(65, 256, 256, 256, num_speakers=3, speaker_embed_dim=8)
	keys = set(dec.state_dict())
	for k in ('feature2hidden.weight', 'offset_predictor.whole_network.2.bias', 'rnn_cell.cell.weight_ih',
			'emission_sampler.to_parameters.mlps.1.whole_network.0.weight', 'embed_speaker.weight'):
		assert k in keys
	assert dec.feature2hidden.weight.shape == (512, 264)
	assert dec.pack_init_parameters()['rnn_type'] == 'LSTM'


def test_temperature_schedule(abcd):
	smp = abcd.model.ABCDSampler(8, 8, 4, 8, temperature_update_freq=10, temperature_anneal_rate=0.1)
	assert smp.temperature == 1.0
	for _ in range(10):
		smp.increment_iter_counts()
	assert abs(smp.temperature - np.exp(-1.0)) < 1e-12
	smp.update_epoch_init_iter_counts()
	assert smp.pack_init_parameters()['epoch_init_iter_counts'] == 10
