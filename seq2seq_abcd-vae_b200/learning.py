# coding: utf-8
"""Training driver for the B200 hot path: `Learner` with the reference's constructor / learn() / checkpoint
surface and CLI flags (/root/reference ABCD-VAE/learning.py:36-347, 351-394), re-organised around one fused device
step:

    zero-copy flat parameter + gradient arenas  ->  encoder -> ABCD sampler -> decoder kernels (autograd Functions
    over the C ABI)  ->  [NCCL all-reduce of the gradient arena when world_size > 1]  ->  one fused
    global-norm-clip + SGD kernel.

Per-batch scalars stay on the device; the log line's three perplexities come from one stats kernel and are read
back once per batch (the reference issues >= 7 blocking .item() calls per batch, learning.py:167-186).
"""
import argparse
import itertools
import json
import math
import os
from logging import DEBUG, FileHandler, Formatter, NullHandler, getLogger

import numpy as np
import torch

from . import _lib as L
from .modules import data_utils, model

logger = getLogger(__name__)


def _dist_rank_world(group=None):
	if torch.distributed.is_available() and torch.distributed.is_initialized():
		return torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
	return 0, 1


def update_log_handler(file_dir, write=True):
	"""(Re)attach <file_dir>/history.log; an existing log means the job is being resumed.  With write=False (ranks > 0 of
	a data-parallel job) the decision is the same but nothing is written: rank 0 owns history.log."""
	for h in list(logger.handlers):
		logger.removeHandler(h)
	log_file_path = os.path.join(file_dir, 'history.log')
	retrieval = os.path.isfile(log_file_path)
	if not write:
		logger.addHandler(NullHandler())
		logger.setLevel(DEBUG)
		return retrieval, log_file_path
	handler = FileHandler(filename=log_file_path)
	handler.setLevel(DEBUG)
	handler.setFormatter(Formatter('{asctime} - {levelname} - {message}', style='{'))
	logger.setLevel(DEBUG)
	logger.addHandler(handler)
	if retrieval:
		logger.info("LEARNING RETRIEVED.")
	else:
		logger.info("Logger set up.")
		logger.info("PyTorch ver.: {ver}".format(ver=torch.__version__))
	return retrieval, log_file_path


class FlatArena(object):
	"""Re-homes the parameters of several modules into ONE contiguous fp32 buffer (and their .grad into a second
	one with the same layout) so that gradient all-reduce, global-norm clipping and the SGD update are each a single
	pass over flat memory.  Every tensor starts on a 16-byte boundary."""

	def __init__(self, modules, device):
		self.params = list(itertools.chain.from_iterable(m.parameters() for m in modules))
		self.offsets = []
		off = 0
		for p in self.params:
			assert p.dtype == torch.float32
			self.offsets.append(off)
			off += (p.numel() + 3) // 4 * 4
		self.numel = off
		self.flat = torch.zeros(off, dtype=torch.float32, device=device)
		self.grad = torch.zeros(off, dtype=torch.float32, device=device)
		for p, o in zip(self.params, self.offsets):
			n = p.numel()
			self.flat[o:o + n].copy_(p.data.reshape(-1))
			p.data = self.flat[o:o + n].view(p.shape)
			p.grad = self.grad[o:o + n].view(p.shape)
		self.momentum_buf = None
		self.partial = torch.empty(L.load().abcd_optim_partial_count(), dtype=torch.float32, device=device)
		self.norm = torch.zeros((), dtype=torch.float32, device=device)

	def rebind(self):
		"""Point .grad back into the arena (autograd may have replaced it, e.g. after set_to_none)."""
		for p, o in zip(self.params, self.offsets):
			if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
				p.grad = self.grad[o:o + p.numel()].view(p.shape)

	def clip_and_step(self, lr, momentum, max_norm):
		"""clip_grad_norm_(max_norm) + SGD.step() + zero_grad() in two kernel launches."""
		if momentum != 0.0 and self.momentum_buf is None:
			self.momentum_buf = torch.zeros_like(self.flat)
		L.call('abcd_gradnorm_clip_sgd', L.ptr(self.flat), L.ptr(self.grad), L.ptr(self.momentum_buf) if momentum != 0.0 else None,
			self.numel, float(lr), float(momentum), float(max_norm), 1, L.ptr(self.norm), L.ptr(self.partial), L.stream_ptr())
		return self.norm


class Learner(object):
	"""ABCD-VAE learner (reference: ABCD-VAE/learning.py:36-347).  The latent-variable part is factored into four small
	hooks (_build_sampler, _sampler_from_checkpoint, _latent, _log_batch) so that plain/learning.py can swap in the isotropic
	Gaussian latent of the reference's plain/ variant and share everything else."""
	has_pretraining = True

	def __init__(self, input_size, encoder_rnn_hidden_size, decoder_rnn_hidden_size, mlp_hidden_size, num_feature_categories,
				feature_dim, save_dir, encoder_rnn_type='LSTM', decoder_rnn_type='LSTM', encoder_rnn_layers=1,
				bidirectional_encoder=True, bidirectional_decoder=False, right2left_decoder_weight=0.5,
				encoder_hidden_dropout=0.0, decoder_input_dropout=0.0, device='cuda', seed=1111,
				emission_distribution='isotropic_gaussian', decoder_self_feedback=True, esn_leak=1.0, num_speakers=None,
				speaker_embed_dim=None, prior_concentration=1.0, process_group=None):
		self._init_common(save_dir, device, process_group)
		if self.retrieval:
			self.last_epoch = self.retrieve_model(device=device)
			logger.info('Model retrieved.')
			return
		self._seed_everything(seed)
		if encoder_hidden_dropout > 0.0 and encoder_rnn_layers == 1:
			logger.warning('Non-zero dropout cannot be used for the single-layer encoder RNN (because there is no non-top hidden layers).')
			encoder_hidden_dropout = 0.0
		# creation order encoder -> sampler -> decoder fixes the seeded initial weights (reference learning.py:90-92)
		self.encoder = model.RNN_Variational_Encoder(input_size, encoder_rnn_hidden_size, rnn_type=encoder_rnn_type,
													rnn_layers=encoder_rnn_layers, hidden_dropout=encoder_hidden_dropout,
													bidirectional=bidirectional_encoder, esn_leak=esn_leak)
		self.feature_sampler = self._build_sampler(self.encoder.hidden_size_total, mlp_hidden_size, num_feature_categories,
													feature_dim, prior_concentration)
		self.decoder = model.RNN_Variational_Decoder(input_size, decoder_rnn_hidden_size, mlp_hidden_size, feature_dim,
													emission_distr_name=emission_distribution, rnn_type=decoder_rnn_type,
													input_dropout=decoder_input_dropout, self_feedback=decoder_self_feedback,
													esn_leak=esn_leak, bidirectional=bidirectional_decoder,
													right2left_weight=right2left_decoder_weight, num_speakers=num_speakers,
													speaker_embed_dim=speaker_embed_dim)
		logger.info('Random seed: {seed}'.format(seed=seed))
		logger.info('Type of RNN used for the encoder: {rnn_type}'.format(rnn_type=encoder_rnn_type))
		logger.info('Type of RNN used for the decoder: {rnn_type}'.format(rnn_type=decoder_rnn_type))
		logger.info("# of hidden units in the encoder RNNs: {hs}".format(hs=encoder_rnn_hidden_size))
		logger.info("# of hidden units in the decoder RNNs: {hs}".format(hs=decoder_rnn_hidden_size))
		logger.info("# of hidden units in the MLPs: {hs}".format(hs=mlp_hidden_size))
		self._place_on_device()

	# -- latent-variable hooks (overridden by plain/learning.py) ------------------------------------------
	def _build_sampler(self, hidden_size_total, mlp_hidden_size, num_feature_categories, feature_dim, prior_concentration):
		logger.info('Data are encoded into one of {num_cat} possible {feature_dim}-dim feature vectors.'.format(num_cat=num_feature_categories, feature_dim=feature_dim))
		logger.info('Discrete categories are assumed to be distributed according to Categorical(pi), with Dirichlet({}) prior on pi.'.format(prior_concentration))
		return model.ABCDSampler(hidden_size_total, mlp_hidden_size, num_feature_categories, feature_dim, prior_concentration=prior_concentration)

	def _sampler_from_checkpoint(self, init_parameters):
		return model.ABCDSampler(**init_parameters)

	def _latent(self, last_hidden, num_strings, is_pretraining):
		"""-> (decoder input features, KL summed over the batch, category logits)  (learning.py:150-152)"""
		if self.feature_sampler.fused_supported():   # one fused kernel after the to_code_like MLP (codebook fits in shared memory)
			return self.feature_sampler.forward_sample_kl(last_hidden, num_strings, no_sample=is_pretraining)
		category_logits = self.feature_sampler(last_hidden)
		features = self.feature_sampler.sample(category_logits, no_sample=is_pretraining)
		kl_loss = self.feature_sampler.kl_divergence(category_logits, num_strings)
		return features, kl_loss, category_logits

	def _log_batch(self, batch_ix, num_batches, loss, logits):
		"""One read-back per batch: the loss and the three perplexities of learning.py:171-186 (with world_size > 1 the
		perplexities are those of rank 0's shard of the batch)."""
		stats = torch.cat([loss.detach().reshape(1), self.feature_sampler.perplexities(logits)]).tolist()
		logger.info('{batch_ix}/{num_batches} training batches complete. mean loss: {loss:5.4f}. Perplexity of the posterior clustering probs.: {cluster_perplex:5.4f}. Perplexity of the mean clustering probs. over minibatch: {batch_perplex:5.4f}. Perplexity of the posterior Dirichlet shape: {shape_perplex:5.4f}'.format(
			batch_ix=batch_ix, num_batches=num_batches, loss=stats[0], cluster_perplex=stats[1], batch_perplex=stats[2], shape_perplex=stats[3]))

	# -- set-up -----------------------------------------------------------------------------------------
	def _init_common(self, save_dir, device, process_group):
		self.save_dir = save_dir
		self.process_group = process_group
		self.rank, self.world_size = _dist_rank_world(process_group)
		if not str(device).startswith('cuda'):
			raise ValueError('the B200 hot path has no CPU fallback: pass -d cuda (got {!r})'.format(device))
		self.device = torch.device(device)
		self.retrieval, self.log_file_path = False, None
		if save_dir is not None:
			# every rank decides "resume or fresh start" on the same file-system state, THEN rank 0 creates history.log
			self.retrieval = os.path.isfile(os.path.join(save_dir, 'history.log'))
			if self.world_size > 1:
				torch.distributed.barrier(group=process_group)
			_, self.log_file_path = update_log_handler(save_dir, write=self.rank == 0)
		logger.info('Device: {device}'.format(device=device))

	def _seed_everything(self, seed):
		torch.manual_seed(seed)
		torch.cuda.manual_seed_all(seed)
		# in-kernel Philox noise: ranks of a data-parallel job draw from disjoint streams
		model.manual_noise_seed(seed, stream=self.rank)

	def _modules(self):
		return (self.encoder, self.feature_sampler, self.decoder)

	def _place_on_device(self):
		L.ensure_device(self.device)
		for m in self._modules():
			m.to(self.device)
		self.parameters = lambda: itertools.chain.from_iterable(m.parameters() for m in self._modules())
		self.arena = FlatArena(self._modules(), self.device)
		self._n_enc_params = len(list(self.encoder.parameters()))
		import collections
		self.use_cuda_graph = getattr(self, 'use_cuda_graph', False)   # opt-in: Learner.use_cuda_graph = True (bench.py, -G)
		self.max_cuda_graphs = 8
		self._side_stream = torch.cuda.Stream(device=self.device)
		self._graphs = collections.OrderedDict()
		self._graph_seen = {}
		self._replayed_launches = 0

	# -- one fused optimisation step ----------------------------------------------------------------------
	def forward_losses(self, packed_input, is_offset, speaker, num_strings, is_pretraining, hidden_grad_hook=None):
		"""Encoder -> sampler -> decoder; returns the three summed losses and the category logits (None for plain/).
		`hidden_grad_hook` (optional) is registered on the encoder output: autograd calls it once the sampler's and the
		decoder's backward have been enqueued and before the encoder's backward starts."""
		last_hidden = self.encoder(packed_input)
		if hidden_grad_hook is not None and last_hidden.requires_grad:
			last_hidden.register_hook(hidden_grad_hook)
		features, kl_loss, category_logits = self._latent(last_hidden, num_strings, is_pretraining)
		emission_loss, end_prediction_loss, _, _, _ = self.decoder(
			features, batch_sizes=packed_input.batch_sizes, speaker=speaker, ground_truth_out=packed_input.data,
			ground_truth_offset=is_offset.data, return_outputs=False)
		return emission_loss, end_prediction_loss, kl_loss, category_logits

	def train_step(self, packed_input, is_offset, speaker, num_strings, is_pretraining, global_batch_size=None):
		"""forward + backward + (all-reduce) + clip + SGD for one (local shard of a) batch, all on the device.
		Gradients are SUMS over segments divided by the GLOBAL batch size, so shards add up exactly
		(reference: loss / batch_sizes[0], learning.py:155-158).  With world_size > 1 the caller passes this rank's shard and
		`global_batch_size` (Learner.train does; a full batch on every rank would be summed world_size times).

		With `use_cuda_graph` the ~60 kernel launches of a step are recorded once per (input buffers, batch shape,
		hyper-parameters) into a CUDA graph -- the second time that combination is seen -- and replayed afterwards: one
		cudaGraphLaunch per step, kernels back to back, the weight-gradient branch running beside the encoder BPTT.  The noise
		stream position lives on the device (model.noise_like), so replays draw fresh noise exactly like eager steps."""
		if self.world_size > 1 and global_batch_size is None:
			raise ValueError('train_step under a process group needs global_batch_size (the batch must be sharded over the ranks)')
		args = (packed_input, is_offset, speaker, num_strings, is_pretraining, global_batch_size)
		key = self._graph_key(*args) if self.use_cuda_graph else None
		if key is None:
			out = self._train_step_body(*args)
		else:
			entry = self._graphs.get(key)
			if entry is None and self._graph_seen.get(key, 0) == 0:
				self._graph_seen[key] = 1
				if len(self._graph_seen) > 64:
					self._graph_seen.clear()
				out = self._train_step_body(*args)      # first sight: a host-launched step (also fills descriptor caches)
			else:
				if entry is None:
					entry = self._capture(key, args)
				else:
					self._graphs.move_to_end(key)
					self._replayed_launches += entry['launches']
					model._noise_state['offset'] += entry['noise']
				entry['graph'].replay()
				out = entry['out']
		if self.has_pretraining and not is_pretraining:
			self.feature_sampler.increment_iter_counts()
		return out

	def _train_step_body(self, packed_input, is_offset, speaker, num_strings, is_pretraining, global_batch_size):
		self.arena.rebind()
		pending = []
		hook = None
		side = self._side_stream
		if self.world_size > 1:
			# the arena is laid out [encoder | sampler | decoder]: the sampler/decoder part of the gradient is final before the
			# encoder's backward starts, so its all-reduce (5.8 of the 6.9 MB) overlaps the encoder BPTT and weight-gradient GEMMs
			tail = self.arena.grad[self.arena.offsets[self._n_enc_params]:]

			def hook(grad):
				# sampler gradients are enqueued on this stream, decoder weight gradients on the side stream: reduce after both
				side.wait_stream(torch.cuda.current_stream())
				with torch.cuda.stream(side):
					pending.append(torch.distributed.all_reduce(tail, op=torch.distributed.ReduceOp.SUM, group=self.process_group, async_op=True))
				return grad
		nvtx = torch.cuda.nvtx
		nvtx.range_push('abcd/forward')
		emission_loss, end_loss, kl_loss, logits = self.forward_losses(packed_input, is_offset, speaker, num_strings, is_pretraining,
																		hidden_grad_hook=hook)
		batch_size = int(packed_input.batch_sizes[0]) if global_batch_size is None else global_batch_size
		loss = (emission_loss + end_loss + kl_loss) / batch_size
		nvtx.range_pop()
		nvtx.range_push('abcd/backward')
		# kernels add straight into the flat gradient arena; the decoder's weight-gradient GEMMs run on a second stream beside
		# the sampler backward and the encoder BPTT (which leaves 36 of the 148 SMs idle)
		with model.accumulate_into_grad(), model.weight_gradient_stream(side):
			loss.backward()
		torch.cuda.current_stream().wait_stream(side)
		nvtx.range_pop()
		nvtx.range_push('abcd/allreduce+clip+sgd')
		if self.world_size > 1:
			if pending:
				head = self.arena.grad[:self.arena.offsets[self._n_enc_params]]
				torch.distributed.all_reduce(head, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
				with torch.cuda.stream(side):
					for w in pending:
						w.wait()
				torch.cuda.current_stream().wait_stream(side)
			else:   # the hook did not fire (encoder frozen): one all-reduce over everything
				torch.distributed.all_reduce(self.arena.grad, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
		lr = self.optimizer.param_groups[0]['lr']
		momentum = self.optimizer.param_groups[0]['momentum']
		self.arena.clip_and_step(lr, momentum, self.gradient_clip)
		model.noise_commit(self.device)
		nvtx.range_pop()
		return loss, emission_loss, end_loss, kl_loss, logits

	# -- CUDA-graph replay of a recorded step ---------------------------------------------------------------
	def _graph_key(self, packed_input, is_offset, speaker, num_strings, is_pretraining, global_batch_size):
		"""Everything a recorded step bakes in: buffer addresses, the batch layout, and every scalar passed by value."""
		smp, dec = self.feature_sampler, self.decoder
		if getattr(smp, 'injected_gumbel_noise', None) is not None or getattr(smp, 'injected_noise', None) is not None \
				or dec.injected_emission_noise is not None or not torch.is_grad_enabled():
			return None   # parity hooks hand in fresh tensors every step: host-launched
		bs = packed_input.batch_sizes
		g = self.optimizer.param_groups[0]
		return (packed_input.data.data_ptr(), is_offset.data.data_ptr(), speaker.data_ptr() if speaker is not None else 0,
				bs.numel(), bs.numpy().tobytes(), float(num_strings), bool(is_pretraining), global_batch_size,
				float(g['lr']), float(g['momentum']), float(self.gradient_clip), float(getattr(smp, 'temperature', 0.0)),
				model._noise_key(), self.arena.momentum_buf is None)

	def _capture(self, key, args):
		lib = L.load()
		torch.cuda.synchronize(self.device)
		if self.arena.momentum_buf is None and self.optimizer.param_groups[0]['momentum'] != 0.0:
			self.arena.momentum_buf = torch.zeros_like(self.arena.flat)
		noise0 = model._noise_state['offset']
		launches0 = lib.abcd_launch_count()
		graph = torch.cuda.CUDAGraph()
		with torch.cuda.graph(graph):
			out = self._train_step_body(*args)
		entry = {'graph': graph, 'out': out, 'launches': int(lib.abcd_launch_count() - launches0),
				'noise': model._noise_state['offset'] - noise0}
		self._graphs[key] = entry
		while len(self._graphs) > self.max_cuda_graphs:
			self._graphs.popitem(last=False)
		return entry

	def release_graphs(self):
		"""Drop the recorded steps.  Call before torch.distributed.destroy_process_group(): a live CUDA graph that contains
		NCCL kernels keeps the communicator busy and the tear-down blocks."""
		torch.cuda.synchronize(self.device)
		self._graphs.clear()
		self._graph_seen.clear()
		import gc
		gc.collect()
		torch.cuda.synchronize(self.device)

	def kernel_launches(self):
		"""Kernels of this library enqueued so far, including those replayed from recorded graphs."""
		return int(L.load().abcd_launch_count()) + self._replayed_launches

	def launch_mode(self):
		return 'cuda-graph replay ({} recorded step(s))'.format(len(self._graphs)) if self.use_cuda_graph and self._graphs else 'host-launched kernels'

	# -- epochs -----------------------------------------------------------------------------------------
	def _move(self, packed_input, is_offset, speaker):
		packed_input = packed_input.to(self.device)
		is_offset = is_offset.to(self.device)
		speaker = speaker.to(self.device) if self.decoder.embed_speaker is not None else None
		return packed_input, is_offset, speaker

	def _reduce_totals(self, totals):
		if self.world_size > 1:
			torch.distributed.all_reduce(totals, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
		return totals

	def train(self, dataloader, is_pretraining=False):
		"""One epoch.  Under a process group the DataLoader hands every rank its strided shard of each length-sorted batch
		(data_utils.DataLoader(rank, world_size)); losses are divided by the global batch size and summed over ranks."""
		for m in self._modules():
			m.train()
		num_batches = dataloader.get_num_batches()
		num_strings = len(dataloader.dataset)
		totals = torch.zeros(3, dtype=torch.float64, device=self.device)
		for batch_ix, (packed_input, is_offset, speaker, _) in enumerate(dataloader, 1):
			packed_input, is_offset, speaker = self._move(packed_input, is_offset, speaker)
			gb = dataloader.last_global_batch_size if self.world_size > 1 else None
			loss, em, end, kl, logits = self.train_step(packed_input, is_offset, speaker, num_strings, is_pretraining, global_batch_size=gb)
			part = torch.stack([em.detach(), end.detach(), kl.detach()]).double()
			totals += part / self.world_size if getattr(dataloader, 'last_replicated', False) else part
			if self.world_size > 1:
				loss = loss.detach().clone()
				torch.distributed.all_reduce(loss, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
			self._log_batch(batch_ix, num_batches, loss, logits)
		return self._log_epoch('training', self._reduce_totals(totals), num_strings)

	def _log_epoch(self, phase, totals, num_strings):
		emission_loss, end_prediction_loss, kl_loss = (totals / num_strings).tolist()
		mean_loss = emission_loss + end_prediction_loss + kl_loss
		logger.info('mean {} emission negative pdf loss (per string): {:5.4f}'.format(phase, emission_loss))
		logger.info('mean {} end-prediction loss (per string): {:5.4f}'.format(phase, end_prediction_loss))
		logger.info('mean {} KL (per string): {:5.4f}'.format(phase, kl_loss))
		logger.info('mean {} total loss (per string): {:5.4f}'.format(phase, mean_loss))
		if phase == 'training' and self.has_pretraining:
			self.feature_sampler.update_epoch_init_iter_counts()
		return mean_loss

	def test_or_validate(self, dataloader, is_pretraining=False):
		"""Same forward as training (it also samples, reference learning.py:224-226), no update."""
		for m in self._modules():
			m.eval()
		num_batches = dataloader.get_num_batches()
		num_strings = len(dataloader.dataset)
		totals = torch.zeros(3, dtype=torch.float64, device=self.device)
		with torch.no_grad():
			for batch_ix, (packed_input, is_offset, speaker, _) in enumerate(dataloader, 1):
				packed_input, is_offset, speaker = self._move(packed_input, is_offset, speaker)
				em, end, kl, _ = self.forward_losses(packed_input, is_offset, speaker, num_strings, is_pretraining)
				model.noise_commit(self.device)
				part = torch.stack([em, end, kl]).double()
				totals += part / self.world_size if getattr(dataloader, 'last_replicated', False) else part
				logger.info('{batch_ix}/{num_batches} validation batches complete.'.format(batch_ix=batch_ix, num_batches=num_batches))
		return self._log_epoch('validation', self._reduce_totals(totals), num_strings)

	def setup_optimizer(self, learning_rate, momentum, gradient_clip=None, patience=0):
		"""torch.optim.SGD / ReduceLROnPlateau are kept as host-side bookkeeping (lr, state_dict layout of the
		reference's checkpoints); the update itself is the fused kernel."""
		self.optimizer = torch.optim.SGD(self.parameters(), lr=learning_rate, momentum=momentum)
		if self.retrieval:
			self.optimizer.load_state_dict(self.checkpoint['optimizer'])
			self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer)
			self.lr_scheduler.load_state_dict(self.checkpoint['lr_scheduler'])
			self._momentum_from_optimizer_state()
		else:
			self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, patience=patience)
			self.gradient_clip = gradient_clip

	def _momentum_from_optimizer_state(self):
		"""Fill the flat momentum arena from torch.optim.SGD's per-parameter `momentum_buffer`s -- the form in which both the
		reference's checkpoints and ours carry momentum (learning.py:306-307)."""
		found = False
		for p, o in zip(self.arena.params, self.arena.offsets):
			buf = self.optimizer.state.get(p, {}).get('momentum_buffer')
			if buf is None:
				continue
			if self.arena.momentum_buf is None:
				self.arena.momentum_buf = torch.zeros_like(self.arena.flat)
			self.arena.momentum_buf[o:o + p.numel()].copy_(buf.reshape(-1).to(self.device))
			found = True
		self.optimizer.state.clear()   # the arena is the live copy from here on
		return found

	def _momentum_to_optimizer_state(self):
		"""Expose the momentum arena as per-parameter `momentum_buffer`s so that optimizer.state_dict() has the reference's
		layout (a reference Learner can resume from our checkpoint and vice versa)."""
		if self.arena.momentum_buf is None:
			return
		for p, o in zip(self.arena.params, self.arena.offsets):
			self.optimizer.state[p]['momentum_buffer'] = self.arena.momentum_buf[o:o + p.numel()].view(p.shape).clone()

	def _make_dataloaders(self, train_dataset, valid_dataset, batch_size_train, batch_size_valid):
		train_dataloader = data_utils.DataLoader(train_dataset, batch_size=batch_size_train, shuffle=True, device=self.device,
												rank=self.rank, world_size=self.world_size)
		valid_dataloader = data_utils.DataLoader(valid_dataset, batch_size=batch_size_valid, device=self.device,
												rank=self.rank, world_size=self.world_size)
		return train_dataloader, valid_dataloader

	def learn(self, train_dataset, valid_dataset, num_epochs, batch_size_train, batch_size_valid, pretrain_epochs=0,
			learning_rate=0.1, momentum=0.9, gradient_clip=0.25, patience=0):
		train_dataloader, valid_dataloader = self._make_dataloaders(train_dataset, valid_dataset, batch_size_train, batch_size_valid)
		self.setup_optimizer(learning_rate, momentum, gradient_clip, patience)
		if self.retrieval:
			initial_epoch = self.last_epoch + 1
			logger.info('To be restarted from the beginning of epoch #: {epoch}'.format(epoch=initial_epoch))
		else:
			initial_epoch = 1
			logger.info("START LEARNING.")
			logger.info("max # of epochs: {ep}".format(ep=num_epochs))
			if self.has_pretraining:
				logger.info("first {} epochs are for pretraining w/o gumbel-softmax sampling.".format(pretrain_epochs))
			logger.info("batch size for training data: {size}".format(size=batch_size_train))
			logger.info("batch size for validation data: {size}".format(size=batch_size_valid))
			logger.info("initial learning rate: {lr}".format(lr=learning_rate))
			logger.info("momentum for SGD: {momentum}".format(momentum=momentum))
			logger.info("gradient clipping: {gc}".format(gc=self.gradient_clip))
		for epoch in range(initial_epoch, num_epochs + 1):
			logger.info('START OF EPOCH: {:3d}'.format(epoch))
			logger.info('current learning rate: {lr}'.format(lr=self.optimizer.param_groups[0]['lr']))
			is_pretrain = self.has_pretraining and epoch <= pretrain_epochs
			logger.info('start of TRAINING phase.')
			self.train(train_dataloader, is_pretrain)
			logger.info('end of TRAINING phase.')
			logger.info('start of VALIDATION phase.')
			mean_valid_loss = self.test_or_validate(valid_dataloader, is_pretrain)
			logger.info('end of VALIDATION phase.')
			self.lr_scheduler.step(mean_valid_loss)
			if self.has_pretraining and epoch == pretrain_epochs:
				self.lr_scheduler.best = math.inf   # forget the pretraining optimum
				logger.info('END OF PRETRAINING.')
			self.save_model(epoch)
			logger.info('END OF EPOCH: {:3d}'.format(epoch))
		logger.info('END OF TRAINING')

	# -- checkpoints (same dict keys as the reference, learning.py:298-313) -------------------------------
	def save_model(self, epoch):
		"""checkpoint.pt with the reference's keys; momentum travels as optimizer.state[...]['momentum_buffer'] like torch's
		own SGD.  Under a process group only rank 0 writes (parameters are replicated) and everybody waits for the file."""
		if self.rank == 0:
			self._momentum_to_optimizer_state()
			checkpoint = {
				'epoch': epoch,
				'encoder': {k: v.detach().cpu().clone() for k, v in self.encoder.state_dict().items()},
				'encoder_init_parameters': self.encoder.pack_init_parameters(),
				'feature_sampler': {k: v.detach().cpu().clone() for k, v in self.feature_sampler.state_dict().items()},
				'feature_sampler_init_parameters': self.feature_sampler.pack_init_parameters(),
				'decoder': {k: v.detach().cpu().clone() for k, v in self.decoder.state_dict().items()},
				'decoder_init_parameters': self.decoder.pack_init_parameters(),
				'optimizer': self.optimizer.state_dict(),
				'lr_scheduler': self.lr_scheduler.state_dict(),
				'gradient_clip': self.gradient_clip,
				'random_state': torch.get_rng_state(),
				'random_state_cuda': torch.cuda.get_rng_state_all(),
				'noise_state': dict(model._noise_state),   # in-kernel Philox stream (not in the reference; ignored by it)
			}
			torch.save(checkpoint, os.path.join(self.save_dir, 'checkpoint.pt'))
			self.optimizer.state.clear()
			logger.info('Config successfully saved.')
		if self.world_size > 1:
			torch.distributed.barrier(group=self.process_group)

	def retrieve_model(self, checkpoint_path=None, device='cuda'):
		"""Loads checkpoints written by this Learner or by the reference's (learning.py:315-347): same keys; a reference
		checkpoint simply lacks 'noise_state'."""
		if checkpoint_path is None:
			checkpoint_path = os.path.join(self.save_dir, 'checkpoint.pt')
		self.checkpoint = torch.load(checkpoint_path, map_location='cpu', weights_only=False)
		self.encoder = model.RNN_Variational_Encoder(**self.checkpoint['encoder_init_parameters'])
		self.feature_sampler = self._sampler_from_checkpoint(dict(self.checkpoint['feature_sampler_init_parameters']))
		self.decoder = model.RNN_Variational_Decoder(**dict(self.checkpoint['decoder_init_parameters']))
		self.encoder.load_state_dict(self.checkpoint['encoder'], strict=False)
		self.feature_sampler.load_state_dict(self.checkpoint['feature_sampler'])
		self.decoder.load_state_dict(self.checkpoint['decoder'])
		self._place_on_device()
		self.gradient_clip = self.checkpoint['gradient_clip']
		try:
			torch.set_rng_state(self.checkpoint['random_state'])
		except RuntimeError:
			logger.warning('Failed to retrieve random_state.')
		cuda_states = self.checkpoint.get('random_state_cuda')
		if cuda_states is not None and len(cuda_states) == torch.cuda.device_count():
			torch.cuda.set_rng_state_all(cuda_states)   # (unused by the kernels: their noise is the Philox stream below)
		if 'noise_state' in self.checkpoint:
			ns = self.checkpoint['noise_state']
			model._noise_state['seed'] = ns['seed']
			model._noise_state['stream'] = getattr(self, 'rank', 0)
			model.set_noise_offset(ns['offset'])
		return self.checkpoint['epoch']


def get_parameters(argv=None):
	p = argparse.ArgumentParser(description='Train the seq2seq ABCD-VAE on a B200.')
	p.add_argument('input_root', type=str, help='Path to the root directory under which inputs are located.')
	p.add_argument('annotation_file', type=str, help='Path to the annotation csv file.')
	p.add_argument('--annotation_sep', type=str, default=',', help='Separator symbol of the annotation file.')
	p.add_argument('-S', '--save_root', type=str, default=None, help='Path to the directory where results are saved.')
	p.add_argument('-j', '--job_id', type=str, default='NO_JOB_ID', help='Job ID.')
	p.add_argument('-s', '--seed', type=int, default=1111, help='random seed')
	p.add_argument('-d', '--device', type=str, default='cuda', help='Computing device (CUDA only).')
	p.add_argument('-e', '--epochs', type=int, default=20, help='# of epochs to train the model.')
	p.add_argument('--pretrain_epochs', type=int, default=5, help='# of initial epochs w/o gumbel-softmax sampling.')
	p.add_argument('-b', '--batch_size', type=int, default=512, help='Batch size for training.')
	p.add_argument('--validation_batch_size', type=int, default=None, help='Batch size for validation (default: same).')
	p.add_argument('-l', '--learning_rate', type=float, default=1.0, help='Initial learning rate.')
	p.add_argument('-M', '--momentum', type=float, default=0.0, help='Momentum for SGD.')
	p.add_argument('-c', '--clip', type=float, default=1.0, help='Gradient clipping.')
	p.add_argument('-p', '--patience', type=int, default=0, help='# of epochs before updating the learning rate.')
	p.add_argument('-R', '--encoder_rnn_type', type=str, default='LSTM', help='LSTM or GRU.')
	p.add_argument('--decoder_rnn_type', type=str, default=None, help='Decoder RNN (default: same as the encoder).')
	p.add_argument('-K', '--num_feature_categories', type=int, default=128, help='# of discrete categories.')
	p.add_argument('-f', '--feature_dim', type=int, default=256, help='# of dimensions of the codebook features.')
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
	p.add_argument('--prior_concentration', type=float, default=1.0)
	return p.parse_args(argv)


def get_save_dir(save_root, job_id_str):
	save_dir = os.path.join(save_root, job_id_str)
	os.makedirs(save_dir, exist_ok=True)
	return save_dir


def main(argv=None):
	a = get_parameters(argv)
	save_dir = get_save_dir(a.save_root if a.save_root is not None else a.input_root, a.job_id)
	speaker_coding_path = os.path.join(save_dir, 'speaker_coding.json')
	speaker2ix = None
	if os.path.isfile(speaker_coding_path):
		with open(speaker_coding_path, 'r') as f:
			speaker2ix = json.load(f)
	parser = data_utils.Data_Parser(a.input_root, a.annotation_file, annotation_sep=a.annotation_sep, speaker2ix=speaker2ix)
	fs = parser.get_sample_freq()
	num_speakers = parser.get_num_speakers()
	if num_speakers > 0 and speaker2ix is None:
		with open(speaker_coding_path, 'w') as f:
			json.dump(parser.speaker2ix, f)
	fft_frame_length = int(np.floor(a.fft_frame_length * fs))
	fft_step_size = int(np.floor(a.fft_step_size * fs))
	if a.decoder_rnn_type is None:
		a.decoder_rnn_type = a.encoder_rnn_type
	learner = Learner(
		int(fft_frame_length / 2 + 1), a.encoder_rnn_hidden_size, a.decoder_rnn_hidden_size, a.mlp_hidden_size,
		a.num_feature_categories, a.feature_dim, save_dir, encoder_rnn_type=a.encoder_rnn_type,
		decoder_rnn_type=a.decoder_rnn_type, encoder_rnn_layers=a.encoder_rnn_layers,
		encoder_hidden_dropout=a.encoder_hidden_dropout, decoder_input_dropout=a.decoder_input_dropout, device=a.device,
		seed=a.seed, decoder_self_feedback=not a.greedy_decoder, bidirectional_encoder=not a.unidirectional_encoder,
		bidirectional_decoder=a.bidirectional_decoder, right2left_decoder_weight=a.right2left_decoder_weight,
		num_speakers=num_speakers, speaker_embed_dim=a.speaker_embed_dim, prior_concentration=a.prior_concentration)
	stft = data_utils.STFT(fft_frame_length, fft_step_size, window=a.fft_window_type, centering=not a.fft_no_centering, device=a.device)
	transform = data_utils.LogSTFT(stft, epsilon=a.epsilon, normalizer=a.data_normalizer)
	logger.info("log(abs(STFT(wav))) + {eps}) / {normalizer} will be the input.".format(eps=a.epsilon, normalizer=a.data_normalizer))
	logger.info("Sampling frequency of data: {fs}".format(fs=fs))
	train_dataset = parser.get_data(data_type='train', transform=transform, channel=a.channel, device=a.device)
	valid_dataset = parser.get_data(data_type='valid', transform=transform, channel=a.channel, device=a.device)
	if a.validation_batch_size is None:
		a.validation_batch_size = a.batch_size
	learner.learn(train_dataset, valid_dataset, a.epochs, a.batch_size, a.validation_batch_size,
				pretrain_epochs=a.pretrain_epochs, learning_rate=a.learning_rate, momentum=a.momentum,
				gradient_clip=a.clip, patience=a.patience)
	return learner


if __name__ == '__main__':
	main()
