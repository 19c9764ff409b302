# coding: utf-8
"""nn.Module surface of the ABCD-VAE hot path, backed by libabcd_b200.so.

Class names, constructor signatures, forward signatures, attribute names and state_dict keys follow
/root/reference ABCD-VAE/modules/model.py (and plain/modules/model.py for the Gaussian latent) so that these
classes drop in for the reference's; the arithmetic is done by the sm_100a kernels behind include/abcd_b200.h.
torch.nn layers appear only as *parameter containers* (same names, shapes and initialisation order as the
reference, hence identical weights under the same seed); their forward methods are never called.

Out of scope: ESN cells, the bidirectional decoder, decoder input dropout /
greedy decoding and multi-layer encoders raise NotImplementedError instead of silently running something else.
"""
import collections
import ctypes as C
import math

import torch

from .. import _lib as L


def _f32(t):
	assert t.dtype == torch.float32, 'abcd_b200 kernels are fp32'
	return t.contiguous()


# Every backward entry point of the library ACCUMULATES parameter gradients into the buffers it is given.  Inside
# `accumulate_into_grad()` (learning.Learner.train_step wraps loss.backward() in it) the buffers are the leaf
# parameters' existing .grad tensors -- exactly what autograd's AccumulateGrad would do with a returned gradient, minus
# one zero-fill and one add kernel per parameter.  Outside it (torch.autograd.grad, hooks, parameters without a .grad
# yet) fresh zero buffers are filled and returned to autograd as usual.
_DIRECT_GRADS = False


class accumulate_into_grad(object):
	def __enter__(self):
		global _DIRECT_GRADS
		self.prev = _DIRECT_GRADS
		_DIRECT_GRADS = True
		return self

	def __exit__(self, *exc):
		global _DIRECT_GRADS
		_DIRECT_GRADS = self.prev
		return False


# Optional second stream for the decoder's weight-gradient reductions (Learner.train_step): inside
# `weight_gradient_stream(side)` _DecoderFn.backward enqueues the data path (BPTT, d_features) on the current stream and the
# parameter gradients on `side`, ordered by an event; the caller joins `side` before it reads the gradients.
_WGRAD_STREAM = None


class weight_gradient_stream(object):
	def __init__(self, stream):
		self.stream = stream

	def __enter__(self):
		global _WGRAD_STREAM
		self.prev = _WGRAD_STREAM
		_WGRAD_STREAM = self.stream
		return self

	def __exit__(self, *exc):
		global _WGRAD_STREAM
		_WGRAD_STREAM = self.prev
		return False


def _grad_buffers(tensors):
	"""-> (buffers handed to the kernel, values returned to autograd) for parameter-like inputs."""
	bufs, rets = [], []
	for t in tensors:
		g = t.grad if (_DIRECT_GRADS and t.is_leaf and t.requires_grad) else None
		if (g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == t.shape and g.device == t.device
				and not t._backward_hooks):
			bufs.append(g)
			rets.append(None)
		else:
			z = torch.zeros_like(t)
			bufs.append(z)
			rets.append(z)
	return bufs, rets


def _cell_code(rnn_type):
	if rnn_type == 'LSTM':
		return 0
	if rnn_type == 'GRU':
		return 1
	raise NotImplementedError('rnn_type {!r} is outside the B200 hot path (LSTM and GRU only; ESN is out of scope)'.format(rnn_type))


def _ws(nbytes, device):
	return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


_DESC_CACHE_MAX = 8
_DESC_CACHE = collections.OrderedDict()


def as_packed_desc(batch_sizes, device, cache=None):
	"""Build (or fetch from `cache`) the device descriptor for a PackedSequence.batch_sizes tensor.  The cache is ONE small
	LRU shared by the encoder and the decoder (both see the same batch_sizes in a step), keyed by the raw bytes of the
	tensor: shuffled training batches almost never repeat a layout, so it holds a handful of entries, not one per step."""
	if isinstance(batch_sizes, L.Packed):
		return batch_sizes
	cache = _DESC_CACHE
	bs = torch.as_tensor(batch_sizes, dtype=torch.int64, device='cpu')
	key = (str(device), bs.numpy().tobytes())
	desc = cache.get(key)
	if desc is not None:
		cache.move_to_end(key)
		return desc
	desc = L.Packed(bs, device)
	cache[key] = desc
	while len(cache) > _DESC_CACHE_MAX:
		cache.popitem(last=False)
	return desc


# ------------------------------------------------------------------------------------------------------
# Gaussian helper functions (reference: model.py:17-37).  Kept as plain tensor expressions for API parity;
# the training path uses the fused kernels instead.
# ------------------------------------------------------------------------------------------------------
def sample_from_isotropic_gaussian(mean, log_variance):
	noise = noise_like(mean)
	return _GaussLatentFn.apply(mean, log_variance, noise)[0]


def kl_isotropic_to_standard_gaussian(mean, log_variance):
	return _GaussLatentFn.apply(mean, log_variance, None)[1]


def log_pdf_isotropic_gaussian(value, mean, log_variance):
	diff = value - mean
	return -0.5 * (math.log(2 * math.pi) + log_variance + diff * (-log_variance).exp() * diff).sum()


def choose_distribution(distribution_name):
	if distribution_name != 'isotropic_gaussian':
		raise KeyError(distribution_name)
	return (sample_from_isotropic_gaussian, log_pdf_isotropic_gaussian, kl_isotropic_to_standard_gaussian, 2)


# In-kernel Philox noise.  The stream position lives ON THE DEVICE (one uint64 per device): a launch draws at
# *counter + `_noise_rel`, where `_noise_rel` counts the draws already issued since the last commit, and noise_commit()
# advances the device counter by that amount with a one-thread kernel.  Eager steps and CUDA-graph replays of a recorded
# step therefore consume the same stream ('offset' is the host mirror of the committed position, kept for checkpoints).
_noise_state = {'seed': 0x5EED, 'offset': 0, 'stream': 0}
_noise_rel = 0
_noise_ctr = {}


def manual_noise_seed(seed, stream=0):
	"""Seed of the in-kernel Philox stream used for Gumbel / emission noise.  `stream` selects one of 2^64 disjoint key
	streams under the same seed: data-parallel ranks pass their rank, so different shards of a batch never see the same
	draws (the reference has one process and one torch generator)."""
	_noise_state['seed'] = int(seed) & 0xFFFFFFFFFFFFFFFF
	_noise_state['stream'] = int(stream)
	set_noise_offset(0)


def set_noise_offset(offset):
	"""Move the stream position (checkpoint restore)."""
	global _noise_rel
	_noise_state['offset'] = int(offset)
	_noise_rel = 0
	for t in _noise_ctr.values():
		t.fill_(int(offset))


def _noise_key():
	return (_noise_state['seed'] ^ ((_noise_state.get('stream', 0) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def _noise_counter(device):
	key = str(device)
	if key not in _noise_ctr:
		_noise_ctr[key] = torch.tensor([_noise_state['offset']], dtype=torch.int64, device=device)
	return _noise_ctr[key]


def noise_like(t, kind='normal'):
	"""Philox draws with the shape of `t` (N(0,1) or standard Gumbel)."""
	global _noise_rel
	L.ensure_device(t.device)
	out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
	n = out.numel()
	L.call('abcd_philox_fill', L.ptr(out), n, 0 if kind == 'normal' else 1, _noise_key(), _noise_rel, L.ptr(_noise_counter(t.device)), L.stream_ptr())
	_noise_rel += (n + 3) // 4
	return out


def noise_pending():
	"""Draw groups issued since the last commit."""
	return _noise_rel


def noise_commit(device, host_only=False):
	"""Advance the device-resident stream position past the draws issued since the last commit (enqueued on the current
	stream, after the kernels that consumed them).  host_only=True only updates the host mirror: used after replaying a
	CUDA graph that already contains the device-side increment."""
	global _noise_rel
	if _noise_rel == 0:
		return 0
	n = _noise_rel
	if not host_only:
		L.call('abcd_counter_add', L.ptr(_noise_counter(device)), n, L.stream_ptr())
	_noise_state['offset'] += n
	_noise_rel = 0
	return n


class _GaussLatentFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, mean, log_variance, eps):
		L.ensure_device(mean.device)
		mean, log_variance = _f32(mean), _f32(log_variance)
		has_eps = eps is not None
		z = torch.empty_like(mean) if has_eps else None
		kl = torch.empty((), dtype=torch.float32, device=mean.device)
		partial = torch.empty(160, dtype=torch.float32, device=mean.device)
		L.call('abcd_gauss_latent_fwd', mean.numel(), L.ptr(mean), L.ptr(log_variance), L.ptr(eps) if has_eps else None,
			L.ptr(z) if has_eps else None, L.ptr(kl), L.ptr(partial), L.stream_ptr())
		ctx.save_for_backward(mean, log_variance, eps if has_eps else mean)
		ctx.has_eps = has_eps
		if not has_eps:
			z = mean.new_zeros(())
		return z, kl

	@staticmethod
	def backward(ctx, dz, dkl):
		mean, lv, eps = ctx.saved_tensors
		dmean = torch.empty_like(mean)
		dlv = torch.empty_like(mean)
		has_dz = ctx.has_eps and dz is not None
		dz_c = _f32(dz) if has_dz else None
		dkl_c = _f32(dkl) if dkl is not None else None
		L.call('abcd_gauss_latent_bwd', mean.numel(), L.ptr(mean), L.ptr(lv), L.ptr(eps) if ctx.has_eps else L.ptr(mean),
			L.ptr(dz_c), L.ptr(dkl_c), L.ptr(dmean), L.ptr(dlv), L.stream_ptr())
		return dmean, dlv, None


# ------------------------------------------------------------------------------------------------------
# MLP (reference: model.py:303-334)
# ------------------------------------------------------------------------------------------------------
def _mlp_structs(net, grads=None):
	w1, b1, w2, b2 = net[0].weight, net[0].bias, net[2].weight, net[2].bias
	p = L.MlpParamsT(w1.shape[1], w1.shape[0], w2.shape[0], L.ptr(w1.data), L.ptr(b1.data), L.ptr(w2.data), L.ptr(b2.data))
	g = None
	if grads is not None:
		g = L.MlpGradsT(*[L.ptr(t) for t in grads])
	return p, g


class _MlpFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, x, w1, b1, w2, b2):
		L.ensure_device(x.device)
		x = _f32(x)
		rows = x.shape[0]
		hid = torch.empty(rows, w1.shape[0], dtype=torch.float32, device=x.device)
		y = torch.empty(rows, w2.shape[0], dtype=torch.float32, device=x.device)
		p = L.MlpParamsT(w1.shape[1], w1.shape[0], w2.shape[0], L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2))
		L.call('abcd_mlp_fwd', C.byref(p), rows, L.ptr(x), L.ptr(hid), L.ptr(y), L.stream_ptr())
		ctx.save_for_backward(x, hid, w1, b1, w2, b2)
		return y

	@staticmethod
	def backward(ctx, dy):
		x, hid, w1, b1, w2, b2 = ctx.saved_tensors
		dy = _f32(dy)
		rows = x.shape[0]
		grads, rets = _grad_buffers((w1, b1, w2, b2))
		dx = torch.empty_like(x)
		scratch = torch.empty(rows * w1.shape[0], dtype=torch.float32, device=x.device)
		p = L.MlpParamsT(w1.shape[1], w1.shape[0], w2.shape[0], L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2))
		g = L.MlpGradsT(*[L.ptr(t) for t in grads])
		L.call('abcd_mlp_bwd', C.byref(p), C.byref(g), rows, L.ptr(x), L.ptr(hid), L.ptr(dy), L.ptr(dx), L.ptr(scratch), L.stream_ptr())
		return (dx,) + tuple(rets)


class MLP(torch.nn.Module):
	"""Linear -> Tanh -> Linear; parameters live in `whole_network.{0,2}` like the reference's ScriptModule."""

	def __init__(self, input_size, hidden_size, output_size):
		super(MLP, self).__init__()
		self.input_size = input_size
		self.hidden_size = hidden_size
		self.output_size = output_size
		self.whole_network = torch.nn.Sequential(
			torch.nn.Linear(input_size, hidden_size),
			torch.nn.Tanh(),
			torch.nn.Linear(hidden_size, output_size))

	def forward(self, batched_input):
		net = self.whole_network
		return _MlpFn.apply(batched_input, net[0].weight, net[0].bias, net[2].weight, net[2].bias)


class MLP_To_k_Vecs(torch.nn.Module):
	def __init__(self, input_size, hidden_size, output_size, k):
		super(MLP_To_k_Vecs, self).__init__()
		self.input_size = input_size
		self.hidden_size = hidden_size
		self.output_size = output_size
		self.k = k
		self.mlps = torch.nn.ModuleList([MLP(input_size, hidden_size, output_size) for _ in range(k)])

	def forward(self, batched_input):
		return [mlp(batched_input) for mlp in self.mlps]


class Sampler(torch.nn.Module):
	"""Two parallel MLPs -> (mean, log-variance) of an isotropic Gaussian (reference: model.py:676-705)."""

	def __init__(self, input_size, mlp_hidden_size, output_size, distribution_name='isotropic_gaussian'):
		super(Sampler, self).__init__()
		self.distribution_name = distribution_name
		self._sampler, self._log_pdf, self._kl_divergence, num_parameters = choose_distribution(distribution_name)
		self.to_parameters = MLP_To_k_Vecs(input_size, mlp_hidden_size, output_size, num_parameters)

	def forward(self, parameter_seed):
		return self.to_parameters(parameter_seed)

	def sample(self, parameters):
		return self._sampler(*parameters)

	def kl_divergence(self, parameters):
		return self._kl_divergence(*parameters)

	def sample_and_kl(self, parameters):
		"""sample() and kl_divergence() of the same parameters in ONE fused kernel (plain/learning.py:148-149 calls them back
		to back): z = mean + exp(0.5 lv) eps and KL(N(mean, e^lv) || N(0, I)) summed over the batch.  `injected_noise`
		(parity hook) supplies eps for the next call; otherwise it is drawn from the Philox stream."""
		mean, log_variance = parameters
		eps = getattr(self, 'injected_noise', None)
		self.injected_noise = None
		if eps is None:
			eps = noise_like(mean)
		return _GaussLatentFn.apply(mean, log_variance, _f32(eps))

	def log_pdf(self, samples, parameters):
		return self._log_pdf(samples, *parameters)

	def pack_init_parameters(self):
		return {
			'input_size': self.to_parameters.input_size,
			'mlp_hidden_size': self.to_parameters.hidden_size,
			'output_size': self.to_parameters.output_size,
			'distribution_name': self.distribution_name,
		}


# ------------------------------------------------------------------------------------------------------
# Encoder (reference: model.py:40-79)
# ------------------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, x, desc, cell, bidir, hidden, *params):
		lib = L.ensure_device(x.device)
		x = _f32(x)
		ndirs = 2 if bidir else 1
		F = x.shape[1]
		params = [_f32(p) for p in params]
		need_grad = any(ctx.needs_input_grad[5:])
		ws_bytes = lib.abcd_enc_workspace_bytes(cell, F, desc.B, desc.sumT, int(bidir), int(need_grad))
		ws = _ws(ws_bytes, x.device)
		out = torch.empty(desc.B, ndirs * hidden * (2 if cell == 0 else 1), dtype=torch.float32, device=x.device)
		ep = L.EncParamsT()
		ep.cell, ep.F, ep.H, ep.bidir = cell, F, hidden, int(bidir)
		for d in range(ndirs):
			ep.w_ih[d], ep.w_hh[d], ep.b_ih[d], ep.b_hh[d] = [L.ptr(p) for p in params[4 * d:4 * d + 4]]
		L.call('abcd_enc_fwd', C.byref(ep), desc.ref(), L.ptr(x), x.shape[1], L.ptr(out), int(need_grad),
			L.ptr(ws), ws_bytes, L.stream_ptr())
		ctx.desc, ctx.cell, ctx.bidir, ctx.hidden, ctx.ws, ctx.ws_bytes = desc, cell, bidir, hidden, ws, ws_bytes
		ctx.save_for_backward(x, *params)
		return out

	@staticmethod
	def backward(ctx, d_out):
		x, *params = ctx.saved_tensors
		desc = ctx.desc
		ndirs = 2 if ctx.bidir else 1
		d_out = _f32(d_out)
		grads, rets = _grad_buffers(params)
		ep = L.EncParamsT()
		ep.cell, ep.F, ep.H, ep.bidir = ctx.cell, x.shape[1], ctx.hidden, int(ctx.bidir)
		eg = L.EncGradsT()
		for d in range(ndirs):
			ep.w_ih[d], ep.w_hh[d], ep.b_ih[d], ep.b_hh[d] = [L.ptr(p) for p in params[4 * d:4 * d + 4]]
			eg.w_ih[d], eg.w_hh[d], eg.b_ih[d], eg.b_hh[d] = [L.ptr(g) for g in grads[4 * d:4 * d + 4]]
		L.call('abcd_enc_bwd', C.byref(ep), C.byref(eg), desc.ref(), L.ptr(x), x.shape[1], L.ptr(d_out),
			L.ptr(ctx.ws), ctx.ws_bytes, L.stream_ptr())
		return (None, None, None, None, None) + tuple(rets)


class RNN_Variational_Encoder(torch.nn.Module):
	"""Bidirectional LSTM/GRU over a PackedSequence, returning the final states flattened per segment:
	LSTM [h_fwd | c_fwd | h_bwd | c_bwd], GRU [h_fwd | h_bwd]."""

	def __init__(self, input_size, rnn_hidden_size, rnn_type='LSTM', rnn_layers=1, hidden_dropout=0.0, bidirectional=True, esn_leak=1.0):
		super(RNN_Variational_Encoder, self).__init__()
		_cell_code(rnn_type)
		if rnn_layers != 1:
			raise NotImplementedError('multi-layer encoders are outside the B200 hot path (encoder_rnn_layers must be 1)')
		# parameter container only: same names/shapes/initialisation as the reference's torch.nn.LSTM/GRU
		self.rnn = getattr(torch.nn, rnn_type)(input_size, rnn_hidden_size, rnn_layers, dropout=hidden_dropout, bidirectional=bidirectional, batch_first=True)
		self.hidden_size_total = rnn_layers * rnn_hidden_size
		if bidirectional:
			self.hidden_size_total *= 2
		if rnn_type == 'LSTM':
			self.hidden_size_total *= 2
		self._desc_cache = {}

	def _params(self):
		names = ['weight_ih_l0', 'weight_hh_l0', 'bias_ih_l0', 'bias_hh_l0']
		ps = [getattr(self.rnn, n) for n in names]
		if self.rnn.bidirectional:
			ps += [getattr(self.rnn, n + '_reverse') for n in names]
		return ps

	def forward(self, packed_input):
		data, batch_sizes = packed_input.data, packed_input.batch_sizes
		desc = as_packed_desc(batch_sizes, data.device, self._desc_cache)
		cell = _cell_code(self.rnn.mode.split('_')[0])
		return _EncoderFn.apply(data, desc, cell, self.rnn.bidirectional, self.rnn.hidden_size, *self._params())

	def pack_init_parameters(self):
		return {
			'input_size': self.rnn.input_size,
			'rnn_hidden_size': self.rnn.hidden_size,
			'rnn_type': self.rnn.mode.split('_')[0],
			'rnn_layers': self.rnn.num_layers,
			'hidden_dropout': self.rnn.dropout,
			'bidirectional': self.rnn.bidirectional,
		}


# ------------------------------------------------------------------------------------------------------
# ABCD sampler (reference: model.py:538-673)
# ------------------------------------------------------------------------------------------------------
def _sampler_structs(w1, b1, w2, b2, codebook, shape_logits, prior_c, grads=None):
	sp = L.SamplerParamsT()
	sp.mlp = L.MlpParamsT(w1.shape[1], w1.shape[0], w2.shape[0], L.ptr(w1), L.ptr(b1), L.ptr(w2), L.ptr(b2))
	sp.f, sp.K = codebook.shape[0], codebook.shape[1]
	sp.codebook, sp.shape_logits, sp.prior_c = L.ptr(codebook), L.ptr(shape_logits), float(prior_c)
	sg = L.SamplerGradsT()
	if grads is not None:
		gw1, gb1, gw2, gb2, gcb, gsl = grads
		sg.mlp = L.MlpGradsT(L.ptr(gw1), L.ptr(gb1), L.ptr(gw2), L.ptr(gb2))
		sg.codebook, sg.shape_logits = L.ptr(gcb), L.ptr(gsl)
	return sp, sg


class _LogitsFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, x, w1, b1, w2, b2, codebook):
		L.ensure_device(x.device)
		x = _f32(x)
		B = x.shape[0]
		dev = x.device
		hid = torch.empty(B, w1.shape[0], dtype=torch.float32, device=dev)
		xc = torch.empty(B, w2.shape[0], dtype=torch.float32, device=dev)
		logits = torch.empty(B, codebook.shape[1], dtype=torch.float32, device=dev)
		sp, _ = _sampler_structs(w1, b1, w2, b2, codebook, None, 1.0)
		L.call('abcd_sampler_logits_fwd', C.byref(sp), B, L.ptr(x), L.ptr(hid), L.ptr(xc), L.ptr(logits), L.stream_ptr())
		ctx.save_for_backward(x, hid, xc, w1, b1, w2, b2, codebook)
		return logits

	@staticmethod
	def backward(ctx, dlogits):
		x, hid, xc, w1, b1, w2, b2, codebook = ctx.saved_tensors
		dlogits = _f32(dlogits)
		B = x.shape[0]
		grads, rets = _grad_buffers((w1, b1, w2, b2, codebook))
		dx = torch.empty_like(x)
		f, hidden = w2.shape[0], w1.shape[0]
		scratch = torch.empty(B * (f + max(hidden, f)), dtype=torch.float32, device=x.device)
		sp, sg = _sampler_structs(w1, b1, w2, b2, codebook, None, 1.0, grads + [None])
		L.call('abcd_sampler_logits_bwd', C.byref(sp), C.byref(sg), B, L.ptr(x), L.ptr(hid), L.ptr(xc), L.ptr(dlogits),
			L.ptr(dx), L.ptr(scratch), L.stream_ptr())
		return (dx,) + tuple(rets)


class _SampleFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, logits, codebook, mode, tau, noise):
		L.ensure_device(logits.device)
		logits = _f32(logits)
		B, K = logits.shape
		y = torch.empty_like(logits)
		feats = torch.empty(B, codebook.shape[0], dtype=torch.float32, device=logits.device)
		sp = L.SamplerParamsT()
		sp.f, sp.K, sp.codebook = codebook.shape[0], K, L.ptr(codebook)
		L.call('abcd_sampler_sample_fwd', C.byref(sp), B, L.ptr(logits), mode, float(tau), L.ptr(noise), L.ptr(y), L.ptr(feats), L.stream_ptr())
		ctx.save_for_backward(y, codebook)
		ctx.mode, ctx.tau = mode, float(tau)
		ctx.mark_non_differentiable(y)
		ctx.set_materialize_grads(False)
		return feats, y

	@staticmethod
	def backward(ctx, dfeat, _dy):
		y, codebook = ctx.saved_tensors
		if dfeat is None:
			return None, None, None, None, None
		dfeat = _f32(dfeat)
		B, K = y.shape
		(gcb,), (gcb_ret,) = _grad_buffers((codebook,))
		dlogits = torch.empty_like(y)
		scratch = torch.empty(B * K, dtype=torch.float32, device=y.device)
		sp = L.SamplerParamsT()
		sp.f, sp.K, sp.codebook = codebook.shape[0], K, L.ptr(codebook)
		sg = L.SamplerGradsT()
		sg.codebook = L.ptr(gcb)
		L.call('abcd_sampler_sample_bwd', C.byref(sp), C.byref(sg), B, L.ptr(y), ctx.mode, ctx.tau, L.ptr(dfeat), L.ptr(dlogits), L.ptr(scratch), L.stream_ptr())
		return dlogits, gcb_ret, None, None, None


class _KlFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, logits, shape_logits, prior_c, n_data):
		lib = L.ensure_device(logits.device)
		logits = _f32(logits)
		B, K = logits.shape
		ws = torch.empty(lib.abcd_sampler_kl_ws_floats(K) + 64, dtype=torch.float32, device=logits.device)
		kl = torch.empty((), dtype=torch.float32, device=logits.device)
		sp = L.SamplerParamsT()
		sp.K, sp.shape_logits, sp.prior_c = K, L.ptr(shape_logits), float(prior_c)
		L.call('abcd_sampler_kl_fwd', C.byref(sp), B, L.ptr(logits), float(n_data), L.ptr(kl), L.ptr(ws), L.stream_ptr())
		ctx.save_for_backward(logits, shape_logits, ws)
		ctx.prior_c, ctx.n_data = float(prior_c), float(n_data)
		return kl

	@staticmethod
	def backward(ctx, g_kl):
		logits, shape_logits, ws = ctx.saved_tensors
		g_kl = _f32(g_kl)
		B, K = logits.shape
		dlogits = torch.empty_like(logits)
		(dshape,), (dshape_ret,) = _grad_buffers((shape_logits,))
		sp = L.SamplerParamsT()
		sp.K, sp.shape_logits, sp.prior_c = K, L.ptr(shape_logits), ctx.prior_c
		sg = L.SamplerGradsT()
		sg.shape_logits = L.ptr(dshape)
		L.call('abcd_sampler_kl_bwd', C.byref(sp), C.byref(sg), B, L.ptr(logits), ctx.n_data, L.ptr(g_kl), L.ptr(dlogits), L.ptr(ws), L.stream_ptr())
		return dlogits, dshape_ret, None, None


class _SamplerFusedFn(torch.autograd.Function):
	"""logits / sample / features / KL of the ABCD sampler in one kernel each way (abcd_sampler_fused_fwd / _bwd)."""

	@staticmethod
	def forward(ctx, xc, codebook, shape_logits, mode, tau, gumbel, prior_c, n_data):
		global _noise_rel
		lib = L.ensure_device(xc.device)
		xc = _f32(xc)
		B, f = xc.shape
		K = codebook.shape[1]
		dev = xc.device
		logits = torch.empty(B, K, dtype=torch.float32, device=dev)
		y = torch.empty_like(logits)
		feats = torch.empty(B, f, dtype=torch.float32, device=dev)
		kl = torch.empty((), dtype=torch.float32, device=dev)
		ws = torch.empty(lib.abcd_sampler_fused_ws_floats(K) + 64, dtype=torch.float32, device=dev)
		sp = L.SamplerParamsT()
		sp.f, sp.K, sp.codebook, sp.shape_logits, sp.prior_c = f, K, L.ptr(codebook), L.ptr(shape_logits), float(prior_c)
		key, rel, ctr = 0, 0, None
		if mode and gumbel is None:   # in-kernel Philox draws: the same stream positions noise_like() would use for a [B, K] tensor
			key, rel, ctr = _noise_key(), _noise_rel, L.ptr(_noise_counter(dev))
			_noise_rel += (B * K + 3) // 4
		L.call('abcd_sampler_fused_fwd', C.byref(sp), B, L.ptr(xc), int(mode), float(tau), L.ptr(gumbel), key, rel, ctr, float(n_data),
			L.ptr(logits), L.ptr(y), L.ptr(feats), L.ptr(kl), L.ptr(ws), L.stream_ptr())
		ctx.save_for_backward(xc, logits, y, codebook, shape_logits)
		ctx.cfg = (int(mode), float(tau), float(prior_c), float(n_data))
		ctx.mark_non_differentiable(logits, y)
		ctx.set_materialize_grads(False)
		return feats, kl, logits, y

	@staticmethod
	def backward(ctx, d_feats, g_kl, _dl, _dy):
		xc, logits, y, codebook, shape_logits = ctx.saved_tensors
		mode, tau, prior_c, n_data = ctx.cfg
		lib = L.load()
		B, f = xc.shape
		K = codebook.shape[1]
		dev = xc.device
		d_feats = _f32(d_feats) if d_feats is not None else torch.zeros(B, f, dtype=torch.float32, device=dev)
		g_kl = _f32(g_kl) if g_kl is not None else torch.zeros((), dtype=torch.float32, device=dev)
		(gcb, gsl), (gcb_ret, gsl_ret) = _grad_buffers((codebook, shape_logits))
		d_xc = torch.empty_like(xc)
		scratch = torch.empty(B * K, dtype=torch.float32, device=dev)
		ws = torch.empty(lib.abcd_sampler_fused_ws_floats(K) + 64, dtype=torch.float32, device=dev)
		sp = L.SamplerParamsT()
		sp.f, sp.K, sp.codebook, sp.shape_logits, sp.prior_c = f, K, L.ptr(codebook), L.ptr(shape_logits), prior_c
		sg = L.SamplerGradsT()
		sg.codebook, sg.shape_logits = L.ptr(gcb), L.ptr(gsl)
		L.call('abcd_sampler_fused_bwd', C.byref(sp), C.byref(sg), B, L.ptr(xc), L.ptr(logits), L.ptr(y), mode, tau, n_data,
			L.ptr(d_feats), L.ptr(g_kl), L.ptr(d_xc), L.ptr(scratch), L.ptr(ws), L.stream_ptr())
		return d_xc, gcb_ret, gsl_ret, None, None, None, None, None


class ABCDSampler(torch.nn.Module):
	""" "A"ttention-"B"ased "C"ategorical sampler with the "D"irichlet prior."""

	def __init__(self, input_size, mlp_hidden_size, num_categories, feature_dim, prior_concentration=1.0, min_temperature=1.0,
				epoch_init_iter_counts=0, temperature_update_freq=1000, temperature_anneal_rate=1e-5):
		super(ABCDSampler, self).__init__()
		self.num_categories = num_categories
		self.to_code_like = MLP(input_size, mlp_hidden_size, feature_dim)
		self.min_temperature = min_temperature
		self.epoch_init_iter_counts = epoch_init_iter_counts
		self.iter_counts = epoch_init_iter_counts
		self.temperature_update_freq = temperature_update_freq
		self.temperature_anneal_rate = temperature_anneal_rate
		self.update_temperature((self.iter_counts // self.temperature_update_freq) * self.temperature_update_freq)
		self.register_buffer('prior_concentration', torch.tensor(prior_concentration))
		self._prior_c = float(prior_concentration)
		# same creation order as the reference: posterior_shape_logits, then codebook
		self.posterior_shape_logits = torch.nn.Parameter(torch.randn(num_categories), requires_grad=True)
		self.codebook = torch.nn.Parameter(torch.randn((feature_dim, num_categories)), requires_grad=True)
		self.injected_gumbel_noise = None   # parity hook: [B,K] standard Gumbel draws used by the next sample()
		self.last_sample_probs = None

	def forward(self, x):
		"""Unnormalised log posterior q(z|x): scaled dot products between to_code_like(x) and the codebook."""
		net = self.to_code_like.whole_network
		return _LogitsFn.apply(x, net[0].weight, net[0].bias, net[2].weight, net[2].bias, self.codebook)

	def sample(self, logits, no_sample=False):
		"""Gumbel-softmax relaxation y of the category, returned as the codebook-weighted feature y @ codebook^T.
		With no_sample=True the class probabilities are used in place of y (pretraining)."""
		if no_sample:
			mode, noise = 0, None
		else:
			mode = 1
			noise = self.injected_gumbel_noise
			self.injected_gumbel_noise = None
			if noise is None:
				noise = noise_like(logits, kind='gumbel')
			noise = _f32(noise)
		feats, y = _SampleFn.apply(logits, self.codebook, mode, self.temperature, noise)
		self.last_sample_probs = y
		return feats

	def fused_supported(self):
		return bool(L.load().abcd_sampler_fused_supported(self.codebook.shape[0], self.codebook.shape[1]))

	def forward_sample_kl(self, x, entire_data_size, no_sample=False):
		"""forward() + sample() + kl_divergence() of the reference (learning.py:150-152) with ONE kernel after the to_code_like
		MLP (and one for its backward): returns (features, KL summed over the batch, category logits).  The logits come back
		detached (they only feed the perplexity log line and the arg-max)."""
		xc = self.to_code_like(x)
		noise = None
		if not no_sample:
			noise = self.injected_gumbel_noise
			self.injected_gumbel_noise = None
			if noise is not None:
				noise = _f32(noise)
		feats, kl, logits, y = _SamplerFusedFn.apply(xc, self.codebook, self.posterior_shape_logits, 0 if no_sample else 1,
													self.temperature, noise, float(self._prior_c), float(entire_data_size))
		self.last_sample_probs = y
		return feats, kl, logits

	def kl_divergence(self, logits, entire_data_size):
		"""KL(q(pi) q(z|x) || p(pi) p(z|pi)) summed over the batch."""
		self._prior_c = float(self._prior_c)
		return _KlFn.apply(logits, self.posterior_shape_logits, self._prior_c, float(entire_data_size))

	def log_pmf(self, targets, logits):
		return torch.nn.functional.cross_entropy(logits, targets, reduction='sum')

	def perplexities(self, logits):
		"""(cluster, batch-mean, Dirichlet-shape) perplexities of Learner.train's log line, as a [3] device tensor."""
		L.ensure_device(logits.device)
		logits = _f32(logits.detach())
		B, K = logits.shape
		scratch = torch.empty(K, dtype=torch.float32, device=logits.device)
		out = torch.empty(3, dtype=torch.float32, device=logits.device)
		L.call('abcd_sampler_stats', B, K, L.ptr(logits), L.ptr(self.posterior_shape_logits.data), L.ptr(scratch), L.ptr(out), L.stream_ptr())
		return out

	def increment_iter_counts(self):
		self.iter_counts += 1
		if self.iter_counts % self.temperature_update_freq == 0:
			self.update_temperature()

	def update_epoch_init_iter_counts(self):
		self.epoch_init_iter_counts = self.iter_counts

	def update_temperature(self, steps=None):
		if steps is None:
			steps = self.iter_counts
		self.temperature = min(self.min_temperature, math.exp(-self.temperature_anneal_rate * steps))

	def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
		super(ABCDSampler, self)._load_from_state_dict(state_dict, prefix, *args, **kwargs)
		self._prior_c = float(self.prior_concentration.item())

	def pack_init_parameters(self):
		return {
			'input_size': self.to_code_like.input_size,
			'mlp_hidden_size': self.to_code_like.hidden_size,
			'num_categories': self.num_categories,
			'feature_dim': self.to_code_like.output_size,
			'prior_concentration': self._prior_c,
			'min_temperature': self.min_temperature,
			'epoch_init_iter_counts': self.epoch_init_iter_counts,
			'temperature_update_freq': self.temperature_update_freq,
			'temperature_anneal_rate': self.temperature_anneal_rate,
		}


# ------------------------------------------------------------------------------------------------------
# Decoder (reference: model.py:84-300)
# ------------------------------------------------------------------------------------------------------
class RNN_Cell(torch.nn.Module):
	"""Parameter container matching the reference's RNN_Cell (keys rnn_cell.cell.*)."""

	def __init__(self, input_size, hidden_size, model_type='LSTM', input_dropout=0.0, esn_leak=1.0):
		super(RNN_Cell, self).__init__()
		_cell_code(model_type)
		self.drop = torch.nn.Dropout(input_dropout)
		self.mode = model_type
		self.cell = getattr(torch.nn, model_type + 'Cell')(input_size, hidden_size)


_DEC_PARAM_ORDER = ['w_f2h', 'b_f2h', 'w_ih', 'w_hh', 'b_ih', 'b_hh',
					'em_w1_0', 'em_b1_0', 'em_w2_0', 'em_b2_0', 'em_w1_1', 'em_b1_1', 'em_w2_1', 'em_b2_1',
					'off_w1', 'off_b1', 'off_w2', 'off_b2']


def _dec_structs(cell, F, H, M, fin, params, grads=None):
	dp = L.DecParamsT()
	dp.cell, dp.F, dp.H, dp.M, dp.fin = cell, F, H, M, fin
	(dp.w_f2h, dp.b_f2h, dp.w_ih, dp.w_hh, dp.b_ih, dp.b_hh) = [L.ptr(p) for p in params[:6]]
	for i in range(2):
		dp.em_w1[i], dp.em_b1[i], dp.em_w2[i], dp.em_b2[i] = [L.ptr(p) for p in params[6 + 4 * i:10 + 4 * i]]
	dp.off_w1, dp.off_b1, dp.off_w2, dp.off_b2 = [L.ptr(p) for p in params[14:18]]
	dg = L.DecGradsT()
	if grads is not None:
		(dg.w_f2h, dg.b_f2h, dg.w_ih, dg.w_hh, dg.b_ih, dg.b_hh) = [L.ptr(p) for p in grads[:6]]
		for i in range(2):
			dg.em_w1[i], dg.em_b1[i], dg.em_w2[i], dg.em_b2[i] = [L.ptr(p) for p in grads[6 + 4 * i:10 + 4 * i]]
		dg.off_w1, dg.off_b1, dg.off_w2, dg.off_b2 = [L.ptr(p) for p in grads[14:18]]
	return dp, dg


class _DecoderFn(torch.autograd.Function):
	@staticmethod
	def forward(ctx, feat_in, eps, gt, is_offset, desc, cell, dims, *params):
		lib = L.ensure_device(feat_in.device)
		F, H, M, want_outputs = dims
		feat_in, eps = _f32(feat_in), _f32(eps)
		gt = _f32(gt) if gt is not None else None
		is_offset = _f32(is_offset) if is_offset is not None else None
		params = [_f32(p) for p in params]
		dev = feat_in.device
		fin = feat_in.shape[1]
		ws_bytes = lib.abcd_dec_workspace_bytes(cell, F, fin, desc.B, desc.sumT)
		ws = _ws(ws_bytes, dev)
		losses = torch.zeros(2, dtype=torch.float32, device=dev)
		if want_outputs:
			out_x = torch.empty(desc.sumT, F, dtype=torch.float32, device=dev)
			out_mean = torch.empty_like(out_x)
			out_lv = torch.empty_like(out_x)
			out_off = torch.empty(desc.sumT, dtype=torch.float32, device=dev)
		else:   # losses only (Learner.train_step): the kernels skip the unpadded copies
			out_x = out_mean = out_lv = out_off = None
		dp, _ = _dec_structs(cell, F, H, M, fin, params)
		L.call('abcd_dec_fwd', C.byref(dp), desc.ref(), L.ptr(feat_in), L.ptr(eps), eps.shape[1],
			L.ptr(gt), gt.shape[1] if gt is not None else 0, L.ptr(is_offset), L.ptr(losses),
			L.ptr(out_x), L.ptr(out_mean), L.ptr(out_lv), L.ptr(out_off), L.ptr(ws), ws_bytes, L.stream_ptr())
		ctx.desc, ctx.cell, ctx.dims, ctx.ws, ctx.ws_bytes = desc, cell, (F, H, M), ws, ws_bytes
		ctx.has_targets = gt is not None and is_offset is not None
		ctx.save_for_backward(feat_in, eps, gt if gt is not None else eps, is_offset if is_offset is not None else eps, *params)
		if want_outputs:
			ctx.mark_non_differentiable(out_x, out_mean, out_lv, out_off)
		ctx.set_materialize_grads(False)   # no zero-filled [sumT, F] gradients for the non-differentiable outputs
		return losses[0], losses[1], out_x, out_mean, out_lv, out_off

	@staticmethod
	def backward(ctx, g_em, g_off, *_unused):
		if not ctx.has_targets:
			raise RuntimeError('decoder backward needs ground_truth_out and ground_truth_offset')
		feat_in, eps, gt, is_offset, *params = ctx.saved_tensors
		F, H, M = ctx.dims
		dev = feat_in.device
		if g_em is not None and g_off is not None:
			g = torch.stack((g_em.reshape(()), g_off.reshape(()))).to(torch.float32)   # one kernel instead of fill + 2 copies
		else:
			g = torch.zeros(2, dtype=torch.float32, device=dev)
			if g_em is not None:
				g[0] = g_em
			if g_off is not None:
				g[1] = g_off
		grads, rets = _grad_buffers(params)
		d_feat = torch.empty_like(feat_in)
		dp, dg = _dec_structs(ctx.cell, F, H, M, feat_in.shape[1], params, grads)
		side = _WGRAD_STREAM
		if side is None or any(r is not None for r in rets):
			L.call('abcd_dec_bwd', C.byref(dp), C.byref(dg), ctx.desc.ref(), L.ptr(feat_in), L.ptr(eps), eps.shape[1],
				L.ptr(gt), gt.shape[1], L.ptr(is_offset), L.ptr(g), L.ptr(d_feat), L.ptr(ctx.ws), ctx.ws_bytes, L.stream_ptr())
		else:
			# data path here; weight gradients (accumulated straight into the parameters' .grad) on the side stream
			L.call('abcd_dec_bwd_data', C.byref(dp), ctx.desc.ref(), L.ptr(feat_in), L.ptr(eps), eps.shape[1],
				L.ptr(gt), gt.shape[1], L.ptr(is_offset), L.ptr(g), L.ptr(d_feat), L.ptr(ctx.ws), ctx.ws_bytes, L.stream_ptr())
			side.wait_stream(torch.cuda.current_stream())
			with torch.cuda.stream(side):
				L.call('abcd_dec_bwd_weights', C.byref(dp), C.byref(dg), ctx.desc.ref(), L.ptr(feat_in), L.ptr(ctx.ws), ctx.ws_bytes, L.stream_ptr())
			for t in (ctx.ws, feat_in):
				t.record_stream(side)
		return (d_feat, None, None, None, None, None, None) + tuple(rets)


class RNN_Variational_Decoder(torch.nn.Module):
	"""Autoregressive decoder with Gaussian emissions and self-feedback of the sampled frame."""

	def __init__(self, output_size, rnn_hidden_size, mlp_hidden_size, feature_size, emission_distr_name='isotropic_gaussian',
				rnn_type='LSTM', rnn_layers=1, input_dropout=0.0, self_feedback=True, bidirectional=False, right2left_weight=0.5,
				esn_leak=1.0, num_speakers=None, speaker_embed_dim=None):
		super(RNN_Variational_Decoder, self).__init__()
		assert rnn_layers == 1, 'Only rnn_layers=1 is currently supported.'
		_cell_code(rnn_type)
		if bidirectional:
			raise NotImplementedError('the bidirectional decoder is outside the B200 hot path')
		if not self_feedback or input_dropout != 0.0:
			raise NotImplementedError('decoder_input_dropout / greedy_decoder are outside the B200 hot path (must be 0.0 / off)')
		hidden_size_total = rnn_layers * rnn_hidden_size
		if rnn_type == 'LSTM':
			hidden_size_total *= 2
		self.bidirectional = bidirectional
		self.feature_size = feature_size   # without speaker_embed_dim
		if num_speakers is None or speaker_embed_dim is None:
			self.embed_speaker = None
		else:
			self.embed_speaker = torch.nn.Embedding(num_speakers, speaker_embed_dim)
			feature_size += speaker_embed_dim
		# creation order = the reference's, so that seeded initialisation matches
		self.feature2hidden = torch.nn.Linear(feature_size, hidden_size_total)
		self.offset_predictor = MLP(rnn_hidden_size, mlp_hidden_size, 1)
		self.emission_sampler = Sampler(rnn_hidden_size, mlp_hidden_size, output_size, distribution_name=emission_distr_name)
		self.rnn_cell = RNN_Cell(output_size, rnn_hidden_size, model_type=rnn_type, input_dropout=input_dropout, esn_leak=esn_leak)
		self.injected_emission_noise = None   # parity hook: [sumT, F] N(0,1) draws in packed order for the next forward
		self._desc_cache = {}

	def _params(self):
		cell = self.rnn_cell.cell
		ps = [self.feature2hidden.weight, self.feature2hidden.bias, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
		for mlp in self.emission_sampler.to_parameters.mlps:
			net = mlp.whole_network
			ps += [net[0].weight, net[0].bias, net[2].weight, net[2].bias]
		net = self.offset_predictor.whole_network
		ps += [net[0].weight, net[0].bias, net[2].weight, net[2].bias]
		return ps

	def pack_init_parameters(self):
		parameters = {
			'output_size': self.rnn_cell.cell.input_size,
			'rnn_hidden_size': self.rnn_cell.cell.hidden_size,
			'mlp_hidden_size': self.offset_predictor.hidden_size,
			'feature_size': self.feature_size,
			'emission_distr_name': self.emission_sampler.distribution_name,
			'rnn_type': self.rnn_cell.mode,
			'rnn_layers': 1,
			'input_dropout': self.rnn_cell.drop.p,
			'bidirectional': self.bidirectional,
		}
		if self.embed_speaker is not None:
			parameters['num_speakers'] = self.embed_speaker.num_embeddings
			parameters['speaker_embed_dim'] = self.embed_speaker.embedding_dim
		return parameters

	def _length_to_batch_sizes(self, lengths):
		return [(lengths > t).sum() for t in range(int(lengths.max()))]

	def forward(self, features, lengths=None, batch_sizes=None, speaker=None, ground_truth_out=None, ground_truth_offset=None,
				return_outputs=True):
		"""Returns (emission_loss, offset_loss, flatten_out, (mean, log_variance), flatten_offset_weights); the
		flattened tensors are in PackedSequence.data order.  return_outputs=False (not in the reference; used by
		Learner, which only needs the losses) returns None for the flattened tensors and skips their copies."""
		assert (lengths is not None) or (batch_sizes is not None), 'Either lengths or batch_sizes must be given.'
		if lengths is not None:
			batch_sizes = torch.stack(self._length_to_batch_sizes(lengths)).cpu()
		if self.embed_speaker is not None:
			features = torch.cat([features, self.embed_speaker(speaker)], dim=-1)
		desc = as_packed_desc(batch_sizes, features.device, self._desc_cache)
		cell = self.rnn_cell.cell
		F, H, M = cell.input_size, cell.hidden_size, self.offset_predictor.hidden_size
		eps = self.injected_emission_noise
		self.injected_emission_noise = None
		if eps is None:
			eps = noise_like(torch.empty(desc.sumT, F, device=features.device))
		em, off, out_x, mean, lv, off_logit = _DecoderFn.apply(
			features, eps, ground_truth_out, ground_truth_offset, desc, _cell_code(self.rnn_cell.mode), (F, H, M, bool(return_outputs)), *self._params())
		if ground_truth_out is None:
			em = None
		if ground_truth_offset is None:
			off = None
		return em, off, out_x, (mean, lv), off_logit
