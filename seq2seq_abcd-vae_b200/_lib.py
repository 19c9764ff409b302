# coding: utf-8
"""ctypes binding of libabcd_b200.so (include/abcd_b200.h).

There is no CPU or PyTorch fallback behind these calls: if the shared library is missing, or a call is
made without a B200, an exception is raised.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libabcd_b200.so')

c_f32p = C.c_void_p   # device pointers travel as integers
c_i32p = C.c_void_p


class PackedT(C.Structure):
	_fields_ = [('B', C.c_int32), ('T_max', C.c_int32), ('sumT', C.c_int32),
				('batch_sizes', c_i32p), ('offsets', c_i32p), ('lengths', c_i32p)]


class EncParamsT(C.Structure):
	_fields_ = [('cell', C.c_int32), ('F', C.c_int32), ('H', C.c_int32), ('bidir', C.c_int32),
				('w_ih', c_f32p * 2), ('w_hh', c_f32p * 2), ('b_ih', c_f32p * 2), ('b_hh', c_f32p * 2)]


class EncGradsT(C.Structure):
	_fields_ = [('w_ih', c_f32p * 2), ('w_hh', c_f32p * 2), ('b_ih', c_f32p * 2), ('b_hh', c_f32p * 2)]


class MlpParamsT(C.Structure):
	_fields_ = [('in_', C.c_int32), ('hidden', C.c_int32), ('out', C.c_int32),
				('w1', c_f32p), ('b1', c_f32p), ('w2', c_f32p), ('b2', c_f32p)]


class MlpGradsT(C.Structure):
	_fields_ = [('w1', c_f32p), ('b1', c_f32p), ('w2', c_f32p), ('b2', c_f32p)]


class SamplerParamsT(C.Structure):
	_fields_ = [('mlp', MlpParamsT), ('f', C.c_int32), ('K', C.c_int32),
				('codebook', c_f32p), ('shape_logits', c_f32p), ('prior_c', C.c_float)]


class SamplerGradsT(C.Structure):
	_fields_ = [('mlp', MlpGradsT), ('codebook', c_f32p), ('shape_logits', c_f32p)]


class DecParamsT(C.Structure):
	_fields_ = [('cell', C.c_int32), ('F', C.c_int32), ('H', C.c_int32), ('M', C.c_int32), ('fin', C.c_int32),
				('w_f2h', c_f32p), ('b_f2h', c_f32p),
				('w_ih', c_f32p), ('w_hh', c_f32p), ('b_ih', c_f32p), ('b_hh', c_f32p),
				('em_w1', c_f32p * 2), ('em_b1', c_f32p * 2), ('em_w2', c_f32p * 2), ('em_b2', c_f32p * 2),
				('off_w1', c_f32p), ('off_b1', c_f32p), ('off_w2', c_f32p), ('off_b2', c_f32p)]


class DecGradsT(C.Structure):
	_fields_ = [('w_f2h', c_f32p), ('b_f2h', c_f32p),
				('w_ih', c_f32p), ('w_hh', c_f32p), ('b_ih', c_f32p), ('b_hh', c_f32p),
				('em_w1', c_f32p * 2), ('em_b1', c_f32p * 2), ('em_w2', c_f32p * 2), ('em_b2', c_f32p * 2),
				('off_w1', c_f32p), ('off_b1', c_f32p), ('off_w2', c_f32p), ('off_b2', c_f32p)]


# name -> (restype, argtypes); the list doubles as the export check run by the CPU test-suite
P = C.POINTER
SIGNATURES = {
	'abcd_init': (C.c_int, [C.c_int]),
	'abcd_last_error': (C.c_char_p, []),
	'abcd_version': (C.c_char_p, []),
	'abcd_max_active_clusters': (C.c_int, [C.c_int, C.c_int]),
	'abcd_debug_enc_timing': (C.c_int, [P(C.c_ulonglong)]),
	'abcd_set_rnn_tensorcore': (C.c_int, [C.c_int]),
	'abcd_debug_dec_timing': (C.c_int, [P(C.c_ulonglong)]),
	'abcd_measure_cluster_exchange': (C.c_int, [C.c_int, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
	'abcd_launch_count': (C.c_longlong, []),
	'abcd_profile_groups': (C.c_int, []),
	'abcd_profile_name': (C.c_char_p, [C.c_int]),
	'abcd_profile_enable': (None, [C.c_int]),
	'abcd_profile_read': (C.c_int, [C.c_int, P(C.c_float), P(C.c_int)]),
	'abcd_stft_logamp': (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, c_i32p, P(PackedT), C.c_int, C.c_int, C.c_int,
								c_f32p, C.c_float, C.c_float, c_f32p, C.c_int64, C.c_void_p]),
	'abcd_enc_workspace_bytes': (C.c_size_t, [C.c_int] * 6),
	'abcd_enc_fwd': (C.c_int, [P(EncParamsT), P(PackedT), c_f32p, C.c_int, c_f32p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_enc_bwd': (C.c_int, [P(EncParamsT), P(EncGradsT), P(PackedT), c_f32p, C.c_int, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_sampler_logits_fwd': (C.c_int, [P(SamplerParamsT), C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_logits_bwd': (C.c_int, [P(SamplerParamsT), P(SamplerGradsT), C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_sample_fwd': (C.c_int, [P(SamplerParamsT), C.c_int, c_f32p, C.c_int, C.c_float, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_sample_bwd': (C.c_int, [P(SamplerParamsT), P(SamplerGradsT), C.c_int, c_f32p, C.c_int, C.c_float, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_kl_ws_floats': (C.c_size_t, [C.c_int]),
	'abcd_sampler_kl_fwd': (C.c_int, [P(SamplerParamsT), C.c_int, c_f32p, C.c_float, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_kl_bwd': (C.c_int, [P(SamplerParamsT), P(SamplerGradsT), C.c_int, c_f32p, C.c_float, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_fused_supported': (C.c_int, [C.c_int, C.c_int]),
	'abcd_sampler_fused_ws_floats': (C.c_size_t, [C.c_int]),
	'abcd_sampler_fused_fwd': (C.c_int, [P(SamplerParamsT), C.c_int, c_f32p, C.c_int, C.c_float, c_f32p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_float,
									c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_fused_bwd': (C.c_int, [P(SamplerParamsT), P(SamplerGradsT), C.c_int, c_f32p, c_f32p, c_f32p, C.c_int, C.c_float, C.c_float,
									c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_sampler_stats': (C.c_int, [C.c_int, C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_softmax_argmax': (C.c_int, [C.c_int, C.c_int, c_f32p, c_f32p, c_i32p, C.c_void_p]),
	'abcd_mlp_fwd': (C.c_int, [P(MlpParamsT), C.c_int, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_mlp_bwd': (C.c_int, [P(MlpParamsT), P(MlpGradsT), C.c_int, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_gauss_latent_fwd': (C.c_int, [C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_gauss_latent_bwd': (C.c_int, [C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
	'abcd_dec_workspace_bytes': (C.c_size_t, [C.c_int] * 5),
	'abcd_dec_fwd': (C.c_int, [P(DecParamsT), P(PackedT), c_f32p, c_f32p, C.c_int, c_f32p, C.c_int, c_f32p, c_f32p,
							c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_dec_bwd': (C.c_int, [P(DecParamsT), P(DecGradsT), P(PackedT), c_f32p, c_f32p, C.c_int, c_f32p, C.c_int, c_f32p,
							c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_dec_bwd_data': (C.c_int, [P(DecParamsT), P(PackedT), c_f32p, c_f32p, C.c_int, c_f32p, C.c_int, c_f32p,
							c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_dec_bwd_weights': (C.c_int, [P(DecParamsT), P(DecGradsT), P(PackedT), c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
	'abcd_optim_partial_count': (C.c_int, []),
	'abcd_gradnorm_clip_sgd': (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_int,
									c_f32p, c_f32p, C.c_void_p]),
	'abcd_philox_normal': (C.c_int, [c_f32p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
	'abcd_philox_gumbel': (C.c_int, [c_f32p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
	'abcd_philox_fill': (C.c_int, [c_f32p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
	'abcd_counter_add': (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
	'abcd_csv_row_bound': (C.c_int64, [C.c_int64]),
	'abcd_csv_format_long': (C.c_int64, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
	'abcd_sgemm': (C.c_int, [C.c_int, C.c_int, C.c_int, c_f32p, C.c_int64, C.c_int, c_f32p, C.c_int64, C.c_int,
							c_f32p, C.c_int64, C.c_float, c_f32p, C.c_int, C.c_int, C.c_int, c_f32p, C.c_void_p]),
}

_lib = None
_inited_devices = set()
LAUNCH_COUNT = 0   # C-ABI compute calls made through this binding (bench.py reports it)


class AbcdError(RuntimeError):
	pass


def load():
	"""dlopen the shared library and bind every symbol of include/abcd_b200.h.  Raises if it is not built."""
	global _lib
	if _lib is not None:
		return _lib
	if not os.path.isfile(LIB_PATH):
		raise ImportError(
			'libabcd_b200.so is not built ({}). Run `python -c "import __graft_entry__ as g; g.build()"` '
			'(needs nvcc); there is no CPU fallback.'.format(LIB_PATH))
	lib = C.CDLL(LIB_PATH)
	for name, (res, args) in SIGNATURES.items():
		fn = getattr(lib, name)   # AttributeError if the export is missing
		fn.restype = res
		fn.argtypes = args
	_lib = lib
	return lib


def ensure_device(device):
	"""abcd_init on first use of a CUDA device."""
	lib = load()
	if not torch.cuda.is_available():
		raise AbcdError('abcd_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
	idx = device.index if device.index is not None else torch.cuda.current_device()
	if idx not in _inited_devices:
		rc = lib.abcd_init(idx)
		if rc != 0:
			raise AbcdError('abcd_init({}) failed: {}'.format(idx, lib.abcd_last_error().decode()))
		_inited_devices.add(idx)
	return lib


def check(rc, what):
	if rc != 0:
		raise AbcdError('{} failed ({}): {}'.format(what, rc, load().abcd_last_error().decode()))


def ptr(t):
	"""Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
	if t is None:
		return None
	assert t.is_cuda and t.is_contiguous(), 'abcd_b200 ops need contiguous CUDA tensors'
	return t.data_ptr()


def stream_ptr():
	return torch.cuda.current_stream().cuda_stream


def call(name, *args):
	global LAUNCH_COUNT
	lib = load()
	LAUNCH_COUNT += 1
	check(getattr(lib, name)(*args), name)


def profile_enable(on):
	load().abcd_profile_enable(int(bool(on)))


def profile_read_all():
	"""{kernel group: (total milliseconds, launches)} of the spans recorded since the last read."""
	lib = load()
	out = {}
	for i in range(lib.abcd_profile_groups()):
		ms, n = C.c_float(0), C.c_int(0)
		check(lib.abcd_profile_read(i, C.byref(ms), C.byref(n)), 'abcd_profile_read')
		out[lib.abcd_profile_name(i).decode()] = (ms.value, n.value)
	return out


class Packed(object):
	"""Device-side descriptor of a PackedSequence layout (batch_sizes -> offsets, lengths)."""

	def __init__(self, batch_sizes, device):
		bs = torch.as_tensor(batch_sizes, dtype=torch.int64, device='cpu')
		assert bs.dim() == 1 and bs.numel() > 0 and int(bs[0]) > 0, 'empty PackedSequence'
		assert bool((bs[1:] <= bs[:-1]).all()), 'batch_sizes must be non-increasing'
		self.batch_sizes_cpu = bs
		self.B = int(bs[0])
		self.T_max = int(bs.numel())
		self.sumT = int(bs.sum())
		offsets = torch.zeros(self.T_max + 1, dtype=torch.int64)
		offsets[1:] = torch.cumsum(bs, 0)
		# lengths[b] = #steps with batch_size > b
		lengths = (bs[None, :] > torch.arange(self.B)[:, None]).sum(1)
		meta = torch.cat([bs, offsets, lengths]).to(torch.int32)
		self.meta = meta.to(device, non_blocking=True)
		self.batch_sizes = self.meta[:self.T_max]
		self.offsets = self.meta[self.T_max:2 * self.T_max + 1]
		self.lengths = self.meta[2 * self.T_max + 1:]
		self.lengths_cpu = lengths
		self.offsets_cpu = offsets
		self.struct = PackedT(self.B, self.T_max, self.sumT, self.batch_sizes.data_ptr(),
							self.offsets.data_ptr(), self.lengths.data_ptr())

	def ref(self):
		return C.byref(self.struct)
