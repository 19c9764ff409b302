# coding: utf-8
"""Does tcgen05.mma accept its A operand from tensor memory laid out "lane = row, one 32-bit column per k" for kind::tf32?
    python tools/ts_probe.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
pkg = importlib.import_module('seq2seq_abcd-vae_b200')
L = pkg._lib
L.ensure_device(torch.device('cuda'))
# the probe lives in its own test library (python seq2seq_abcd-vae_b200/build.py --probe), not in the product .so
_build = importlib.import_module('seq2seq_abcd-vae_b200.build')
if not os.path.isfile(_build.PROBE_LIB_PATH):
	_build.build_probe()
_probe = C.CDLL(_build.PROBE_LIB_PATH)
_probe.abcd_debug_ts_probe.restype = C.c_int
_probe.abcd_debug_ts_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
_probe.abcd_probe_last_error.restype = C.c_char_p


def probe(A, B, D, K, N, variant):
	rc = _probe.abcd_debug_ts_probe(L.ptr(A), L.ptr(B), L.ptr(D), K, N, variant, L.stream_ptr())
	if rc != 0:
		raise RuntimeError(_probe.abcd_probe_last_error().decode())


torch.manual_seed(0)
for (K, N) in ((8, 16), (32, 64), (64, 96), (256, 80)):
	A = torch.randn(128, K, device='cuda'); B = torch.randn(K, N, device='cuda')
	want = (A.double() @ B.double())
	for variant in (0,):   # a_major = 1 (MN) with a tensor-memory A operand is an illegal instruction
		D = torch.full((128, N), 7.0, device='cuda')
		try:
			probe(A, B, D, K, N, variant)
			torch.cuda.synchronize()
			err = float((D.double() - want).abs().max() / want.abs().max())
			print('K %3d N %3d a_major %d: max rel err %.3e %s' % (K, N, variant, err, 'OK (tf32 single pass)' if err < 5e-3 else 'MISMATCH'))
		except Exception as e:
			print('K %3d N %3d a_major %d: FAILED %r' % (K, N, variant, e))
			raise SystemExit(1)

# M = 64: where do the 64 accumulator rows land in tensor memory, for two placements of the A rows?
K, N = 32, 32
A = torch.randn(128, K, device='cuda'); B = torch.randn(K, N, device='cuda')
want = (A.double() @ B.double())
for variant, what in ((2, 'TS, A rows 0..63 in lanes 0..63'), (6, 'TS, A row 16q+i in lane 32q+i (i<16)'), (10, 'SS, A rows 0..63 as MN-major smem tiles')):
	D = torch.full((128, N), 7.0, device='cuda')
	probe(A, B, D, K, N, variant)
	torch.cuda.synchronize()
	Dd = D.double()
	rows = []
	for lane in range(128):
		hit = [m for m in range(128) if float((Dd[lane] - want[m]).abs().max()) < 5e-3 * float(want.abs().max())]
		rows.append(hit[0] if hit else -1)
	print('M=64,', what, ': accumulator lane -> A row:', rows)
