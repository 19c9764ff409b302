# coding: utf-8
"""Debug aid for the tcgen05 TN GEMM: structured inputs, prints error pattern."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module('seq2seq_abcd-vae_b200')
L = pkg._lib
L.ensure_device(torch.device('cuda'))
torch.manual_seed(0)
for (M, N, K) in ((128, 32, 256), (128, 64, 256), (1024, 256, 4096), (1024, 65, 3000)):
	for mode in ('ones', 'rowid', 'rand'):
		if mode == 'ones':
			A = torch.ones(K, M, device='cuda'); B = torch.ones(K, N, device='cuda')
		elif mode == 'rowid':
			A = torch.zeros(K, M, device='cuda'); A[0] = torch.arange(M, device='cuda').float()
			B = torch.zeros(K, N, device='cuda'); B[0] = torch.arange(N, device='cuda').float() + 1
		else:
			A = torch.randn(K, M, device='cuda'); B = torch.randn(K, N, device='cuda')
		want = A.double().t() @ B.double()
		C = torch.zeros(M, N, device='cuda'); part = torch.empty(2 * M * N, device='cuda')
		L.call('abcd_sgemm', M, N, K, L.ptr(A), M, 0, L.ptr(B), N, 0, L.ptr(C), N, 1.0, None, 0, 0, 2, L.ptr(part), L.stream_ptr())
		torch.cuda.synchronize()
		err = (C.double() - want).abs().max().item() / (want.abs().max().item() + 1e-30)
		print('M%d N%d K%d %-5s rel err %.3e  got[0:2,0:6]=%s want=%s' % (M, N, K, mode, err, C[0:2, 0:6].tolist(), want[0:2, 0:6].tolist()))
		if mode == 'rowid' and err > 1e-5:
			got = C.cpu()
			print('   got[5, :12]', got[5, :12].tolist()); print('   got[:12, 3]', got[:12, 3].tolist())
