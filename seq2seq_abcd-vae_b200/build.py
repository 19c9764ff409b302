# coding: utf-8
"""In-tree build of libabcd_b200.so (sm_100a only) with nvcc.  No torch headers are involved: the library
is a plain C ABI (include/abcd_b200.h) loaded through ctypes."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libabcd_b200.so')
SOURCES = ['api.cu', 'sgemm.cu', 'umma_gemm.cu', 'umma_rm_gemm.cu', 'latency_probe.cu', 'enc.cu', 'dec.cu', 'sampler.cu', 'sampler_fused.cu', 'stft.cu', 'optim.cu', 'csv.cu']
# layout probes used by tools/ (not product code): a separate test library, built on demand by build_probe()
PROBE_LIB_PATH = os.path.join(HERE, 'libabcd_b200_probe.so')
PROBE_SOURCES = ['umma_probe.cu', 'probe_api.cu']
NVCC_FLAGS = [
	'-gencode', 'arch=compute_100a,code=sm_100a',
	'-O3', '-std=c++17', '-lineinfo',
	'-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
	'--expt-relaxed-constexpr',
]


def _nvcc():
	for cand in (shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
		if cand and os.path.isfile(cand):
			return cand
	raise RuntimeError('nvcc not found; the CUDA extension cannot be built')


def _source_hash():
	h = hashlib.sha256()
	files = sorted(os.listdir(CSRC)) + ['../../include/abcd_b200.h']
	for name in files:
		path = os.path.join(CSRC, name)
		if os.path.isfile(path) and (name.endswith(('.cu', '.cuh', '.h'))):
			with open(path, 'rb') as f:
				h.update(name.encode())
				h.update(f.read())
	h.update(' '.join(NVCC_FLAGS).encode())
	return h.hexdigest()


def build(force=False, verbose=False):
	"""Compile every .cu for sm_100a and link the shared library next to this file."""
	stamp = LIB_PATH + '.hash'
	digest = _source_hash()
	if not force and os.path.isfile(LIB_PATH) and os.path.isfile(stamp):
		with open(stamp) as f:
			if f.read().strip() == digest:
				return LIB_PATH
	nvcc = _nvcc()
	obj_dir = os.path.join(HERE, 'build')
	os.makedirs(obj_dir, exist_ok=True)
	procs = []
	objs = []
	for src in SOURCES:
		obj = os.path.join(obj_dir, src.replace('.cu', '.o'))
		objs.append(obj)
		cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
		if verbose:
			cmd.insert(1, '-Xptxas')
			cmd.insert(2, '-v')
		procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
	failed = False
	for src, p in procs:
		out, _ = p.communicate()
		if p.returncode != 0:
			failed = True
			sys.stderr.write('nvcc failed on {}:\n{}\n'.format(src, out))
		elif verbose and out:
			sys.stderr.write(out)
	if failed:
		raise RuntimeError('building libabcd_b200.so failed')
	cmd = [nvcc, '-shared', '-cudart', 'static', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIB_PATH] + objs
	subprocess.check_call(cmd)
	with open(stamp, 'w') as f:
		f.write(digest)
	return LIB_PATH


def build_probe():
	"""libabcd_b200_probe.so: the tcgen05 TS-form layout probe of tools/ts_probe.py (kept out of the product library)."""
	nvcc = _nvcc()
	cmd = [nvcc] + NVCC_FLAGS + ['-shared', '-cudart', 'static', '-o', PROBE_LIB_PATH] + [os.path.join(CSRC, s) for s in PROBE_SOURCES]
	subprocess.check_call(cmd)
	return PROBE_LIB_PATH


if __name__ == '__main__':
	if '--probe' in sys.argv:
		print(build_probe())
		sys.exit(0)
	print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
