# coding: utf-8
"""Turn ncu outputs brought back in gpurun_out/ into small text summaries under profiles/ (tracked).
    python tools/summarize_profiles.py <launches.csv | -> <report.ncu-rep | raw.csv | -> <tag> [<second report> ...]"""
import csv
import json
import os
import re
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__cluster_size', 'launch__block_size',
		'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
		'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
		'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
		'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
		'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
		'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_membar', 'smsp__pcsamp_warps_issue_stalled_wait',
		'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_not_selected', 'smsp__pcsamp_sample_count']


# kernels of bench.py's own measurements (cuBLAS TF32 peak, cluster-exchange probe, STFT front-end timing): not part of a training step
NOT_STEP = ('cutlass', 'cluster_exchange_kernel', 'stft_logamp')


def launches(path, out):
	rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
	agg, other, total, total_other = {}, {}, 0.0, 0.0
	for r in rows:
		name = re.sub(r'\(.*', '', r[4]).replace('void <unnamed>::', '').replace('void at::native::', 'at::')[:64]
		if 'elementwise' in name:
			name = 'torch elementwise (autograd accumulation / fills)'
		us = float(r[14]) / 1000.0
		if any(k in name for k in NOT_STEP):
			a = other.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
			total_other += us
			continue
		a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
		total += us
	out.write('launch list: {} launches of the training steps, {:.1f} us total (ncu gpu__time_duration.sum, --clock-control none; cold-cache, '
			'serialised: compare shares)\n'.format(sum(c for c, _ in agg.values()), total))
	for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
		out.write('  {:<66s} x{:<4d} {:>10.1f} us  {:5.1f}%\n'.format(k, c, v, 100 * v / total))
	if other:
		out.write('not part of a step (bench.py\'s peak / probe / front-end measurements), {:.1f} us:\n'.format(total_other))
		for k, (c, v) in sorted(other.items(), key=lambda kv: -kv[1][1]):
			out.write('  {:<66s} x{:<4d} {:>10.1f} us\n'.format(k, c, v))


GROUPS = (('enc_fwd', 'enc_fwd_kernel'), ('enc_bwd', 'enc_bwd_kernel'), ('dec_fwd', 'dec_fwd_kernel'), ('dec_bwd', 'dec_bwd_kernel'))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def raw(rep, out, tag):
	"""Per-kernel counters of a `ncu --set full` report; also writes profiles/ncu_constants.json -- the per-launch DRAM traffic and
	pipe utilisation of the four persistent kernels that bench.py quotes in its `roofline` / `rooflines_kernels` blocks."""
	# `rep`: an .ncu-rep, or the CSV of its raw page (`ncu -i x.ncu-rep --page raw --csv`, made on the GPU box when the report is too large to travel)
	txt = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
	rows = list(csv.reader(txt.splitlines()))
	hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
	units = rows[1]
	consts = {'_source': 'profiles/{}.txt (ncu --set full --clock-control none; one launch per kernel)'.format(tag)}
	for r in rows[2:]:
		name = r[idx['Kernel Name']]
		out.write('\nkernel: {}\n'.format(name))
		for k in KEYS:
			if k in idx:
				out.write('  {:<72s} {} {}\n'.format(k, r[idx[k]], units[idx[k]]))
		for pat, group in GROUPS:
			if pat in name and group not in consts:
				def val(k):
					return float(r[idx[k]].replace(',', '')) * SCALE.get(units[idx[k]], 1.0)
				consts[group] = {'kernel': re.sub(r'\(.*', '', name).replace('void <unnamed>::', ''),
								'dram_bytes': val('dram__bytes_read.sum') + val('dram__bytes_write.sum'),
								'tensor_pct': val('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
								'fma_pct': val('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'),
								'issue_active_pct': val('smsp__issue_active.avg.pct_of_peak_sustained_active'),
								'duration_ms_under_ncu': val('gpu__time_duration.sum') * (1e-6 if units[idx['gpu__time_duration.sum']] in ('ns', 'nsecond') else
																						1e-3 if units[idx['gpu__time_duration.sum']] in ('us', 'usecond') else 1.0)}
	if len(consts) > 1:
		json.dump(consts, open('profiles/ncu_constants.json', 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
	ll, rep, tag = sys.argv[1:4]
	with open('profiles/{}.txt'.format(tag), 'w') as f:
		f.write('# {} -- summarised by tools/summarize_profiles.py\n'.format(tag))
		if ll != '-':
			launches(ll, f)
		if rep != '-':
			raw(rep, f, tag)
		for extra in sys.argv[4:]:   # further captures of the same run (e.g. the GEMM kernels); they do not touch the constants file
			keep = open('profiles/ncu_constants.json').read() if os.path.exists('profiles/ncu_constants.json') else None
			raw(extra, f, tag)
			if keep is not None:
				open('profiles/ncu_constants.json', 'w').write(keep)
	print(open('profiles/{}.txt'.format(tag)).read()[:3000])
