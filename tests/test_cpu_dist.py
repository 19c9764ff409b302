# coding: utf-8
"""CPU suite, part 3: the N>1 path on 2 gloo ranks -- strided sharding of a packed batch, loss / global-batch
scaling, SUM all-reduce of flat gradient arenas.  The per-rank compute is the CPU oracle (the CUDA kernels need a
GPU); what is under test is the host-side sharding logic and the collective pattern Learner.train_step uses."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import abcd_oracle as O
from oracle import cases as CS


def _free_port():
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	return port


def _worker(rank, world, port, out_dir):
	os.environ['MASTER_ADDR'] = '127.0.0.1'
	os.environ['MASTER_PORT'] = str(port)
	torch.distributed.init_process_group('gloo', rank=rank, world_size=world)
	torch.set_num_threads(2)
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	dist = importlib.import_module('seq2seq_abcd-vae_b200.dist')
	cfg = CS.TRAIN_CASES['train_lstm_gumbel']
	mods = CS.build_modules(pkg.model, cfg)
	enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
	dt = torch.float64
	cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
	B = len(seqs)
	data, bs, _ = O.pack(seqs)
	# shard: sequences, their packed emission noise, their Gumbel rows
	pos = dist.shard_indices(B, rank, world)
	my_seqs = [seqs[i].to(dt) for i in pos]
	my_eps, my_bs = dist.shard_packed(eps.to(dt), bs, rank, world)
	my_data, my_bs2 = dist.shard_packed(data, bs, rank, world)
	assert my_bs.tolist() == my_bs2.tolist() == O.batch_sizes_from_lengths([s.shape[0] for s in my_seqs])
	assert torch.equal(my_data, O.pack([seqs[i] for i in pos])[0])   # shard of a packed batch == packed shard
	r = O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), my_seqs, my_eps, gumbel.to(dt)[pos], cfg['N_data'],
					cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, None, global_batch=B)
	names, flat = [], []
	for m in ('encoder', 'feature_sampler', 'decoder'):
		for k in sorted(r['grads'][m]):
			names.append((m, k)); flat.append(r['grads'][m][k].reshape(-1))
	arena = torch.cat(flat)
	dist.all_reduce_sum_(arena)
	loss = r['loss'].clone()
	torch.distributed.all_reduce(loss)
	if rank == 0:
		torch.save({'arena': arena, 'loss': loss, 'names': names}, os.path.join(out_dir, 'reduced.pt'))
	torch.distributed.destroy_process_group()


def test_two_rank_sharded_step_equals_full_batch(abcd, tmp_path):
	world = 2
	mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
	red = torch.load(os.path.join(str(tmp_path), 'reduced.pt'), weights_only=False)
	cfg = CS.TRAIN_CASES['train_lstm_gumbel']
	mods = CS.build_modules(abcd.model, cfg)
	enc_sd, smp_sd, dec_sd, seqs, eps, gumbel, speaker = CS.build_case(cfg, mods)
	dt = torch.float64
	cast = lambda d: {k: (v.to(dt) if v.dtype.is_floating_point else v) for k, v in d.items()}
	full = O.train_step(cast(enc_sd), cast(smp_sd), cast(dec_sd), [s.to(dt) for s in seqs], eps.to(dt), gumbel.to(dt),
						cfg['N_data'], cfg['pretrain'], cfg['tau'], cfg['rnn'], 1.0, None)
	flat = torch.cat([full['grads'][m][k].reshape(-1) for m, k in red['names']])
	assert abs(red['loss'].item() - full['loss'].item()) < 1e-9 * abs(full['loss'].item())
	assert float((red['arena'] - flat).abs().max()) < 1e-10 * float(flat.abs().max())


def test_shard_helpers(abcd):
	dist = importlib.import_module('seq2seq_abcd-vae_b200.dist')
	lengths = np.array([9, 9, 7, 4, 4, 2, 1])
	parts = [dist.shard_sorted_batch(lengths, r, 3) for r in range(3)]
	assert sorted(np.concatenate([p for p, _ in parts]).tolist()) == list(range(7))
	for _, l in parts:
		assert (np.diff(l) <= 0).all()       # every shard stays sorted -> a valid PackedSequence
	bs = O.batch_sizes_from_lengths(lengths.tolist())
	data = torch.arange(sum(bs)).float()[:, None]
	total = 0
	for r in range(3):
		d, b = dist.shard_packed(data, bs, r, 3)
		assert b.tolist() == O.batch_sizes_from_lengths(parts[r][1].tolist())
		total += d.shape[0]
	assert total == sum(bs)
