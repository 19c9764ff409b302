# coding: utf-8
"""Multi-GPU correctness check of Learner.train_step (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/check_dist_gpu.py

Every rank trains two steps on its strided shard of one global batch (dist.shard_packed, loss divided by the GLOBAL batch
size, gradient all-reduce overlapped with the encoder backward); rank 0 also runs the same two steps on the full batch with
a single-process Learner.  The updated parameters must agree to fp32 round-off (the summation order differs)."""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
	local = int(os.environ.get('LOCAL_RANK', rank))
	torch.cuda.set_device(local)
	dev = torch.device('cuda', local)
	torch.distributed.init_process_group('nccl', device_id=dev)
	pkg = importlib.import_module('seq2seq_abcd-vae_b200')
	learning = importlib.import_module('seq2seq_abcd-vae_b200.learning')
	dist = importlib.import_module('seq2seq_abcd-vae_b200.dist')
	from oracle import cases as CS
	from oracle import abcd_oracle as O
	cfg = dict(rnn='LSTM', F=65, K=128, f=256, lengths=sorted([(i * 7) % 19 + 1 for i in range(96)], reverse=True), pretrain=False,
				tau=0.9, N_data=5000, seed=9)

	def make_learner():
		ln = learning.Learner(cfg['F'], 256, 256, 256, cfg['K'], cfg['f'], None, device=str(dev), seed=3)
		mods = ln._modules()
		sds = CS.build_case(cfg, mods)[:3]
		for m, sd in zip(mods, sds):
			m.load_state_dict({k: v.to(dev) for k, v in sd.items()})
		ln.setup_optimizer(0.5, 0.0, gradient_clip=1.0, patience=0)
		ln.feature_sampler.train()
		ln.feature_sampler.temperature = cfg['tau']
		return ln

	ln = make_learner()
	_, _, _, seqs, eps, gumbel, _ = CS.build_case(cfg, ln._modules())
	packed = torch.nn.utils.rnn.pack_sequence(seqs)
	is_off = O.is_offset_packed(packed.batch_sizes.tolist())
	B = len(seqs)
	data_s, bs_s = dist.shard_packed(packed.data, packed.batch_sizes, rank, world)
	off_s, _ = dist.shard_packed(is_off, packed.batch_sizes, rank, world)
	eps_s, _ = dist.shard_packed(eps, packed.batch_sizes, rank, world)
	gum_s = gumbel[dist.shard_indices(B, rank, world)]
	p_in = torch.nn.utils.rnn.PackedSequence(data_s.to(dev), bs_s)
	p_off = torch.nn.utils.rnn.PackedSequence(off_s.to(dev), bs_s)
	def say(msg):
		print('[rank {}] {}'.format(rank, msg), flush=True)

	phases = os.environ.get('ABCD_DIST_PHASES', 'shard,eager,graph').split(',')
	say('set-up done')
	for i in range(2):
		ln.feature_sampler.injected_gumbel_noise = gum_s.to(dev)
		ln.decoder.injected_emission_noise = eps_s.to(dev)
		ln.train_step(p_in, p_off, None, cfg['N_data'], False, global_batch_size=B)
		torch.cuda.synchronize()
		say('sharded step {} done'.format(i))
	ok = True
	if rank == 0:
		ref = make_learner()
		ref.world_size = 1   # full batch, no collective
		f_in = torch.nn.utils.rnn.PackedSequence(packed.data.to(dev), packed.batch_sizes)
		f_off = torch.nn.utils.rnn.PackedSequence(is_off.to(dev), packed.batch_sizes)
		for _ in range(2):
			ref.feature_sampler.injected_gumbel_noise = gumbel.to(dev)
			ref.decoder.injected_emission_noise = eps.to(dev)
			ref.train_step(f_in, f_off, None, cfg['N_data'], False)
		worst = 0.0
		for (k, a), (_, b) in zip(sorted(dict(ln.encoder.named_parameters(), **{'s.' + k: v for k, v in ln.feature_sampler.named_parameters()},
												**{'d.' + k: v for k, v in ln.decoder.named_parameters()}).items()),
									sorted(dict(ref.encoder.named_parameters(), **{'s.' + k: v for k, v in ref.feature_sampler.named_parameters()},
												**{'d.' + k: v for k, v in ref.decoder.named_parameters()}).items())):
			err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
			worst = max(worst, err)
			if err > 1e-5:
				ok = False
				print('MISMATCH', k, err)
		print('dist check: world {} worst relative parameter difference after 2 steps {:.2e} -> {}'.format(world, worst, 'OK' if ok else 'FAIL'))
	# all ranks must hold identical parameters
	flat = ln.arena.flat.clone()
	torch.distributed.broadcast(flat, 0)
	same = bool(torch.equal(flat, ln.arena.flat))
	print('rank {} parameters identical to rank 0: {}'.format(rank, same))

	# in-kernel (Philox) noise, one stream per rank; the same 5 sharded steps host-launched and as CUDA-graph replays (NCCL
	# all-reduces recorded inside the graph, decoder weight gradients on the side stream) must give bit-identical parameters
	finals = []
	for use_graph in (False, True):
		if ('graph' if use_graph else 'eager') not in phases:
			finals.append((ln.arena.flat.clone(), 'skipped'))
			continue
		l2 = make_learner()
		pkg.model.manual_noise_seed(77, stream=rank)
		l2.use_cuda_graph = use_graph
		for i in range(5):
			l2.train_step(p_in, p_off, None, cfg['N_data'], False, global_batch_size=B)
			torch.cuda.synchronize()
			say('{} step {} done'.format('graph' if use_graph else 'eager', i))
		finals.append((l2.arena.flat.clone(), l2.launch_mode()))
		l2.release_graphs()
		del l2
	graph_same = bool(torch.equal(finals[0][0], finals[1][0])) or 'skipped' in (finals[0][1], finals[1][1])
	bc = finals[1][0].clone()
	torch.distributed.broadcast(bc, 0)
	graph_ranks_same = bool(torch.equal(bc, finals[1][0]))
	print('rank {} graph replay == host-launched: {} ({}); ranks identical: {}'.format(rank, graph_same, finals[1][1], graph_ranks_same))
	torch.distributed.barrier()
	torch.distributed.destroy_process_group()
	sys.exit(0 if (ok and same and graph_same and graph_ranks_same) else 1)


if __name__ == '__main__':
	main()
