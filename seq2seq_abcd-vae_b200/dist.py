# coding: utf-8
"""Segment-sharded data parallelism (one process per GPU, torch.distributed over NCCL/NVLink).

Every loss term of the ABCD-VAE is a SUM over segments (emission / offset sums over frames, the per-item KL term
over b, and the Dirichlet prior term scaled by local_batch / entire_data_size, reference model.py:636-638), so a
batch shards exactly: each rank runs forward/backward on its slice, divides by the GLOBAL batch size, and the
flat gradient arenas are summed with ONE all-reduce before the (replicated, deterministic) clip + SGD step.

Sharding is strided over the length-sorted batch: rank r takes elements r, r+R, r+2R, ... so every shard stays
sorted (a valid PackedSequence), sees the same length distribution and runs the same number of time steps.
"""
import numpy as np
import torch


def shard_indices(n, rank, world_size):
	"""Positions (in the length-sorted batch) owned by `rank`."""
	return np.arange(rank, n, world_size)


def shard_sorted_batch(sorted_lengths, rank, world_size):
	"""(positions, lengths) of this rank's strided shard of a batch whose lengths are sorted descending."""
	pos = shard_indices(len(sorted_lengths), rank, world_size)
	return pos, np.asarray(sorted_lengths)[pos]


def shard_packed(data, batch_sizes, rank, world_size):
	"""Slice a PackedSequence's (data, batch_sizes) into this rank's strided shard, still time-major packed.
	Row (t, b) of the full batch lives at offsets[t] + b; the shard keeps b = rank, rank+R, ... < batch_sizes[t]."""
	bs = [int(v) for v in batch_sizes]
	rows = []
	new_bs = []
	off = 0
	for n in bs:
		keep = list(range(rank, n, world_size))
		if keep:
			rows.extend(off + b for b in keep)
			new_bs.append(len(keep))
		off += n
	idx = torch.as_tensor(rows, dtype=torch.long, device=data.device)
	return data.index_select(0, idx), torch.tensor(new_bs, dtype=torch.int64)


def all_reduce_sum_(flat, group=None):
	"""In-place SUM all-reduce of the flat gradient arena (no averaging: losses are already divided by the
	global batch size)."""
	if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
		torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
	return flat
