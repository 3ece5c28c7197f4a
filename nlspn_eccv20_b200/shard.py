"""Batch sharding for multi-GPU runs.

Every batch image propagates independently (the gather indexes only its own plane, the blend is
elementwise, there is no batch statistic on the path), so the path shards by contiguous batch
ranges, one process per GPU, with NO collective on the data path.  The only cross-rank quantities
are reporting ones: error sums for RMSE/MAE and, in full-model training, the gradient of the scalar
gamma, which rides in DDP's ordinary all-reduce.
"""
from __future__ import annotations

import torch


def shard_range(total: int, world: int, rank: int):
    """Contiguous shard [start, start+count) of `total` images for `rank` of `world`.
    The first ``total % world`` ranks get one extra image (reference: batch_size // num_gpus per
    process, src/main.py:98, which silently drops the remainder; here nothing is dropped)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank: %d/%d" % (world, rank))
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard_inputs(inputs: dict, world: int, rank: int):
    """Slice every tensor of `inputs` along dim 0 to this rank's shard (views, no copy)."""
    B = next(v for v in inputs.values() if torch.is_tensor(v)).shape[0]
    s, n = shard_range(B, world, rank)
    return {k: (v[s:s + n] if torch.is_tensor(v) else v) for k, v in inputs.items()}


def error_sums(pred, gt, t_valid=1e-4):
    """Per-shard sufficient statistics of RMSE / MAE (src/metric/nlspnmetric.py:40,53-60):
    -> tensor [sum of squared error, sum of absolute error, number of valid pixels] (float64)."""
    mask = gt > t_valid
    d = (pred[mask] - gt[mask]).double()
    return torch.stack([(d * d).sum(), d.abs().sum(), mask.sum().double()])


def reduce_rmse_mae(sums, group=None):
    """All-reduce the sufficient statistics (SUM) and return (rmse, mae) of the whole batch."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        sums = sums.clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    sq, ab, n = (float(x) for x in sums.tolist())
    return (sq / (n + 1e-8)) ** 0.5, ab / (n + 1e-8)
