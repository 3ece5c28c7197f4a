"""CPU, world_size 2 over gloo: the N>1 host path -- batch sharding and the metric reduction --
gives exactly the single-process result.  The per-shard compute here is the CPU oracle (test
infrastructure standing in for the GPU kernel); on the GPU box bench.py runs the same sharding
with the CUDA path (one process per GPU, NCCL only for the timing/metric reduction)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nlspn_eccv20_b200.shard import shard_inputs, shard_range, error_sums, reduce_rmse_mae
    from nlspn_eccv20_b200.synth import make_inputs
    from oracle import nlspn_oracle as O
    B, H, W, K, T = 5, 20, 26, 3, 4                      # 5 images over 2 ranks: ragged shards
    full = make_inputs(B, H, W, K, seed=42, conf_mean=3.0)
    mine = shard_inputs(full, world, rank)
    s, n = shard_range(B, world, rank)
    assert mine["feat_init"].shape[0] == n
    out = O.nlspn_forward(mine["feat_init"].numpy(), mine["guidance"].numpy(), mine["confidence"].numpy(),
                          mine["feat_fix"].numpy(), 4.0, K, T)
    pred = torch.from_numpy(out["feat_result"]).clamp(min=0)
    rmse, mae = reduce_rmse_mae(error_sums(pred, mine["gt"]))
    # gather the shards on rank 0 for the exactness check
    parts = [None] * world
    dist.all_gather_object(parts, (s, n, out["feat_result"]))
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"),
                np.concatenate([p[2] for p in sorted(parts, key=lambda p: p[0])], 0))
        np.save(os.path.join(out_dir, "metrics.npy"), np.array([rmse, mae]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from nlspn_eccv20_b200.shard import shard_range
    for total in (0, 1, 5, 16, 64):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert sum(n for _, n in spans) == total
            pos = 0
            for s, n in spans:
                assert s == pos
                pos += n
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_rank_gloo_sharded_equals_unsharded(tmp_path, oracle):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from nlspn_eccv20_b200.shard import error_sums
    from nlspn_eccv20_b200.synth import make_inputs
    full = make_inputs(5, 20, 26, 3, seed=42, conf_mean=3.0)
    ref = oracle.nlspn_forward(full["feat_init"].numpy(), full["guidance"].numpy(), full["confidence"].numpy(),
                               full["feat_fix"].numpy(), 4.0, 3, 4)
    got = np.load(tmp_path / "gathered.npy")
    assert np.array_equal(got, ref["feat_result"])          # images are independent: bit-identical
    sq, ab, n = error_sums(torch.from_numpy(ref["feat_result"]).clamp(min=0), full["gt"]).tolist()
    rmse, mae = np.load(tmp_path / "metrics.npy")
    assert abs(rmse - (sq / n) ** 0.5) < 1e-9 and abs(mae - ab / n) < 1e-9
