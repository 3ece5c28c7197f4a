"""Synthetic workloads of the propagation path (SURVEY 8d "stable set" / "signed set").

No dataset ships with the reference (.MISSING_LARGE_BLOBS), so bench.py and the tests use
seeded synthetic tensors with the shapes, value ranges and sparse-depth densities of the
reference's data pipeline (src/config.py:124-141, src/data/nyu.py:164-182).  Generated on the
CPU generator (device-independent streams) and moved to the requested device, so the CPU
oracle and the GPU path see bit-identical inputs.
"""
from __future__ import annotations

import torch

SEED = 7240  # reference default seed, src/config.py:58-61

SHAPES = {"nyu": (228, 304, 10.0), "kitti": (352, 1216, 90.0)}


def smooth_field(g, B, H, W, lo, hi, cell=32):
    gh, gw = max(2, H // cell + 1), max(2, W // cell + 1)
    grid = lo + (hi - lo) * torch.rand(B, 1, gh, gw, generator=g)
    f = torch.nn.functional.interpolate(grid, size=(H, W), mode="bicubic", align_corners=True)
    return f.clamp(lo, hi)


def make_inputs(B, H, W, K, max_depth=10.0, seed=SEED, signed=False, conf_mean=0.0,
                off_sigma=2.0, smooth_offsets=False, density=None, num_sample=None,
                device="cpu", pin=False):
    """-> dict(feat_init, guidance, confidence, feat_fix, gt, rgb) of fp32 tensors.

    stable set: non-negative raw affinities (convex update); signed=True gives the
    robustness set.  conf_mean=0 -> sigmoid(N(0,1)) ("random confidence", the timing set);
    conf_mean=3 -> the parity set.  num_sample: exactly that many fixed pixels per image
    (NYU: 500, config.py:137-141); density: Bernoulli (KITTI LiDAR-like, 0.05)."""
    g = torch.Generator().manual_seed(int(seed))
    N = K * K - 1
    gt = smooth_field(g, B, H, W, 0.5, max_depth)
    if signed:
        feat_init = max_depth * torch.rand(B, 1, H, W, generator=g)
        aff_raw = torch.randn(B, N, H, W, generator=g)
    else:
        feat_init = (gt + 0.005 * max_depth * torch.randn(B, 1, H, W, generator=g)).clamp(min=0)
        aff_raw = torch.randn(B, N, H, W, generator=g).abs()
    off_raw = off_sigma * torch.randn(B, 2 * N, H, W, generator=g)
    if smooth_offsets:
        off_raw = torch.nn.functional.avg_pool2d(off_raw, 9, stride=1, padding=4) * 9.0 / 3.0
    guidance = torch.cat([off_raw, aff_raw], 1).contiguous()
    confidence = torch.sigmoid(conf_mean + torch.randn(B, 1, H, W, generator=g))
    if num_sample is not None:
        mask = torch.zeros(B, H * W)
        for b in range(B):
            idx = torch.randperm(H * W, generator=g)[:num_sample]
            mask[b, idx] = 1.0
        mask = mask.view(B, 1, H, W)
    else:
        mask = (torch.rand(B, 1, H, W, generator=g) < (0.05 if density is None else density)).float()
    feat_fix = gt * mask
    rgb = None
    out = dict(feat_init=feat_init, guidance=guidance, confidence=confidence, feat_fix=feat_fix, gt=gt)
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    if str(device) != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    out["rgb"] = rgb
    return out


def workload(name, B, K=3, **kw):
    H, W, md = SHAPES[name]
    if name == "nyu":
        kw.setdefault("num_sample", 500)
    else:
        kw.setdefault("density", 0.05)
    return make_inputs(B, H, W, K, max_depth=md, **kw)


def rmse_mae(pred, gt, t_valid=1e-4):
    """RMSE / MAE with the reference's masking (src/metric/nlspnmetric.py:25,40,53-60)."""
    mask = gt > t_valid
    n = mask.sum()
    diff = pred[mask] - gt[mask]
    rmse = torch.sqrt((diff ** 2).sum() / (n + 1e-8))
    mae = diff.abs().sum() / (n + 1e-8)
    return float(rmse), float(mae)
