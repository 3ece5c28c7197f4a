"""TEST INFRASTRUCTURE / CPU BASELINE ONLY -- the reference path restated over torchvision.

BASELINE.json's north_star names ``torchvision.ops.deform_conv2d(mask=...)`` as the CPU
stand-in for the reference's CUDA-only DCNv2 extension.  This module restates, without
importing anything from /root/reference (which does not exist on the GPU box):

  * ``DeformStep``  -- ModulatedDeformConvFunction (src/model/modulated_deform_conv_func.py:15-56)
                       over torch.ops.torchvision.deform_conv2d / _deform_conv2d_backward, with the
                       reference's "coordinate <= -1 => zero offset gradient" rule
                       (modulated_deform_im2col_cuda.cuh:88-92,308-311) applied to the stand-in.
  * ``propagate``   -- nlspnmodel.py:323-377 (offset insertion, TGASS/AS/ASS/TC normalisation,
                       mask/confidence fix-up, T x {premultiply, gather, blend, clip}).

Parity status: PINNED by tests/test_oracle_golden.py::test_torchvision_port_* against the golden
vectors produced by the unmodified reference (oracle/gen_golden.py).
Only tests/ and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import os
import time
from concurrent.futures import ThreadPoolExecutor

import torch
import torchvision  # noqa: F401  registers torch.ops.torchvision


def _keep_mask(offset, K, pad):
    """1 where the reference keeps grad_offset: both sampling coordinates > -1."""
    B, _, H, W = offset.shape
    dt = offset.dtype
    hs = torch.arange(H).view(1, 1, H, 1) - pad
    ws = torch.arange(W).view(1, 1, 1, W) - pad
    keep = torch.empty_like(offset)
    for t in range(K * K):
        i, j = divmod(t, K)
        h_im = (hs + i).to(dt) + offset[:, 2 * t:2 * t + 1]      # int first, one float add
        w_im = (ws + j).to(dt) + offset[:, 2 * t + 1:2 * t + 2]
        ok = ((h_im > -1) & (w_im > -1)).to(dt)
        keep[:, 2 * t:2 * t + 1] = ok
        keep[:, 2 * t + 1:2 * t + 2] = ok
    return keep


class DeformStep(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias, K):
        pad = (K - 1) // 2
        ctx.K, ctx.pad = K, pad
        out = torch.ops.torchvision.deform_conv2d(x, weight, offset, mask, bias, 1, 1, pad, pad,
                                                  1, 1, 1, 1, True)
        ctx.save_for_backward(x, offset, mask, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, offset, mask, weight, bias = ctx.saved_tensors
        pad = ctx.pad
        gi, gw, go, gm, gb = torch.ops.torchvision._deform_conv2d_backward(
            g.contiguous(), x, weight, offset, mask, bias, 1, 1, pad, pad, 1, 1, 1, 1, True)
        go = go * _keep_mask(offset, ctx.K, pad)
        return gi, go, gm, gw, gb, None


def normalize_affinity(aff, gamma, affinity):
    """nlspnmodel.py:179-201 + :261-269 (without the centre insertion's list juggling)."""
    if affinity == "TC":
        aff = torch.tanh(aff) / gamma
    elif affinity == "TGASS":
        aff = torch.tanh(aff) / (gamma + 1e-8)
    abs_sum = aff.abs().sum(1, keepdim=True) + 1e-4
    if affinity in ("ASS", "TGASS"):
        abs_sum = torch.where(abs_sum < 1.0, torch.ones_like(abs_sum), abs_sum)
    if affinity in ("AS", "ASS", "TGASS"):
        aff = aff / abs_sum
    ref = 1.0 - aff.sum(1, keepdim=True)
    n = aff.shape[1] // 2
    return torch.cat([aff[:, :n], ref, aff[:, n:]], 1)


def insert_center_offset(off, K):
    """nlspnmodel.py:252-259."""
    B, _, H, W = off.shape
    N = K * K - 1
    o = off.view(B, N, 2, H, W)
    z = torch.zeros(B, 1, 2, H, W, dtype=off.dtype)
    return torch.cat([o[:, :N // 2], z, o[:, N // 2:]], 1).reshape(B, 2 * K * K, H, W)


def fixed_local_step(x, aff):
    """nlspnmodel.py:209-224: replicate-pad by one, weight the 9 shifted copies, sum."""
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode="replicate")
    H, W = x.shape[2], x.shape[3]
    cols = [xp[:, :, i:i + H, j:j + W] for i in range(3) for j in range(3)]
    return (torch.cat(cols, 1) * aff).sum(1, keepdim=True)


class DeformStep1x1(torch.autograd.Function):
    """1x1 modulated deformable gather, pad 0 (the upstream conf_prop sampling call)."""

    @staticmethod
    def forward(ctx, x, offset, mask, weight, bias):
        out = torch.ops.torchvision.deform_conv2d(x, weight, offset, mask, bias, 1, 1, 0, 0, 1, 1, 1, 1, True)
        ctx.save_for_backward(x, offset, mask, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x, offset, mask, weight, bias = ctx.saved_tensors
        gi, gw, go, gm, gb = torch.ops.torchvision._deform_conv2d_backward(
            g.contiguous(), x, weight, offset, mask, bias, 1, 1, 0, 0, 1, 1, 1, 1, True)
        return gi, None, None, None, None          # offsets are detached upstream; mask/w/b are constants


def upstream_affinity(aff_raw, offset, confidence, gamma, K, affinity, legacy):
    """UPSTREAM (zzangjinsun/NLSPN_ECCV20) affinity path -- *parity unpinned*: that repository is
    not under /root/reference; this restates BASELINE.json's north-star prose ("conf_prop confidence
    bilinear sampling at the predicted offsets", SURVEY 0.2 right-hand column) from the public
    upstream code as recalled: tanh/gamma, multiply each neighbour's affinity by the confidence
    sampled with a 1x1 deformable gather at that neighbour's (detached) offset -- the tap
    displacement is added to the offset only with --legacy -- then abs-sum normalisation."""
    N = K * K - 1
    if affinity == "TC":
        aff = torch.tanh(aff_raw) / gamma
    elif affinity == "TGASS":
        aff = torch.tanh(aff_raw) / (gamma + 1e-8)
    else:
        aff = aff_raw
    if confidence is not None:
        B, _, H, W = aff.shape
        ones = torch.ones(B, 1, H, W, dtype=aff.dtype)
        w1 = torch.ones(1, 1, 1, 1, dtype=aff.dtype)
        b0 = torch.zeros(1, dtype=aff.dtype)
        confs = []
        pad = (K - 1) // 2
        for t in range(K * K):
            hh, ww = divmod(t, K)
            if hh == pad and ww == pad:
                continue
            off = offset[:, 2 * t:2 * t + 2].detach().clone()
            if legacy:
                off[:, 0] = off[:, 0] + (hh - pad)
                off[:, 1] = off[:, 1] + (ww - pad)
            confs.append(DeformStep1x1.apply(confidence, off.contiguous(), ones, w1, b0))
        aff = aff * torch.cat(confs, 1)
    abs_sum = aff.abs().sum(1, keepdim=True) + 1e-4
    if affinity in ("ASS", "TGASS"):
        abs_sum = torch.where(abs_sum < 1.0, torch.ones_like(abs_sum), abs_sum)
    if affinity in ("AS", "ASS", "TGASS"):
        aff = aff / abs_sum
    ref = 1.0 - aff.sum(1, keepdim=True)
    return torch.cat([aff[:, :N // 2], ref, aff[:, N // 2:]], 1)


def propagate_upstream(feat_init, guidance, confidence, feat_fix, gamma, K, T, affinity="TGASS",
                       preserve_input=True, legacy=False):
    """Upstream loop: blend BEFORE every iteration only, no confidence pre-multiply, last output not
    blended.  -> dict(feat_result, list_feat, offset, aff)."""
    N = K * K - 1
    offset = insert_center_offset(guidance[:, :2 * N], K)
    aff = upstream_affinity(guidance[:, 2 * N:], offset, confidence, gamma, K, affinity, legacy)
    preserve = preserve_input and feat_fix is not None
    if preserve:
        m = (feat_fix > 0).to(feat_init.dtype)
    w = torch.ones(1, 1, K, K, dtype=feat_init.dtype)
    b = torch.zeros(1, dtype=feat_init.dtype)
    x = feat_init
    out = []
    for _ in range(T):
        if preserve:
            x = (1.0 - m) * x + m * feat_fix
        x = DeformStep.apply(x, offset, aff, w, b, K)
        out.append(x)
    return dict(feat_result=x, list_feat=out, offset=offset, aff=aff, confidence=None)


def propagate(feat_init, guidance, confidence, feat_fix, gamma, K, T, affinity="TGASS",
              preserve_input=True, always_clip=False, use_offset=True):
    """-> dict(feat_result, list_feat (list of T), offset, aff, confidence)."""
    N = K * K - 1
    if use_offset:
        offset = insert_center_offset(guidance[:, :2 * N], K)
        aff = normalize_affinity(guidance[:, 2 * N:], gamma, affinity)
    else:
        offset = None
        aff = normalize_affinity(guidance, gamma, affinity)
    preserve = preserve_input and feat_fix is not None
    if preserve:
        m = (feat_fix > 0).to(feat_init.dtype)
        if confidence is not None:
            confidence = (1.0 - m) * confidence + m
    w = torch.ones(1, 1, K, K, dtype=feat_init.dtype)
    b = torch.zeros(1, dtype=feat_init.dtype)
    x = feat_init
    if preserve:
        x = (1.0 - m) * x + m * feat_fix
    if always_clip:
        x = torch.clamp(x, min=0)
    out = []
    for _ in range(T):
        s = x * confidence if confidence is not None else x
        x = DeformStep.apply(s, offset, aff, w, b, K) if use_offset else fixed_local_step(s, aff)
        if preserve:
            x = (1.0 - m) * x + m * feat_fix
        if always_clip:
            x = torch.clamp(x, min=0)
        out.append(x)
    return dict(feat_result=x, list_feat=out, offset=offset, aff=aff, confidence=confidence)


# ---------------------------------------------------------------------------------------------
# CPU baseline timing: torchvision's CPU kernels do not scale with intra-op threads,
# so the fair "all host cores" figure runs one image per worker thread (results are bit-identical
# to the batched call).
# ---------------------------------------------------------------------------------------------
def _one_image(args):
    inp, gamma, K, T, backward = args
    torch.set_num_threads(1)
    fi = inp["feat_init"].clone().requires_grad_(backward)
    gd = inp["guidance"].clone().requires_grad_(backward)
    cf = inp["confidence"].clone().requires_grad_(backward)
    gam = torch.tensor([gamma], requires_grad=backward)
    if backward:
        out = propagate(fi, gd, cf, inp["feat_fix"], gam, K, T)
        loss = torch.clamp(out["feat_result"], min=0).sum()
        loss.backward()
    else:
        with torch.no_grad():
            out = propagate(fi, gd, cf, inp["feat_fix"], gam, K, T)
    return float(out["feat_result"].sum())


def time_batch_parallel(images, gamma, K, T, backward=True, workers=None):
    """images: list of per-image input dicts ([1,...] tensors).  Returns wall seconds."""
    workers = workers or os.cpu_count() or 1
    jobs = [(im, gamma, K, T, backward) for im in images]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(_one_image, jobs))
    return time.perf_counter() - t0
