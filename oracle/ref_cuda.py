"""TEST INFRASTRUCTURE -- the reference propagation path on the GPU through the reference's
*own* CUDA kernels (oracle/_ref/DCN_ref.so, the "patched build" of
/root/reference/src/model/deformconv/src/cuda/modulated_deform_conv_cuda.cu made by
oracle/build_ref_cuda.py).

Only the native kernels travel to the GPU box (as a compiled module); the reference's Python
does not, so the two thin Python layers above them are restated here:

  * ``RefDeformStep``  -- ModulatedDeformConvFunction, src/model/modulated_deform_conv_func.py:15-56
                          (same argument order, same saved tensors, the native call with
                          input, weight, bias, offset, mask -- vision.cpp:9-10).
  * ``propagate``      -- nlspnmodel.py:323-377 with the module logic of
                          oracle/torchvision_port.py (pinned against the unmodified reference by
                          tests/test_oracle_golden.py), device-aware.

Only tests/ and tests/perf_reference_cuda.py may import this; the product never does.
"""
from __future__ import annotations

import importlib.util
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "DCN_ref.so")
_mod = None


def available() -> bool:
    return os.path.exists(SO)


def load():
    global _mod
    if _mod is None:
        if not available():
            raise FileNotFoundError(SO + " (build it with python oracle/build_ref_cuda.py where /root/reference exists)")
        spec = importlib.util.spec_from_file_location("DCN_ref", SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


class RefDeformStep(torch.autograd.Function):
    """modulated_deform_conv_func.py:15-56 over the reference's native CUDA entry points."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, K):
        ctx.K = K
        pad = (K - 1) // 2
        out = load().modulated_deform_conv_forward(input, weight, bias, offset, mask,
                                                   K, K, 1, 1, pad, pad, 1, 1, 1, 1, 64)
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        K = ctx.K
        pad = (K - 1) // 2
        gi, go, gm, gw, gb = load().modulated_deform_conv_backward(
            input, weight, bias, offset, mask, grad_output.contiguous(),
            K, K, 1, 1, pad, pad, 1, 1, 1, 1, 64)
        return gi, go, gm, gw, gb, None


def _insert_center_offset(off, K):
    B, _, H, W = off.shape
    N = K * K - 1
    o = off.view(B, N, 2, H, W)
    z = torch.zeros(B, 1, 2, H, W, dtype=off.dtype, device=off.device)
    return torch.cat([o[:, :N // 2], z, o[:, N // 2:]], 1).reshape(B, 2 * K * K, H, W)


def propagate(feat_init, guidance, confidence, feat_fix, gamma, K, T, affinity="TGASS",
              preserve_input=True, always_clip=False):
    """nlspnmodel.py:323-377, offsets on (the DCN path).  All tensors on one CUDA device."""
    from oracle.torchvision_port import normalize_affinity
    N = K * K - 1
    offset = _insert_center_offset(guidance[:, :2 * N], K)
    aff = normalize_affinity(guidance[:, 2 * N:], gamma, affinity)
    preserve = preserve_input and feat_fix is not None
    if preserve:
        m = (feat_fix > 0).to(feat_init.dtype)
        if confidence is not None:
            confidence = (1.0 - m) * confidence + m
    w = torch.ones(1, 1, K, K, dtype=feat_init.dtype, device=feat_init.device)
    b = torch.zeros(1, dtype=feat_init.dtype, device=feat_init.device)
    x = feat_init
    if preserve:
        x = (1.0 - m) * x + m * feat_fix
    if always_clip:
        x = torch.clamp(x, min=0)
    out = []
    for _ in range(T):
        s = x * confidence if confidence is not None else x
        x = RefDeformStep.apply(s.contiguous(), offset, aff, w, b, K)
        if preserve:
            x = (1.0 - m) * x + m * feat_fix
        if always_clip:
            x = torch.clamp(x, min=0)
        out.append(x)
    return dict(feat_result=x, list_feat=out, offset=offset, aff=aff, confidence=confidence)
