"""Thin torch-facing wrappers over the C ABI: tensors in, tensors out, current stream.

torch is plumbing here (device memory, streams); all arithmetic happens in libnlspn_b200.so.
Every wrapper raises if a tensor is not a contiguous fp32 CUDA tensor -- there is no
CPU path.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

__all__ = ["forward", "prologue_fwd", "propagate_fwd", "backward", "dcn_forward", "dcn_backward",
           "step_fwd", "step_bwd", "debug_indices", "device_info"]


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _chk(name, t, shape=None, optional=False, dtype=torch.float32):
    if t is None:
        if optional:
            return None
        raise RuntimeError("%s is required" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (nlspn_eccv20_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s (got %s)" % (name, str(dtype).replace("torch.", ""), t.dtype))
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise RuntimeError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return t.contiguous()


def _gamma(gamma, dev):
    """gamma as a 1-element fp32 device tensor (no host sync when it already is one)."""
    if torch.is_tensor(gamma):
        g = gamma.detach().reshape(-1)[:1]
        if g.device != dev or g.dtype != torch.float32:
            g = g.to(device=dev, dtype=torch.float32)
        return g.contiguous()
    return torch.full((1,), float(gamma), device=dev, dtype=torch.float32)


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _flags(preserve_input, always_clip, use_offset=True, conf_mode="premul", blend="post", legacy=False):
    if conf_mode not in ("premul", "sampled", "none") or blend not in ("post", "pre"):
        raise ValueError("conf_mode must be premul|sampled|none and blend post|pre")
    return (_lib.FLAG_PRESERVE_INPUT if preserve_input else 0) | \
           (_lib.FLAG_ALWAYS_CLIP if always_clip else 0) | \
           (0 if use_offset else _lib.FLAG_NO_OFFSET) | \
           (_lib.FLAG_BLEND_PRE if blend == "pre" else 0) | \
           (_lib.FLAG_CONF_SAMPLED if conf_mode == "sampled" else 0) | \
           (_lib.FLAG_LEGACY if (legacy and conf_mode == "sampled") else 0)


def device_info(device=None):
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    sm, l2 = ctypes.c_int(0), ctypes.c_int(0)
    _lib.check(_lib.load().nlspn_device_info(dev.index or 0, ctypes.byref(sm), ctypes.byref(l2)),
               "nlspn_device_info")
    return dict(sm_count=sm.value, l2_bytes=l2.value)


def prologue_fwd(guidance, confidence, feat_init, feat_fix, gamma, K, affinity="TGASS",
                 preserve_input=True, always_clip=False):
    """-> (offset [B,2KK,H,W], aff [B,KK,H,W], conf_fixed [B,1,H,W] | None, src0 [B,1,H,W])."""
    lib = _lib.load()
    B, _, H, W = feat_init.shape
    N = K * K - 1
    feat_init = _chk("feat_init", feat_init, (B, 1, H, W))
    guidance = _chk("guidance", guidance, (B, 3 * N, H, W))
    confidence = _chk("confidence", confidence, (B, 1, H, W), optional=True)
    feat_fix = _chk("feat_fix", feat_fix, (B, 1, H, W), optional=True)
    preserve = bool(preserve_input and feat_fix is not None)
    dev = feat_init.device
    opt = dict(device=dev, dtype=torch.float32)
    offset = torch.empty((B, 2 * K * K, H, W), **opt)
    aff = torch.empty((B, K * K, H, W), **opt)
    conf_fixed = torch.empty((B, 1, H, W), **opt) if confidence is not None else None
    src0 = torch.empty((B, 1, H, W), **opt)
    gam = _gamma(gamma, dev)
    with torch.cuda.device(dev):
        rc = lib.nlspn_prologue_fwd(_ptr(guidance), _ptr(confidence), _ptr(feat_init), _ptr(feat_fix),
                                    _ptr(gam), _lib.AFFINITY[affinity], _flags(preserve, always_clip),
                                    B, H, W, K, _ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(src0),
                                    _stream(dev))
    _lib.check(rc, "nlspn_prologue_fwd")
    return offset, aff, conf_fixed, src0


def forward(guidance, confidence, feat_init, feat_fix, gamma, K, T, affinity="TGASS",
            preserve_input=True, always_clip=False, keep_src=True, use_offset=True,
            conf_mode="premul", blend="post", legacy=False):
    """Fused prologue + T iterations (group-major).
    -> (offset | None, aff, conf_fixed | None, src [S,B,1,H,W], list_feat [T,B,1,H,W]).
    use_offset=False: fixed-local propagation (nlspnmodel.py:209-224), guidance is [B,N,H,W]."""
    lib = _lib.load()
    B, _, H, W = feat_init.shape
    N = K * K - 1
    feat_init = _chk("feat_init", feat_init, (B, 1, H, W))
    guidance = _chk("guidance", guidance, (B, (3 if use_offset else 1) * N, H, W))
    confidence = _chk("confidence", confidence, (B, 1, H, W), optional=True)
    feat_fix = _chk("feat_fix", feat_fix, (B, 1, H, W), optional=True)
    preserve = bool(preserve_input and feat_fix is not None)
    dev = feat_init.device
    opt = dict(device=dev, dtype=torch.float32)
    if conf_mode == "none":
        confidence = None
    if conf_mode == "sampled" and confidence is None:
        raise RuntimeError("conf_mode='sampled' needs a confidence map")
    offset = torch.empty((B, 2 * K * K, H, W), **opt) if use_offset else None
    aff = torch.empty((B, K * K, H, W), **opt)
    premul = confidence is not None and conf_mode == "premul"
    conf_fixed = torch.empty((B, 1, H, W), **opt) if premul else None
    use_src = premul or blend == "pre"            # the gather source differs from list_feat
    S = 1 if not use_src else (T if keep_src else min(T, 2))
    src = torch.empty((S, B, 1, H, W), **opt)
    list_feat = torch.empty((T, B, 1, H, W), **opt)
    gam = _gamma(gamma, dev)
    with torch.cuda.device(dev):
        rc = lib.nlspn_forward(_ptr(guidance), _ptr(confidence), _ptr(feat_init), _ptr(feat_fix),
                               _ptr(gam), _lib.AFFINITY[affinity],
                               _flags(preserve, always_clip, use_offset, conf_mode, blend, legacy),
                               B, H, W, K, T, _ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(src), S,
                               _ptr(list_feat), _stream(dev))
    _lib.check(rc, "nlspn_forward")
    return offset, aff, conf_fixed, src, list_feat


def propagate_fwd(offset, aff, conf_fixed, feat_fix, src, list_feat, K, T,
                  preserve_input=True, always_clip=False):
    """Runs T iterations in place: src [S,B,1,H,W] (plane 0 filled), list_feat [T,B,1,H,W]."""
    lib = _lib.load()
    S, B, _, H, W = src.shape
    offset = _chk("offset", offset, (B, 2 * K * K, H, W))
    aff = _chk("aff", aff, (B, K * K, H, W))
    conf_fixed = _chk("conf_fixed", conf_fixed, (B, 1, H, W), optional=True)
    feat_fix = _chk("feat_fix", feat_fix, (B, 1, H, W), optional=True)
    _chk("list_feat", list_feat, (T, B, 1, H, W))
    if not (src.is_contiguous() and list_feat.is_contiguous()):
        raise RuntimeError("src and list_feat must be contiguous (they are written in place)")
    preserve = bool(preserve_input and feat_fix is not None)
    dev = src.device
    with torch.cuda.device(dev):
        rc = lib.nlspn_propagate_fwd(_ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(feat_fix),
                                     _flags(preserve, always_clip), B, H, W, K, T, _ptr(src), S,
                                     _ptr(list_feat), _stream(dev))
    _lib.check(rc, "nlspn_propagate_fwd")
    return list_feat


def backward(guidance, feat_init, feat_fix, offset, aff, conf_fixed, src, list_feat, g_list,
             gamma, K, T, affinity="TGASS", preserve_input=True, always_clip=False,
             g_offset_ext=None, g_aff_ext=None, per_iteration=False, use_offset=True,
             conf_mode="premul", blend="post", legacy=False, confidence=None, deterministic=False):
    """g_list: sequence of T tensors [B,1,H,W] or None.  -> (g_init, g_guidance, g_conf, g_gamma).
    deterministic: NLSPN_FLAG_DETERMINISTIC -- bit-identical gradients from run to run (slower)."""
    lib = _lib.load()
    B, _, H, W = feat_init.shape
    N = K * K - 1
    S = src.shape[0]
    dev = feat_init.device
    opt = dict(device=dev, dtype=torch.float32)
    feat_fix = _chk("feat_fix", feat_fix, (B, 1, H, W), optional=True)
    preserve = bool(preserve_input and feat_fix is not None)
    keep = [_chk("g_list[%d]" % i, g, (B, 1, H, W), optional=True) for i, g in enumerate(g_list)]
    if len(keep) != T:
        raise RuntimeError("g_list must have T entries")
    ptrs = (ctypes.c_void_p * T)(*[(g.data_ptr() if g is not None else None) for g in keep])
    g_offset_ext = _chk("g_offset_ext", g_offset_ext, (B, 2 * K * K, H, W), optional=True)
    g_aff_ext = _chk("g_aff_ext", g_aff_ext, (B, K * K, H, W), optional=True)
    g_init = torch.empty((B, 1, H, W), **opt)
    g_guid = torch.empty((B, (3 if use_offset else 1) * N, H, W), **opt)
    sampled = conf_mode == "sampled"
    confidence = _chk("confidence", confidence, (B, 1, H, W), optional=not sampled)
    g_conf = torch.empty((B, 1, H, W), **opt) if (conf_fixed is not None or sampled) else None
    g_gamma = torch.empty((1,), device=dev, dtype=torch.float64)
    flags = _flags(preserve, always_clip, use_offset, conf_mode, blend, legacy) \
        | (_lib.FLAG_BWD_PER_ITERATION if per_iteration else 0) | (_lib.FLAG_DETERMINISTIC if deterministic else 0)
    nbytes = lib.nlspn_backward_workspace_bytes_ex(B, H, W, K, T, flags)
    ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
    gam = _gamma(gamma, dev)
    with torch.cuda.device(dev):
        rc = lib.nlspn_backward(_ptr(guidance.contiguous()), _ptr(feat_init.contiguous()), _ptr(feat_fix),
                                _ptr(confidence if sampled else None),
                                _ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(src), S, _ptr(list_feat),
                                ptrs, _ptr(g_offset_ext), _ptr(g_aff_ext), _ptr(gam),
                                _lib.AFFINITY[affinity], flags,
                                B, H, W, K, T, _ptr(g_init), _ptr(g_guid), _ptr(g_conf), _ptr(g_gamma), _ptr(ws),
                                nbytes, _stream(dev))
    _lib.check(rc, "nlspn_backward")
    return g_init, g_guid, g_conf, g_gamma


_DCN_INTS = ("kernel_h", "kernel_w", "stride_h", "stride_w", "pad_h", "pad_w", "dilation_h",
             "dilation_w", "group", "deformable_group", "im2col_step")


def _dcn_dtype(input):
    """The single-step operator serves float32 (tuned kernels) and float64 (the reference dispatches
    both, modulated_deform_conv_cuda.cu:93,224); anything else raises."""
    if input.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("DCN: float32 or float64 tensors only (got %s)" % input.dtype)
    return input.dtype


def dcn_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, im2col_step):
    lib = _lib.load()
    dt = _dcn_dtype(input)
    input = _chk("input", input, dtype=dt)
    B, C, H, W = input.shape
    weight, bias, offset, mask = (_chk(n, t, dtype=dt) for n, t in
                                  (("weight", weight), ("bias", bias), ("offset", offset), ("mask", mask)))
    KK = kernel_h * kernel_w
    if tuple(offset.shape) != (B, 2 * KK, H, W) or tuple(mask.shape) != (B, KK, H, W):
        raise RuntimeError("offset/mask shape does not match input and kernel size")
    out = torch.empty_like(input)
    dev = input.device
    with torch.cuda.device(dev):
        fn = lib.nlspn_dcn_forward if dt == torch.float32 else lib.nlspn_dcn_forward_f64
        rc = fn(_ptr(input), _ptr(weight), _ptr(bias), _ptr(offset), _ptr(mask),
                kernel_h, kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                dilation_w, group, deformable_group, im2col_step, B, C, H, W,
                _ptr(out), _stream(dev))
    _lib.check(rc, "nlspn_dcn_forward")
    return out


def dcn_backward(input, weight, bias, offset, mask, grad_output, kernel_h, kernel_w, stride_h,
                 stride_w, pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, im2col_step):
    lib = _lib.load()
    dt = _dcn_dtype(input)
    input = _chk("input", input, dtype=dt)
    B, C, H, W = input.shape
    weight, bias, offset, mask = (_chk(n, t, dtype=dt) for n, t in
                                  (("weight", weight), ("bias", bias), ("offset", offset), ("mask", mask)))
    grad_output = _chk("grad_output", grad_output, (B, C, H, W), dtype=dt)
    gi, go, gm = torch.empty_like(input), torch.empty_like(offset), torch.empty_like(mask)
    gw, gb = torch.empty_like(weight), torch.empty_like(bias)
    dev = input.device
    with torch.cuda.device(dev):
        if dt == torch.float32:
            # vector-RED scatter into blocked planes + TMA-delivered gather source (kernels_step.cuh)
            nbytes = lib.nlspn_dcn_backward_workspace_bytes(B, H, W, kernel_h)
            ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
            rc = lib.nlspn_dcn_backward_ws(_ptr(input), _ptr(weight), _ptr(bias), _ptr(offset), _ptr(mask),
                                           _ptr(grad_output), kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                           pad_w, dilation_h, dilation_w, group, deformable_group, im2col_step,
                                           B, C, H, W, _ptr(gi), _ptr(go), _ptr(gm), _ptr(gw), _ptr(gb),
                                           _ptr(ws), nbytes, _stream(dev))
        else:
            rc = lib.nlspn_dcn_backward_f64(_ptr(input), _ptr(weight), _ptr(bias), _ptr(offset), _ptr(mask),
                                            _ptr(grad_output), kernel_h, kernel_w, stride_h, stride_w, pad_h,
                                            pad_w, dilation_h, dilation_w, group, deformable_group, im2col_step,
                                            B, C, H, W, _ptr(gi), _ptr(go), _ptr(gm), _ptr(gw), _ptr(gb),
                                            _stream(dev))
    _lib.check(rc, "nlspn_dcn_backward")
    return gi, go, gm, gw, gb


def step_fwd(src_prev, offset, aff, conf_fixed, feat_fix, K, preserve_input=True, always_clip=False):
    """One fused iteration (nlspnmodel.py:350-361).  -> (out, src_next | None).  offset=None: fixed-local."""
    lib = _lib.load()
    src_prev = _chk("src_prev", src_prev)
    B, _, H, W = src_prev.shape
    offset = _chk("offset", offset, (B, 2 * K * K, H, W), optional=True)
    aff = _chk("aff", aff, (B, K * K, H, W))
    conf_fixed = _chk("conf_fixed", conf_fixed, (B, 1, H, W), optional=True)
    feat_fix = _chk("feat_fix", feat_fix, (B, 1, H, W), optional=True)
    preserve = bool(preserve_input and feat_fix is not None)
    out = torch.empty_like(src_prev)
    src_next = torch.empty_like(src_prev) if conf_fixed is not None else None
    dev = src_prev.device
    with torch.cuda.device(dev):
        rc = lib.nlspn_step_fwd(_ptr(src_prev), _ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(feat_fix),
                                _flags(preserve, always_clip, offset is not None), B, H, W, K, _ptr(out),
                                _ptr(src_next), _stream(dev))
    _lib.check(rc, "nlspn_step_fwd")
    return out, src_next


def step_bwd(src_prev, offset, aff, conf_fixed, feat_fix, out, g_out, g_src_next, K,
             preserve_input=True, always_clip=False):
    """-> (g_src_prev, g_offset | None, g_aff, g_conf | None)."""
    lib = _lib.load()
    B, _, H, W = src_prev.shape
    dev = src_prev.device
    g_out = _chk("g_out", g_out, (B, 1, H, W), optional=True)
    g_src_next = _chk("g_src_next", g_src_next, (B, 1, H, W), optional=True)
    preserve = bool(preserve_input and feat_fix is not None)
    flags = _flags(preserve, always_clip, offset is not None)
    g_src_prev = torch.empty_like(src_prev)
    g_off = torch.empty_like(offset) if offset is not None else None
    g_aff = torch.empty_like(aff)
    g_conf = torch.empty_like(conf_fixed) if conf_fixed is not None else None
    nbytes = lib.nlspn_step_bwd_workspace_bytes(B, H, W, K, flags)
    ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
    with torch.cuda.device(dev):
        rc = lib.nlspn_step_bwd(_ptr(src_prev), _ptr(offset), _ptr(aff), _ptr(conf_fixed), _ptr(feat_fix), _ptr(out),
                                _ptr(g_out), _ptr(g_src_next), flags, B, H, W, K, _ptr(g_src_prev), _ptr(g_off),
                                _ptr(g_aff), _ptr(g_conf), _ptr(ws), nbytes, _stream(dev))
    _lib.check(rc, "nlspn_step_bwd")
    return g_src_prev, g_off, g_aff, g_conf


def debug_indices(offset, K):
    lib = _lib.load()
    offset = _chk("offset", offset)
    B, _, H, W = offset.shape
    idx = torch.empty((B, K * K, 3, H, W), device=offset.device, dtype=torch.int32)
    with torch.cuda.device(offset.device):
        rc = lib.nlspn_debug_indices(_ptr(offset), B, H, W, K, _ptr(idx), _stream(offset.device))
    _lib.check(rc, "nlspn_debug_indices")
    return idx
