"""The three final head convolutions in front of the propagation as ONE tcgen05 implicit GEMM (SURVEY 8f row f3,
first half; C entries nlspn_heads_pack / nlspn_heads_fwd, csrc/kernels_head.cuh).

Replaces, in the reference's ``NLSPNModel.forward`` (nlspnmodel.py:297,301,313; layers :69-86)::

    pred_init  = id_dec0(cat(id_fd1, fe1))            # conv3x3 128 -> 1,  ReLU
    off_aff    = off_aff_dec0(cat(off_aff_fd1, fe1))  # conv3x3 128 -> 3N
    confidence = cf_dec0(cat(cf_fd1, fe1))            # conv3x3 128 -> 1,  Sigmoid

without the three concatenations.  Arithmetic: TF32 products, fp32 accumulation -- what cuDNN uses for these layers
under PyTorch's default ``torch.backends.cudnn.allow_tf32 = True``.  Backward: the weight and bias gradients are tcgen05
too (nlspn_heads_grad_prep / nlspn_heads_wgrad, csrc/kernels_head_wgrad.cuh; W % 4 == 0), the data gradients stock torch
(cuDNN), all from the saved inputs.  There is no fallback in here: callers that cannot use it
(CPU tensors, other channel counts) keep their stock layers -- see ``model.NLSPNModel``."""
from __future__ import annotations

import ctypes

import torch
from torch.nn import functional as TF

from . import _lib

__all__ = ["fused_heads", "fused_heads_prologue", "prologue_supported", "wgrad_supported", "grad_prep", "weight_grads", "dgrad_one", "dgrad_wide", "dgrad_supported", "FusedHeadsFunction", "supported"]

CIN = 64       # channels of each of the four tensors (64 + 64 = the reference's 128-channel concatenations)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def supported(id_fd1, oa_fd1, cf_fd1, fe1, K) -> bool:
    ts = (id_fd1, oa_fd1, cf_fd1, fe1)
    return all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == CIN for t in ts) \
        and all(t.shape == fe1.shape for t in ts) and K in (3, 5, 7)


def _prepare(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, K):
    """Validation + the packed weight matrix and the bias vector of one call (on the current stream)."""
    lib = _lib.load()
    N3 = 3 * (K * K - 1)
    if not supported(id_fd1, oa_fd1, cf_fd1, fe1, K):
        raise RuntimeError("fused_heads: four CUDA float32 [B,64,H,W] tensors and prop_kernel in (3, 5, 7) are required")
    if tuple(w_id.shape) != (1, 2 * CIN, 3, 3) or tuple(w_oa.shape) != (N3, 2 * CIN, 3, 3) or tuple(w_cf.shape) != (1, 2 * CIN, 3, 3):
        raise RuntimeError("fused_heads: weights must be [1,128,3,3], [%d,128,3,3], [1,128,3,3]" % N3)
    dev = fe1.device
    ins = [t.contiguous() for t in (id_fd1, oa_fd1, cf_fd1, fe1)]
    packed = torch.empty((lib.nlspn_heads_packed_floats(K),), device=dev, dtype=torch.float32)
    bias = torch.cat([b_id.reshape(1), b_oa.reshape(-1), b_cf.reshape(1)]).to(dev, torch.float32).contiguous()
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_pack(_ptr(w_id.contiguous()), _ptr(w_oa.contiguous()), _ptr(w_cf.contiguous()), K,
                                        _ptr(packed), st), "nlspn_heads_pack")
    return lib, ins, packed, bias, st


def prologue_supported(W, K) -> bool:
    """Can the head GEMM run the propagation's prologue as its epilogue for this width / prop_kernel?"""
    return bool(_lib.load().nlspn_heads_prologue_supported(int(W), int(K)))


def fused_heads_prologue(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, feat_fix, gamma, prop_kernel=3,
                         affinity="TGASS", preserve_input=True, always_clip=False, conf_prop=True, src0=None,
                         want_guidance=False):
    """Heads + prologue in ONE kernel (nlspn_heads_prologue_fwd; inference -- no autograd): the 3N-channel `guidance`
    tensor never reaches HBM unless ``want_guidance``.  Replaces nlspnmodel.py:297-313 followed by :252-269, :179-201,
    :328-351.  -> dict(pred_init, confidence, guidance|None, offset, aff, conf_fixed|None, src0); continue with
    ``functional.propagate_fwd``.  ``src0``: optional [B,1,H,W] view to write the first gather source into (plane 0
    of the propagation's ``src`` buffer)."""
    K = int(prop_kernel)
    lib, ins, packed, bias, st = _prepare(id_fd1.detach(), oa_fd1.detach(), cf_fd1.detach(), fe1.detach(), w_id.detach(),
                                          b_id.detach(), w_oa.detach(), b_oa.detach(), w_cf.detach(), b_cf.detach(), K)
    B, C, H, W = fe1.shape
    N3 = 3 * (K * K - 1)
    dev = fe1.device
    opt = dict(device=dev, dtype=torch.float32)
    preserve = bool(preserve_input and feat_fix is not None)
    if feat_fix is not None:
        if not (feat_fix.is_cuda and feat_fix.dtype == torch.float32 and tuple(feat_fix.shape) == (B, 1, H, W)):
            raise RuntimeError("fused_heads_prologue: feat_fix must be a CUDA float32 [B,1,H,W] tensor")
        feat_fix = feat_fix.detach().contiguous()
    gam = gamma.detach().reshape(-1)[:1].to(device=dev, dtype=torch.float32).contiguous() if torch.is_tensor(gamma) \
        else torch.full((1,), float(gamma), **opt)
    out = dict(pred_init=torch.empty((B, 1, H, W), **opt), confidence=torch.empty((B, 1, H, W), **opt),
               guidance=torch.empty((B, N3, H, W), **opt) if want_guidance else None,
               offset=torch.empty((B, 2 * K * K, H, W), **opt), aff=torch.empty((B, K * K, H, W), **opt),
               conf_fixed=torch.empty((B, 1, H, W), **opt) if conf_prop else None,
               src0=torch.empty((B, 1, H, W), **opt) if src0 is None else src0)
    if not (out["src0"].is_cuda and out["src0"].is_contiguous() and tuple(out["src0"].shape) == (B, 1, H, W)):
        raise RuntimeError("fused_heads_prologue: src0 must be a contiguous CUDA [B,1,H,W] tensor")
    flags = (_lib.FLAG_PRESERVE_INPUT if preserve else 0) | (_lib.FLAG_ALWAYS_CLIP if always_clip else 0)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_prologue_fwd(_ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), _ptr(ins[3]), _ptr(packed), _ptr(bias),
                                                _ptr(feat_fix if preserve else None), _ptr(gam), _lib.AFFINITY[affinity], flags,
                                                B, H, W, K, _ptr(out["pred_init"]), _ptr(out["confidence"]), _ptr(out["guidance"]),
                                                _ptr(out["offset"]), _ptr(out["aff"]), _ptr(out["conf_fixed"]), _ptr(out["src0"]), st),
                   "nlspn_heads_prologue_fwd")
    return out


def _forward(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, K):
    lib, ins, packed, bias, st = _prepare(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, K)
    B, C, H, W = fe1.shape
    N3 = 3 * (K * K - 1)
    dev = fe1.device
    pred_init = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    guidance = torch.empty((B, N3, H, W), device=dev, dtype=torch.float32)
    confidence = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_fwd(_ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), _ptr(ins[3]), _ptr(packed), _ptr(bias),
                                       B, H, W, K, _ptr(pred_init), _ptr(guidance), _ptr(confidence), st),
                   "nlspn_heads_fwd")
    return pred_init, guidance, confidence


class FusedHeadsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, K):
        outs = _forward(id_fd1.detach(), oa_fd1.detach(), cf_fd1.detach(), fe1.detach(), w_id.detach(), b_id.detach(),
                        w_oa.detach(), b_oa.detach(), w_cf.detach(), b_cf.detach(), K)
        ctx.save_for_backward(id_fd1, oa_fd1, cf_fd1, fe1, w_id, w_oa, w_cf, outs[0], outs[2])
        ctx.K = int(K)
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_init, g_guid, g_conf):
        """W % 4 == 0: `_backward_native` (weight gradients on tcgen05).  Otherwise:
        stock cuDNN data / weight gradients of the same three layers, WITHOUT re-building the three 128-channel
        concatenations: a convolution is linear in its input channels, so each head's own 64-channel branch gets its own
        (64-channel) gradient calls, and the shared fe1 gets ONE call pair with all 3N + 2 output channels at once
        (instead of three 128-channel calls whose fe1 halves are then added).  KITTI B = 8: 16.5 -> see DESIGN 12."""
        id_fd1, oa_fd1, cf_fd1, fe1, w_id, w_oa, w_cf, pred_init, confidence = ctx.saved_tensors
        need_in, need_w = ctx.needs_input_grad[:4], (ctx.needs_input_grad[4], ctx.needs_input_grad[6], ctx.needs_input_grad[8])
        # (the split-K partials of the native weight / bias gradients are added with fp32 atomics: under
        #  torch.use_deterministic_algorithms(True) the stock path below, which torch makes deterministic, is taken)
        if wgrad_supported(fe1.shape[3], ctx.K) and not torch.are_deterministic_algorithms_enabled() \
                and all(t.is_contiguous() and t.data_ptr() % 16 == 0 for t in (id_fd1, oa_fd1, cf_fd1, fe1)):
            return _backward_native(ctx.K, need_in, need_w, id_fd1, oa_fd1, cf_fd1, fe1, w_id, w_oa, w_cf, pred_init, confidence,
                                    g_init, g_guid, g_conf)
        z = torch.zeros_like
        g_init = z(pred_init) if g_init is None else g_init * (pred_init > 0).to(g_init.dtype)        # ReLU
        g_conf = z(confidence) if g_conf is None else g_conf * confidence * (1.0 - confidence)         # Sigmoid
        g_guid = torch.zeros((fe1.shape[0], w_oa.shape[0]) + tuple(fe1.shape[2:]), device=fe1.device) if g_guid is None else g_guid
        gs = [g_init.contiguous(), g_guid.contiguous(), g_conf.contiguous()]
        grads_in, g_w_own = [None, None, None, None], [None, None, None]
        for k, (x, w, g) in enumerate(((id_fd1, w_id, gs[0]), (oa_fd1, w_oa, gs[1]), (cf_fd1, w_cf, gs[2]))):
            w_own = w[:, :CIN].contiguous()
            if need_in[k]:
                grads_in[k] = torch.nn.grad.conv2d_input(x.shape, w_own, g, stride=1, padding=1)
            if need_w[k]:
                g_w_own[k] = torch.nn.grad.conv2d_weight(x, w_own.shape, g, stride=1, padding=1)
        g_all = torch.cat(gs, 1)                                                  # [B, 3N + 2, H, W]
        w_fe = torch.cat((w_id[:, CIN:], w_oa[:, CIN:], w_cf[:, CIN:]), 0).contiguous()      # [3N + 2, 64, 3, 3]
        if need_in[3]:
            grads_in[3] = torch.nn.grad.conv2d_input(fe1.shape, w_fe, g_all, stride=1, padding=1)
        g_w = [None, None, None]
        if any(need_w):
            g_w_fe = torch.nn.grad.conv2d_weight(fe1, w_fe.shape, g_all, stride=1, padding=1)
            n_oa = w_oa.shape[0]
            parts = (g_w_fe[:1], g_w_fe[1:1 + n_oa], g_w_fe[1 + n_oa:])
            g_w = [torch.cat((g_w_own[k], parts[k]), 1) if need_w[k] else None for k in range(3)]
        g_b = [g.sum(dim=(0, 2, 3)) for g in gs]
        return grads_in[0], grads_in[1], grads_in[2], grads_in[3], g_w[0], g_b[0], g_w[1], g_b[1], g_w[2], g_b[2], None


def wgrad_supported(W, K) -> bool:
    """Do the tcgen05 weight-gradient kernels (nlspn_heads_grad_prep / nlspn_heads_wgrad) cover this width / prop_kernel?"""
    return bool(_lib.load().nlspn_heads_wgrad_supported(int(W), int(K)))


def grad_prep(pred_init, confidence, g_init, g_guid, g_conf, K):
    """nlspn_heads_grad_prep: -> (g_shift [3,B,3N+2,H,W], g_bias [3N+2]); a None gradient is a zero gradient.  g_shift[1] is
    the concatenated gradient (channel 0 = init, 1 = confidence, 2.. = guidance) with the ReLU / Sigmoid derivatives applied,
    g_shift[0] / g_shift[2] the same shifted by one pixel (g[.., x + 1] / g[.., x - 1], zero outside the row)."""
    lib = _lib.load()
    B, _, H, W = pred_init.shape
    NT = 3 * (K * K - 1) + 2
    dev = pred_init.device
    gin = [None if g is None else g.to(torch.float32).contiguous() for g in (g_init, g_guid, g_conf)]
    g_shift = torch.empty((3, B, NT, H, W), device=dev, dtype=torch.float32)
    g_bias = torch.empty((NT,), device=dev, dtype=torch.float32)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_grad_prep(_ptr(gin[0]), _ptr(pred_init.contiguous()), _ptr(gin[1]), _ptr(gin[2]),
                                             _ptr(confidence.contiguous()), B, H, W, K, _ptr(g_shift), _ptr(g_bias), st),
                   "nlspn_heads_grad_prep")
    return g_shift, g_bias


def weight_grads(id_fd1, oa_fd1, cf_fd1, fe1, g_shift, K):
    """nlspn_heads_wgrad: -> dw_all [3N+2,128,3,3] (row 0 = dw_id, 1 = dw_cf, 2.. = dw_oa); a None branch tensor leaves
    its 64 input channels zero."""
    lib = _lib.load()
    B, _, H, W = fe1.shape
    dev = fe1.device
    dw_all = torch.empty((3 * (K * K - 1) + 2, 2 * CIN, 3, 3), device=dev, dtype=torch.float32)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_wgrad(_ptr(id_fd1), _ptr(oa_fd1), _ptr(cf_fd1), _ptr(fe1), _ptr(g_shift), B, H, W, K,
                                         _ptr(dw_all), st), "nlspn_heads_wgrad")
    return dw_all


def dgrad_one(g_all, w_id, w_cf, K):
    """nlspn_heads_dgrad_one: -> (d_id_fd1, d_cf_fd1) [B,64,H,W] from g_all = grad_prep(..)[0][1]; a None weight skips that head."""
    lib = _lib.load()
    B, _, H, W = g_all.shape
    dev = g_all.device
    outs = [None if w is None else torch.empty((B, CIN, H, W), device=dev, dtype=torch.float32) for w in (w_id, w_cf)]
    ws = [None if w is None else w.detach().to(torch.float32).contiguous() for w in (w_id, w_cf)]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_dgrad_one(_ptr(g_all), _ptr(ws[0]), _ptr(ws[1]), B, H, W, K, _ptr(outs[0]), _ptr(outs[1]), st),
                   "nlspn_heads_dgrad_one")
    return outs[0], outs[1]


def dgrad_supported(W, K) -> bool:
    """Does nlspn_heads_dgrad_wide (the wide data gradients as a tcgen05 GEMM) cover this width / prop_kernel?"""
    return bool(_lib.load().nlspn_heads_dgrad_supported(int(W), int(K)))


def dgrad_wide(g_shift, w_id, w_oa, w_cf, K, want_oa=True):
    """nlspn_heads_dgrad_pack + nlspn_heads_dgrad_wide: -> (d_oa_fd1 | None, d_fe1) [B,64,H,W] from g_shift = grad_prep(..)[0]."""
    lib = _lib.load()
    _, B, _, H, W = g_shift.shape
    dev = g_shift.device
    opt = dict(device=dev, dtype=torch.float32)
    packed = torch.empty((lib.nlspn_heads_dgrad_packed_floats(K),), **opt)
    d_oa = torch.empty((B, CIN, H, W), **opt) if want_oa else None
    d_fe = torch.empty((B, CIN, H, W), **opt)
    ws = [w.detach().to(torch.float32).contiguous() for w in (w_id, w_oa, w_cf)]
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_dgrad_pack(_ptr(ws[0]), _ptr(ws[1]), _ptr(ws[2]), K, _ptr(packed), st), "nlspn_heads_dgrad_pack")
        _lib.check(lib.nlspn_heads_dgrad_wide(_ptr(g_shift), _ptr(packed), B, H, W, K, _ptr(d_oa), _ptr(d_fe), st),
                   "nlspn_heads_dgrad_wide")
    return d_oa, d_fe


def _backward_native(K, need_in, need_w, id_fd1, oa_fd1, cf_fd1, fe1, w_id, w_oa, w_cf, pred_init, confidence, g_init, g_guid, g_conf):
    """Activation derivatives + concatenation + bias sums in one kernel (nlspn_heads_grad_prep), every weight gradient on
    tcgen05 (nlspn_heads_wgrad, csrc/kernels_head_wgrad.cuh); the data gradients of the two one-channel heads as an fp32
    stencil (nlspn_heads_dgrad_one); the two wide data gradients (guidance branch, fe1) as one tcgen05 GEMM over the shifted
    copies (nlspn_heads_dgrad_wide; prop_kernel 3 -- otherwise cuDNN, fed from the concatenated gradient)."""
    N3 = 3 * (K * K - 1)
    NT = N3 + 2
    g_shift, g_bias = grad_prep(pred_init, confidence, g_init, g_guid, g_conf, K)
    g_all = g_shift[1]                                                            # [B, 3N + 2, H, W]
    grads_in = [None, None, None, None]
    own = ((id_fd1, w_id, slice(0, 1)), (oa_fd1, w_oa, slice(2, NT)), (cf_fd1, w_cf, slice(1, 2)))      # channel order: init, confidence, guidance
    if need_in[0] or need_in[2]:
        # the one-channel heads: an fp32 nine-tap stencil at the rate of its 64-channel store (nlspn_heads_dgrad_one)
        grads_in[0], grads_in[2] = dgrad_one(g_all, w_id if need_in[0] else None, w_cf if need_in[2] else None, K)
    if need_in[3] and dgrad_supported(fe1.shape[3], K):
        # the two wide ones as one tcgen05 GEMM over the shifted copies (nlspn_heads_dgrad_wide)
        grads_in[1], grads_in[3] = dgrad_wide(g_shift, w_id, w_oa, w_cf, K, want_oa=need_in[1])
    else:
        if need_in[1]:
            # the guidance head has no activation: its upstream gradient is usable as it came
            g = g_guid if (g_guid is not None and g_guid.is_contiguous()) else g_all[:, 2:].contiguous()
            grads_in[1] = torch.nn.grad.conv2d_input(oa_fd1.shape, w_oa[:, :CIN].contiguous(), g, stride=1, padding=1)
        if need_in[3]:
            w_fe = torch.cat((w_id[:, CIN:], w_cf[:, CIN:], w_oa[:, CIN:]), 0).contiguous()      # [3N + 2, 64, 3, 3], g_all's order
            grads_in[3] = torch.nn.grad.conv2d_input(fe1.shape, w_fe, g_all, stride=1, padding=1)
    g_w = [None, None, None]
    if any(need_w):
        dw_all = weight_grads(id_fd1 if need_w[0] else None, oa_fd1 if need_w[1] else None, cf_fd1 if need_w[2] else None,
                              fe1, g_shift, K)
        g_w = [dw_all[sl] if need_w[k] else None for k, (_, _, sl) in enumerate(own)]
    g_b = [g_bias[sl] for _, _, sl in own]
    return grads_in[0], grads_in[1], grads_in[2], grads_in[3], g_w[0], g_b[0], g_w[1], g_b[1], g_w[2], g_b[2], None


def fused_heads(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, prop_kernel=3):
    """-> (pred_init [B,1,H,W], guidance [B,3N,H,W], confidence [B,1,H,W]); differentiable."""
    return FusedHeadsFunction.apply(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, int(prop_kernel))


def reference_heads(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf):
    """The same three layers in stock torch ops (what model.NLSPNModel runs without fusion); used by tests/tools."""
    a = torch.relu(TF.conv2d(torch.cat((id_fd1, fe1), 1), w_id, b_id, 1, 1))
    g = TF.conv2d(torch.cat((oa_fd1, fe1), 1), w_oa, b_oa, 1, 1)
    c = torch.sigmoid(TF.conv2d(torch.cat((cf_fd1, fe1), 1), w_cf, b_cf, 1, 1))
    return a, g, c
