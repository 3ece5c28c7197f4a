"""The three final head convolutions in front of the propagation as ONE tcgen05 implicit GEMM (SURVEY 8f row f3,
first half; C entries nlspn_heads_pack / nlspn_heads_fwd, csrc/kernels_head.cuh).

Replaces, in the reference's ``NLSPNModel.forward`` (nlspnmodel.py:297,301,313; layers :69-86)::

    pred_init  = id_dec0(cat(id_fd1, fe1))            # conv3x3 128 -> 1,  ReLU
    off_aff    = off_aff_dec0(cat(off_aff_fd1, fe1))  # conv3x3 128 -> 3N
    confidence = cf_dec0(cat(cf_fd1, fe1))            # conv3x3 128 -> 1,  Sigmoid

without the three concatenations.  Arithmetic: TF32 products, fp32 accumulation -- what cuDNN uses for these layers
under PyTorch's default ``torch.backends.cudnn.allow_tf32 = True``.  The backward (data and weight gradients of the
same layers) is stock torch (cuDNN), from the saved inputs.  There is no fallback in here: callers that cannot use it
(CPU tensors, other channel counts) keep their stock layers -- see ``model.NLSPNModel``."""
from __future__ import annotations

import ctypes

import torch
from torch.nn import functional as TF

from . import _lib

__all__ = ["fused_heads", "FusedHeadsFunction", "supported"]

CIN = 64       # channels of each of the four tensors (64 + 64 = the reference's 128-channel concatenations)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def supported(id_fd1, oa_fd1, cf_fd1, fe1, K) -> bool:
    ts = (id_fd1, oa_fd1, cf_fd1, fe1)
    return all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[1] == CIN for t in ts) \
        and all(t.shape == fe1.shape for t in ts) and K in (3, 5, 7)


def _forward(id_fd1, oa_fd1, cf_fd1, fe1, w_id, b_id, w_oa, b_oa, w_cf, b_cf, K):
    lib = _lib.load()
    B, C, H, W = fe1.shape
    N3 = 3 * (K * K - 1)
    if not supported(id_fd1, oa_fd1, cf_fd1, fe1, K):
        raise RuntimeError("fused_heads: four CUDA float32 [B,64,H,W] tensors and prop_kernel in (3, 5, 7) are required")
    if tuple(w_id.shape) != (1, 2 * CIN, 3, 3) or tuple(w_oa.shape) != (N3, 2 * CIN, 3, 3) or tuple(w_cf.shape) != (1, 2 * CIN, 3, 3):
        raise RuntimeError("fused_heads: weights must be [1,128,3,3], [%d,128,3,3], [1,128,3,3]" % N3)
    dev = fe1.device
    ins = [t.contiguous() for t in (id_fd1, oa_fd1, cf_fd1, fe1)]
    packed = torch.empty((lib.nlspn_heads_packed_floats(K),), device=dev, dtype=torch.float32)
    bias = torch.cat([b_id.reshape(1), b_oa.reshape(-1), b_cf.reshape(1)]).to(dev, torch.float32).contiguous()
    pred_init = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    guidance = torch.empty((B, N3, H, W), device=dev, dtype=torch.float32)
    confidence = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(lib.nlspn_heads_pack(_ptr(w_id.contiguous()), _ptr(w_oa.contiguous()), _ptr(w_cf.contiguous()), K,
                                        _ptr(packed), st), "nlspn_heads_pack")
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
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_init, g_guid, g_conf):
        id_fd1, oa_fd1, cf_fd1, fe1, w_id, w_oa, w_cf, pred_init, confidence = ctx.saved_tensors
        z = torch.zeros_like
        g_init = z(pred_init) if g_init is None else g_init * (pred_init > 0).to(g_init.dtype)        # ReLU
        g_conf = z(confidence) if g_conf is None else g_conf * confidence * (1.0 - confidence)         # Sigmoid
        g_guid = torch.zeros((fe1.shape[0], w_oa.shape[0]) + tuple(fe1.shape[2:]), device=fe1.device) if g_guid is None else g_guid
        grads_in = [None, None, None, None]
        g_fe1 = None
        g_w, g_b = [], []
        for k, (x, w, g) in enumerate(((id_fd1, w_id, g_init), (oa_fd1, w_oa, g_guid), (cf_fd1, w_cf, g_conf))):
            g = g.contiguous()
            xin = torch.cat((x, fe1), 1)
            gi = torch.nn.grad.conv2d_input(xin.shape, w, g, stride=1, padding=1)
            grads_in[k] = gi[:, :CIN]
            g_fe1 = gi[:, CIN:] if g_fe1 is None else g_fe1 + gi[:, CIN:]
            g_w.append(torch.nn.grad.conv2d_weight(xin, w.shape, g, stride=1, padding=1))
            g_b.append(g.sum(dim=(0, 2, 3)))
        grads_in[3] = g_fe1
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
