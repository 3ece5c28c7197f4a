"""NLSPN -- drop-in module for the reference's propagation (nlspnmodel.py:323-377).

``NLSPN.forward(feat_init, guidance, confidence, feat_fix, rgb)`` returns
``(feat_result, list_feat, offset, aff, aff_const)`` (the north-star signature; in the
reference fork the same five values are the output-dict entries 'pred' (pre-clamp),
'pred_inter', 'offset', 'aff', 'gamma', nlspnmodel.py:379-381).  The body runs entirely in
libnlspn_b200.so through one ``torch.autograd.Function``.

Fork semantics (the parity target): the state is pre-multiplied by the confidence
before EVERY gather (nlspnmodel.py:350-351), confidence is forced to 1 on fixed pixels (:334),
and the input-preserving blend runs once before iteration 1 and after every iteration
(:341-344,355-357).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as F_

__all__ = ["NLSPN", "NLSPNFunction", "nlspn_propagate", "NLSPNStepFunction", "nlspn_step", "GraphedNLSPN"]


class NLSPNFunction(torch.autograd.Function):
    """(feat_init, guidance, confidence|None, feat_fix|None, gamma) ->
    (list_feat[0..T-1], offset, aff, conf_fixed|None).  Backward: one fused reverse replay."""

    @staticmethod
    def forward(ctx, feat_init, guidance, confidence, feat_fix, gamma, K, T, affinity,
                preserve_input, always_clip, use_offset=True, conf_mode="premul", blend="post", legacy=False,
                need_grad=None, deterministic=False):
        if need_grad is None:   # direct .apply() callers: grad mode is already off in here, so only the inputs count
            need_grad = any(torch.is_tensor(t) and t.requires_grad for t in (feat_init, guidance, confidence, gamma))
        gamma_val = gamma.detach() if torch.is_tensor(gamma) else float(gamma)   # stays on the device
        feat_init_c = feat_init.detach().contiguous()
        guidance_c = guidance.detach().contiguous()
        conf_c = confidence.detach().contiguous() if confidence is not None else None
        fix_c = feat_fix.detach().contiguous() if feat_fix is not None else None
        preserve = bool(preserve_input and fix_c is not None)
        offset, aff, conf_fixed, src, list_feat = F_.forward(
            guidance_c, conf_c, feat_init_c, fix_c, gamma_val, K, T, affinity, preserve, always_clip,
            keep_src=need_grad, use_offset=use_offset, conf_mode=conf_mode, blend=blend, legacy=legacy)
        ctx.cfg = (K, T, affinity, preserve, always_clip, None if torch.is_tensor(gamma_val) else gamma_val)
        ctx.use_offset = bool(use_offset)
        ctx.mode = (conf_mode, blend, bool(legacy))
        ctx.deterministic = bool(deterministic)
        ctx.conf_raw = conf_c if (conf_mode == "sampled" and conf_c is not None) else None
        ctx.conf_grad = conf_c is not None and conf_mode != "none"
        ctx.has_conf = conf_fixed is not None
        ctx.gamma_is_tensor = torch.is_tensor(gamma)
        # gamma goes through save_for_backward so that an in-place update between forward and backward is caught
        ctx.save_for_backward(feat_init_c, guidance_c, fix_c, offset, aff, conf_fixed, src, list_feat,
                              gamma_val if torch.is_tensor(gamma_val) else None)
        ctx.set_materialize_grads(False)
        # outputs: T states, aff, [offset], [conf_fixed]
        outs = tuple(list_feat[t] for t in range(T)) + (aff,)
        if offset is not None:
            outs = outs + (offset,)
        if conf_fixed is not None:
            outs = outs + (conf_fixed,)
        return outs

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        K, T, affinity, preserve, always_clip, gamma_val = ctx.cfg
        feat_init, guidance, feat_fix, offset, aff, conf_fixed, src, list_feat, gamma_t = ctx.saved_tensors
        if gamma_t is not None:
            gamma_val = gamma_t
        g_list = list(grads[:T])
        g_aff_ext = grads[T]
        i = T + 1
        g_off_ext = None
        if ctx.use_offset:
            g_off_ext = grads[i]
            i += 1
        g_cf_ext = grads[i] if ctx.has_conf else None
        conf_mode, blend, legacy = ctx.mode
        g_init, g_guid, g_conf, g_gamma = F_.backward(
            guidance, feat_init, feat_fix, offset, aff, conf_fixed, src, list_feat, g_list, gamma_val,
            K, T, affinity, preserve, always_clip, g_off_ext, g_aff_ext, use_offset=ctx.use_offset,
            conf_mode=conf_mode, blend=blend, legacy=legacy, confidence=ctx.conf_raw,
            deterministic=ctx.deterministic)
        if g_cf_ext is not None:
            # conf_fixed = (1-m)*confidence + m, nlspnmodel.py:334
            m = (feat_fix > 0).to(g_cf_ext.dtype) if preserve else 0.0
            g_conf = g_conf + (1.0 - m) * g_cf_ext
        g_gam = g_gamma.to(torch.float32) if ctx.gamma_is_tensor else None
        return g_init, g_guid, (g_conf if ctx.conf_grad else None), None, g_gam, \
            None, None, None, None, None, None, None, None, None, None, None


class NLSPNStepFunction(torch.autograd.Function):
    """ONE fused iteration of the loop body nlspnmodel.py:350-361 (nlspn_step_fwd / nlspn_step_bwd):
    (src_prev, offset|None, aff, conf_fixed|None, feat_fix|None) -> out [, src_next = out * conf_fixed]."""

    @staticmethod
    def forward(ctx, src_prev, offset, aff, conf_fixed, feat_fix, K, preserve_input, always_clip):
        src_c = src_prev.detach().contiguous()
        off_c = offset.detach().contiguous() if offset is not None else None
        aff_c = aff.detach().contiguous()
        conf_c = conf_fixed.detach().contiguous() if conf_fixed is not None else None
        fix_c = feat_fix.detach().contiguous() if feat_fix is not None else None
        preserve = bool(preserve_input and fix_c is not None)
        out, src_next = F_.step_fwd(src_c, off_c, aff_c, conf_c, fix_c, K, preserve, always_clip)
        ctx.cfg = (K, preserve, bool(always_clip))
        ctx.save_for_backward(src_c, off_c, aff_c, conf_c, fix_c, out)
        ctx.set_materialize_grads(False)
        return out if src_next is None else (out, src_next)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_out, g_src_next=None):
        K, preserve, always_clip = ctx.cfg
        src, off, aff, conf, fix, out = ctx.saved_tensors
        g_src, g_off, g_aff, g_conf = F_.step_bwd(src, off, aff, conf, fix, out, g_out, g_src_next, K,
                                                  preserve, always_clip)
        return g_src, g_off, g_aff, g_conf, None, None, None, None


def nlspn_step(src_prev, offset, aff, conf_fixed, feat_fix, prop_kernel=3, preserve_input=True, always_clip=False):
    """One propagation iteration on the PRE-MULTIPLIED state: -> (out, src_next | None).
    ``offset`` [B,2K^2,H,W] incl. the zero centre pair, or None for the fork's fixed-local propagation;
    ``aff`` [B,K^2,H,W] normalised incl. the centre weight.  Differentiable in src_prev, offset, aff, conf_fixed."""
    r = NLSPNStepFunction.apply(src_prev, offset, aff, conf_fixed, feat_fix, prop_kernel, preserve_input, always_clip)
    return r if isinstance(r, tuple) else (r, None)


def nlspn_propagate(feat_init, guidance, confidence, feat_fix, gamma, prop_kernel=3, prop_time=18,
                    affinity="TGASS", preserve_input=True, always_clip=False, use_offset=True,
                    conf_mode="premul", blend="post", legacy=False, deterministic=False):
    """Functional form.  -> (feat_result, list_feat, offset|None, aff, conf_fixed|None).
    use_offset=False selects the fork's fixed-local propagation (nlspnmodel.py:209-224); guidance
    then holds the N raw affinities only and `offset` is None (as in nlspnmodel.py:306-308)."""
    # a backward can only follow when autograd is recording (eval / no_grad / graph capture keep two src planes)
    need_grad = torch.is_grad_enabled() and any(
        torch.is_tensor(t) and t.requires_grad for t in (feat_init, guidance, confidence, gamma))
    outs = NLSPNFunction.apply(feat_init, guidance, confidence, feat_fix, gamma, prop_kernel,
                               prop_time, affinity, preserve_input, always_clip, use_offset,
                               conf_mode if confidence is not None else "none", blend, legacy, need_grad,
                               deterministic)
    T = prop_time
    list_feat = list(outs[:T])
    aff = outs[T]
    i = T + 1
    offset = None
    if use_offset:
        offset = outs[i]
        i += 1
    conf_fixed = outs[i] if len(outs) > i else None
    return list_feat[-1], list_feat, offset, aff, conf_fixed


class NLSPN(nn.Module):
    """Owns the same parameters, with the same names and shapes, as the reference model's
    propagation state (nlspnmodel.py:93-121): ``aff_scale_const``, ``w``, ``b``, ``w_conf``.

    ``args`` may be the reference's argparse Namespace (attributes prop_kernel, prop_time,
    affinity, affinity_gamma, conf_prop, preserve_input, always_clip) or None with kwargs."""

    def __init__(self, args=None, **kw):
        super().__init__()

        def opt(name, default):
            if name in kw:
                return kw[name]
            return getattr(args, name, default) if args is not None else default

        self.prop_kernel = int(opt("prop_kernel", 3))
        self.prop_time = int(opt("prop_time", 18))
        self.affinity = str(opt("affinity", "TGASS"))
        self.affinity_gamma = float(opt("affinity_gamma", 0.5))
        self.conf_prop = bool(opt("conf_prop", True))
        self.preserve_input = bool(opt("preserve_input", True))
        self.always_clip = bool(opt("always_clip", False))
        # args.offset: deformable gather (True; the north-star path) or the fork's fixed-local 3x3
        # propagation (False; the fork's command-line default, src/config.py:272-275)
        self.offset = bool(opt("offset", True))
        # Fork semantics (the parity target) are the defaults.  conf_mode='sampled' + blend='pre'
        # (+ legacy) select the UPSTREAM semantics the north-star prose describes;
        # those are parity-unpinned: checked only against oracle/torchvision_port.py's restatement.
        self.conf_mode = str(opt("conf_mode", "premul"))
        self.blend = str(opt("blend", "post"))
        self.legacy = bool(opt("legacy", False))
        # bit-identical gradients from run to run (NLSPN_FLAG_DETERMINISTIC; the reference itself is not:
        # deformconv/test.py:627-631); slower backward
        self.deterministic = bool(opt("deterministic", False))
        assert (self.prop_kernel % 2) == 1, \
            'only odd kernel is supported but k_f = {}'.format(self.prop_kernel)   # nlspnmodel.py:29-30
        if self.prop_kernel not in (3, 5, 7):
            raise NotImplementedError("prop_kernel must be 3, 5 or 7")
        self.num_neighbors = self.prop_kernel * self.prop_kernel - 1              # :32
        self.idx_ref = self.num_neighbors // 2                                    # :91
        self.ch_f = 1
        if self.affinity == 'TC':                                                 # :93-104
            self.aff_scale_const = nn.Parameter(self.num_neighbors * torch.ones(1))
            self.aff_scale_const.requires_grad = False
        elif self.affinity == 'TGASS':
            self.aff_scale_const = nn.Parameter(self.affinity_gamma * self.num_neighbors * torch.ones(1))
        elif self.affinity in ('AS', 'ASS'):
            self.aff_scale_const = nn.Parameter(torch.ones(1))
            self.aff_scale_const.requires_grad = False
        else:
            raise NotImplementedError
        # dummy gathering parameters kept for state-dict compatibility (:107-114)
        self.w = nn.Parameter(torch.ones((self.ch_f, 1, self.prop_kernel, self.prop_kernel)))
        self.b = nn.Parameter(torch.zeros(self.ch_f))
        self.w.requires_grad = False
        self.b.requires_grad = False
        self.w_conf = nn.Parameter(torch.ones((1, 1, 1, 1)))
        self.w_conf.requires_grad = False

    def forward(self, feat_init, guidance, confidence=None, feat_fix=None, rgb=None):
        assert self.ch_f == feat_init.shape[1]                                    # :318
        if self.conf_prop:
            assert confidence is not None                                         # :320-321
        else:
            confidence = None
        if self.preserve_input:
            assert feat_fix is not None and feat_init.shape == feat_fix.shape      # :329
        else:
            feat_fix = None
        feat_result, list_feat, offset, aff, _ = nlspn_propagate(
            feat_init, guidance, confidence, feat_fix, self.aff_scale_const, self.prop_kernel,
            self.prop_time, self.affinity, self.preserve_input, self.always_clip, self.offset,
            self.conf_mode, self.blend, self.legacy, self.deterministic)
        return feat_result, list_feat, offset, aff, self.aff_scale_const.data

    def graphed(self, feat_init, guidance, confidence=None, feat_fix=None, warmup=2):
        """Inference forward captured once in a CUDA graph for these shapes (see GraphedNLSPN)."""
        return GraphedNLSPN(self, feat_init, guidance, confidence, feat_fix, warmup=warmup)


class GraphedNLSPN:
    """CUDA-graph replay of the inference forward (no autograd) for one fixed input shape.

    Small batches are launch-bound: one NYU frame's forward is a prologue + one persistent kernel
    (~85 us of GPU time) under ~100 us of Python / allocator / launch glue per call.  The graph holds
    the kernel launches (cooperative and PDL launches are capturable), the static input buffers and
    the output buffers; ``__call__`` copies new inputs into the static buffers (skipped for tensors
    that already ARE the static buffers, ``.inputs``) and replays.  The returned tensors are the
    graph's own output buffers: they are overwritten by the next call.
    """

    def __init__(self, module, feat_init, guidance, confidence=None, feat_fix=None, warmup=2):
        if not feat_init.is_cuda:
            raise RuntimeError("GraphedNLSPN needs CUDA tensors (nlspn_eccv20_b200 has no CPU path)")
        self.module = module
        dev = feat_init.device
        self.inputs = [None if t is None else t.detach().clone().contiguous()
                       for t in (feat_init, guidance, confidence, feat_fix)]
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.no_grad(), torch.cuda.stream(side):
            for _ in range(max(1, warmup)):            # lazy initialisation happens outside the capture
                module(*self.inputs)
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.outputs = module(*self.inputs)

    def __call__(self, feat_init, guidance, confidence=None, feat_fix=None, rgb=None):
        for dst, src in zip(self.inputs, (feat_init, guidance, confidence, feat_fix)):
            if (dst is None) != (src is None):
                raise RuntimeError("GraphedNLSPN: optional inputs must match the captured call")
            if dst is not None and src.data_ptr() != dst.data_ptr():
                if src.shape != dst.shape:
                    raise RuntimeError("GraphedNLSPN: captured for shape %s, got %s" % (tuple(dst.shape), tuple(src.shape)))
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.outputs
