"""TEST INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference Python on CPU.

Only usable where /root/reference is mounted (the build container); nothing on
the GPU box imports this file.  It is the tool that (a) generates the golden
vectors under tests/golden/ (see oracle/gen_golden.py) and (b) pins the
restatements in oracle/nlspn_oracle.c and oracle/torchvision_port.py.

How the reference is made to run without a GPU
----------------------------------------------
The reference's propagation calls ``ModulatedDeformConvFunction`` which does
``import DCN`` (reference src/model/modulated_deform_conv_func.py:13), a CUDA-only
pybind module (src/model/deformconv/src/vision.cpp:9-10; the CPU side is a stub,
src/model/deformconv/src/cpu/modulated_deform_cpu.cpp:25,47).  BASELINE.json's
north_star prescribes ``torchvision.ops.deform_conv2d(mask=...)`` as the CPU
stand-in, so a module named ``DCN`` is registered in ``sys.modules`` *before* the
reference is imported; it exposes the two functions with the reference's native
positional signature (src/model/deformconv/src/modulated_deform_conv.h:10-25,46-62).

One correction is applied to the stand-in's backward: the reference returns a
zero offset-gradient whenever a sampling coordinate is <= -1
(modulated_deform_im2col_cuda.cuh:88-92 and :308-311) whereas torchvision uses a
one-sided derivative at exactly -1.  The shim multiplies torchvision's
``grad_offset`` by ``[h_im > -1 and w_im > -1]`` where the coordinates are formed
with the reference's own fp32 expression (``(h - pad + i) + offset``,
modulated_deform_im2col_cuda.cuh:178-179).
"""
from __future__ import annotations

import sys
import types
from argparse import Namespace

import torch

REFERENCE_SRC = "/root/reference/src"


def _coord_valid_mask(offset, kh, kw, sh, sw, ph, pw, dh, dw):
    """[B, 2*kh*kw, Ho, Wo] 0/1 mask: 1 where the reference keeps grad_offset."""
    B, _, Ho, Wo = offset.shape
    dt = offset.dtype
    hs = (torch.arange(Ho) * sh - ph).view(1, 1, Ho, 1)
    ws = (torch.arange(Wo) * sw - pw).view(1, 1, 1, Wo)
    keep = torch.ones_like(offset)
    for i in range(kh):
        for j in range(kw):
            k = i * kw + j
            # integer part first, then one floating add (cuh:178-179)
            h_im = (hs + i * dh).to(dt) + offset[:, 2 * k:2 * k + 1]
            w_im = (ws + j * dw).to(dt) + offset[:, 2 * k + 1:2 * k + 2]
            ok = ((h_im > -1) & (w_im > -1)).to(dt)
            keep[:, 2 * k:2 * k + 1] = ok
            keep[:, 2 * k + 1:2 * k + 2] = ok
    return keep


def make_dcn_shim(correct_minus_one: bool = True) -> types.ModuleType:
    import torchvision  # noqa: F401  (registers torch.ops.torchvision)
    from torchvision.ops import deform_conv2d

    mod = types.ModuleType("DCN")

    def modulated_deform_conv_forward(input, weight, bias, offset, mask,
                                      kh, kw, sh, sw, ph, pw, dh, dw,
                                      group, deformable_group, im2col_step):
        if not (input.is_contiguous() and weight.is_contiguous()):
            raise RuntimeError("input/weight tensor has to be contiguous")
        return deform_conv2d(input, offset, weight, bias, (sh, sw), (ph, pw),
                             (dh, dw), mask)

    def modulated_deform_conv_backward(input, weight, bias, offset, mask, grad_output,
                                       kh, kw, sh, sw, ph, pw, dh, dw,
                                       group, deformable_group, im2col_step):
        gi, gw, go, gm, gb = torch.ops.torchvision._deform_conv2d_backward(
            grad_output.contiguous(), input, weight, offset, mask, bias,
            sh, sw, ph, pw, dh, dw, group, deformable_group, True)
        if correct_minus_one:
            go = go * _coord_valid_mask(offset, kh, kw, sh, sw, ph, pw, dh, dw)
        return gi, go, gm, gw, gb

    mod.modulated_deform_conv_forward = modulated_deform_conv_forward
    mod.modulated_deform_conv_backward = modulated_deform_conv_backward
    return mod


def import_reference(correct_minus_one: bool = True):
    """Return the reference's ``model.nlspnmodel`` module, imported unmodified."""
    sys.modules["DCN"] = make_dcn_shim(correct_minus_one)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import importlib
    return importlib.import_module("model.nlspnmodel")


def reference_args(**kw) -> Namespace:
    """Hand-built args (src/config.py parses sys.argv at import, so it is not used)."""
    d = dict(prop_kernel=3, network="resnet34", from_scratch=True, offset=True,
             zero_init_aff=False, conf_prop=True, affinity="TGASS", affinity_gamma=0.5,
             use_GRU=False, use_S2D=False, lr=1e-3, preserve_input=True,
             always_clip=False, prop_time=18, max_depth=10.0, patch_height=228,
             patch_width=304, GRU_hidden_dim=16, GRU_input_dim=16)
    d.update(kw)
    return Namespace(**d)


def build_reference_model(**kw):
    mod = import_reference()
    return mod.NLSPNModel(reference_args(**kw))


def reference_propagate(model, pred_init, off_raw, aff_raw, confidence, dep):
    """The hot path, nlspnmodel.py:323-377, driven through the reference's OWN methods.

    The statements are the reference's, in the reference's order; only the
    encoder/decoder that produces the five inputs is skipped.  gen_golden.py
    checks this against a real ``NLSPNModel.forward`` with the five inputs
    captured by forward hooks.
    Returns the output-dict entries the path owns.
    """
    a = model.args
    off = model._off_insert(off_raw) if off_raw is not None else None      # :323-324
    aff = model._affinity_normalization(aff_raw)                            # :325
    mask_fix = None
    if a.preserve_input:                                                    # :328-334
        mask_fix = torch.sum(dep > 0.0, dim=1, keepdim=True).detach()
        mask_fix = (mask_fix > 0.0).type_as(dep)
        if confidence is not None:
            confidence = (1.0 - mask_fix) * confidence + mask_fix
    new_pred = pred_init
    list_pred = []
    for k in range(1, a.prop_time + 1):                                     # :340-363
        if k == 1:
            if a.preserve_input:
                new_pred = (1.0 - mask_fix) * new_pred + mask_fix * dep
            if a.always_clip:
                new_pred = torch.clamp(new_pred, min=0)
        if confidence is not None:
            new_pred = model._propagate_once(new_pred * confidence, off, aff)
        else:
            new_pred = model._propagate_once(new_pred, off, aff)
        if a.preserve_input:
            new_pred = (1.0 - mask_fix) * new_pred + mask_fix * dep
        if a.always_clip:
            new_pred = torch.clamp(new_pred, min=0)
        list_pred.append(new_pred)
    feat_result = new_pred
    pred = new_pred if a.always_clip else torch.clamp(new_pred, min=0)      # :375-377
    return dict(pred=pred, feat_result=feat_result, pred_inter=list_pred, offset=off,
                aff=aff, gamma=model.aff_scale_const.data, confidence=confidence)
