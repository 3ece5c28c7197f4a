"""Drop-in for the reference's native ``DCN`` extension module (boundary B1).

Exposes ``modulated_deform_conv_forward`` / ``modulated_deform_conv_backward`` with the exact
positional signatures of the pybind module (reference src/model/deformconv/src/vision.cpp:9-10,
src/model/deformconv/src/modulated_deform_conv.h:10-25,46-62) so that the UNMODIFIED
``src/model/modulated_deform_conv_func.py`` / ``src/model/nlspnmodel.py`` run on B200:

    import nlspn_eccv20_b200.dcn as dcn; dcn.install_as_DCN()   # before importing the reference

Supported domain: what nlspnmodel.py:107-121,205-208 passes (C=1, groups=1, stride 1, dil 1,
pad (K-1)/2, K in {3,5,7}).  Anything else raises RuntimeError -- never a silent fallback.
Errors mirror the reference: non-contiguous input/weight and CPU tensors raise RuntimeError
(modulated_deform_conv_cuda.cu:39-46, modulated_deform_conv.h:43).
"""
from __future__ import annotations

import sys

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import functional as F_

__all__ = ["modulated_deform_conv_forward", "modulated_deform_conv_backward",
           "ModulatedDeformConvFunction", "install_as_DCN"]


def _check(input, weight, bias, offset, mask):
    for name, t in (("input", input), ("weight", weight), ("bias", bias), ("offset", offset), ("mask", mask)):
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU (%s must be a CUDA tensor)" % name)
    if not input.is_contiguous():
        raise RuntimeError("input tensor has to be contiguous")
    if not weight.is_contiguous():
        raise RuntimeError("weight tensor has to be contiguous")
    if weight.shape[0] != 1 or weight.shape[1] != 1 or input.shape[1] != 1:
        raise RuntimeError("nlspn_eccv20_b200.dcn supports C_in = C_out = 1 only (the NLSPN domain)")


def modulated_deform_conv_forward(input, weight, bias, offset, mask, kernel_h, kernel_w,
                                  stride_h, stride_w, pad_h, pad_w, dilation_h, dilation_w,
                                  group, deformable_group, im2col_step):
    _check(input, weight, bias, offset, mask)
    if tuple(weight.shape[2:4]) != (kernel_h, kernel_w):
        raise RuntimeError("Input shape and kernel shape wont match: (%d x %d vs %d x %d)." %
                           (weight.shape[2], weight.shape[3], kernel_h, kernel_w))
    return F_.dcn_forward(input, weight, bias, offset, mask, kernel_h, kernel_w, stride_h, stride_w,
                          pad_h, pad_w, dilation_h, dilation_w, group, deformable_group, im2col_step)


def modulated_deform_conv_backward(input, weight, bias, offset, mask, grad_output, kernel_h,
                                   kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                   dilation_w, group, deformable_group, im2col_step):
    _check(input, weight, bias, offset, mask)
    gi, go, gm, gw, gb = F_.dcn_backward(input, weight, bias, offset, mask, grad_output, kernel_h,
                                         kernel_w, stride_h, stride_w, pad_h, pad_w, dilation_h,
                                         dilation_w, group, deformable_group, im2col_step)
    return [gi, go, gm, gw, gb]


class ModulatedDeformConvFunction(Function):
    """Mirror of reference src/model/modulated_deform_conv_func.py:15-56 over this module."""

    @staticmethod
    def forward(ctx, input, offset, mask, weight, bias, stride, padding, dilation, groups,
                deformable_groups, im2col_step):
        ctx.stride = _pair(stride)
        ctx.padding = _pair(padding)
        ctx.dilation = _pair(dilation)
        ctx.kernel_size = _pair(weight.shape[2:4])
        ctx.groups = groups
        ctx.deformable_groups = deformable_groups
        ctx.im2col_step = im2col_step
        output = modulated_deform_conv_forward(
            input, weight, bias, offset, mask, ctx.kernel_size[0], ctx.kernel_size[1],
            ctx.stride[0], ctx.stride[1], ctx.padding[0], ctx.padding[1], ctx.dilation[0],
            ctx.dilation[1], ctx.groups, ctx.deformable_groups, ctx.im2col_step)
        ctx.save_for_backward(input, offset, mask, weight, bias)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        input, offset, mask, weight, bias = ctx.saved_tensors
        grad_input, grad_offset, grad_mask, grad_weight, grad_bias = modulated_deform_conv_backward(
            input, weight, bias, offset, mask, grad_output.contiguous(), ctx.kernel_size[0],
            ctx.kernel_size[1], ctx.stride[0], ctx.stride[1], ctx.padding[0], ctx.padding[1],
            ctx.dilation[0], ctx.dilation[1], ctx.groups, ctx.deformable_groups, ctx.im2col_step)
        return grad_input, grad_offset, grad_mask, grad_weight, grad_bias, \
            None, None, None, None, None, None


def install_as_DCN():
    """Register this module as ``DCN`` so ``import DCN`` in the reference resolves here."""
    sys.modules["DCN"] = sys.modules[__name__]
    return sys.modules[__name__]
