"""NLSPNModel -- the caller on either side of the propagation path (BASELINE config 4).

The reference's full network (src/model/nlspnmodel.py:23-161 constructor, :271-383 forward) is a
ResNet encoder, a shared transposed-convolution decoder and three 3x3 heads (initial depth,
offsets+affinities, confidence) that feed the propagation.  The dense layers are stock cuDNN
convolutions and are expressed here with plain ``torch.nn`` layers; what this repository
contributes is the propagation itself (lines :323-377), which is the parent class ``NLSPN``
(one fused autograd Function over libnlspn_b200.so).

State-dict compatibility is the contract: every parameter and buffer has the reference's name and
shape (``conv1_rgb.0.weight`` ... ``conv2.0.conv1.weight`` (torchvision ResNet block names,
common.py:27-42) ... ``off_aff_dec0.0.weight``, ``aff_scale_const``, ``w``, ``b``, ``w_conf``), so a
reference checkpoint loads with ``strict=True`` (tests/test_fullmodel.py checks this against the
unmodified reference when /root/reference is present).  The fork's ConvGRU / S2D options
(``use_GRU``, ``use_S2D``) re-estimate affinities between iterations and therefore use the
single-step operator (``nlspn_eccv20_b200.dcn``); they are not part of this class.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .nlspn import NLSPN

__all__ = ["NLSPNModel", "NLSPNLoss", "train_step"]


def _cbr(cin, cout, stride=1, bn=True, relu=True, zero_init=False):
    """3x3 conv [+BN] [+ReLU] as an nn.Sequential with the reference's child indices (common.py:45-67)."""
    conv = nn.Conv2d(cin, cout, 3, stride, 1, bias=not bn)
    if zero_init:
        nn.init.zeros_(conv.weight)
        if conv.bias is not None:
            nn.init.zeros_(conv.bias)
    mods = [conv]
    if bn:
        mods.append(nn.BatchNorm2d(cout))
    if relu:
        mods.append(nn.ReLU(inplace=True))
    return nn.Sequential(*mods)


def _up(cin, cout):
    """stride-2 3x3 transposed conv + BN + ReLU (common.py:70-91 as called at nlspnmodel.py:63-67)."""
    return nn.Sequential(nn.ConvTranspose2d(cin, cout, 3, 2, 1, 1, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _Residual(nn.Module):
    """Two-conv residual block; child names follow torchvision's BasicBlock so that ResNet
    checkpoints (pretrained/resnet34.pth, common.py:20-42) and reference checkpoints load."""

    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        y = self.bn2(self.conv2(self.relu(self.bn1(self.conv1(x)))))
        return self.relu(y + (x if self.downsample is None else self.downsample(x)))


def _stage(cin, cout, blocks, stride):
    return nn.Sequential(*[_Residual(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(blocks)])


def _crop_cat(fd, fe):
    """nlspnmodel.py:163-177: the decoder feature may carry one extra row/column; crop, then concat."""
    return torch.cat((fd[:, :, :fe.shape[2], :fe.shape[3]], fe), 1)


class NLSPNModel(NLSPN):
    """``forward(sample)`` with ``sample = {'rgb': [B,3,H,W], 'dep': [B,1,H,W]}`` returns the
    reference's output dict (nlspnmodel.py:379-383): pred, pred_init, pred_inter, offset, aff, gamma,
    confidence."""

    DEPTHS = {"resnet18": (2, 2, 2), "resnet34": (3, 4, 6)}

    def __init__(self, args=None, **kw):
        super().__init__(args, **kw)
        opt = lambda k, d: kw.get(k, getattr(args, k, d) if args is not None else d)
        if opt("use_GRU", False) or opt("use_S2D", False):
            raise NotImplementedError("use_GRU / use_S2D: run the reference model over nlspn_eccv20_b200.dcn.install_as_DCN()")
        network = opt("network", "resnet34")
        if network not in self.DEPTHS:
            raise NotImplementedError(network)
        n1, n2, n3 = self.DEPTHS[network]
        N = self.num_neighbors
        self.max_depth = float(opt("max_depth", 10.0))
        # encoder (nlspnmodel.py:34-58)
        self.conv1_rgb = _cbr(3, 32, bn=False)
        self.conv1_dep = _cbr(1, 32, bn=False)
        self.conv2 = _stage(64, 64, n1, 1)
        self.conv3 = _stage(64, 128, n2, 2)
        self.conv4 = _stage(128, 256, n3, 2)
        self.conv5 = _cbr(256, 256, stride=2)
        # shared decoder (:60-67)
        self.dec4 = _up(256, 128)
        self.dec3 = _up(128 + 256, 64)
        self.dec2 = _up(64 + 128, 64)
        # heads (:69-86)
        self.id_dec1 = _cbr(128, 64)
        self.id_dec0 = _cbr(128, 1, bn=False)
        self.off_aff_dec1 = _cbr(128, 64)
        self.off_aff_dec0 = _cbr(128, (3 if self.offset else 1) * N, bn=False, relu=False,
                                 zero_init=bool(opt("zero_init_aff", False)))
        if self.conf_prop:
            self.cf_dec1 = _cbr(128, 64)
            self.cf_dec0 = nn.Sequential(nn.Conv2d(128, 1, 3, 1, 1), nn.Sigmoid())
        # one optimiser group holding every trainable parameter (:152-161)
        self.param_groups = [{"params": [p for p in self.parameters() if p.requires_grad],
                              "lr": float(opt("lr", 1e-3))}]

    def heads(self, rgb, dep):
        """Encoder-decoder up to the three head outputs (nlspnmodel.py:271-315):
        -> (pred_init [B,1,H,W], guidance [B,3N,H,W] (or [B,N,H,W] without offsets), confidence|None)."""
        fe1 = torch.cat((self.conv1_rgb(rgb), self.conv1_dep(dep)), 1)
        fe2 = self.conv2(fe1)
        fe3 = self.conv3(fe2)
        fe4 = self.conv4(fe3)
        fd4 = self.dec4(self.conv5(fe4))
        fd3 = self.dec3(_crop_cat(fd4, fe4))
        fd2 = self.dec2(_crop_cat(fd3, fe3))
        trunk = _crop_cat(fd2, fe2)
        pred_init = self.id_dec0(_crop_cat(self.id_dec1(trunk), fe1))
        guidance = self.off_aff_dec0(_crop_cat(self.off_aff_dec1(trunk), fe1))
        confidence = self.cf_dec0(_crop_cat(self.cf_dec1(trunk), fe1)) if self.conf_prop else None
        return pred_init, guidance, confidence

    def forward(self, sample):
        rgb, dep = sample["rgb"], sample["dep"]
        pred_init, guidance, confidence = self.heads(rgb, dep)
        # the propagation: one fused op; `dep` is both the fixed-pixel mask source and the preserved values
        from .nlspn import nlspn_propagate
        feat_result, list_feat, offset, aff, conf_fixed = nlspn_propagate(
            pred_init, guidance, confidence, dep if self.preserve_input else None, self.aff_scale_const,
            self.prop_kernel, self.prop_time, self.affinity, self.preserve_input, self.always_clip,
            self.offset, self.conf_mode, self.blend, self.legacy)
        pred = feat_result if self.always_clip else torch.clamp(feat_result, min=0)      # :375-377
        return {"pred": pred, "pred_init": pred_init, "pred_inter": list_feat, "offset": offset, "aff": aff,
                "gamma": self.aff_scale_const.data, "confidence": conf_fixed if conf_fixed is not None else confidence}


class NLSPNLoss(nn.Module):
    """'1.0*L1+1.0*L2' (config.py:96-99) with the reference's masking and per-image normalisation
    (src/loss/submodule/l1loss.py:27-42, l2loss.py:27-42, nlspnloss.py:30-58): both operands are clamped
    to [0, max_depth], pixels with gt <= 1e-4 are ignored, each image's sum is divided by its valid
    count, and the per-image losses are summed (main.py:222 then divides by the batch size)."""

    def __init__(self, max_depth=10.0, w_l1=1.0, w_l2=1.0):
        super().__init__()
        self.max_depth, self.w_l1, self.w_l2 = float(max_depth), float(w_l1), float(w_l2)

    def forward(self, pred, gt):
        gt = torch.clamp(gt, min=0, max=self.max_depth)
        pred = torch.clamp(pred, min=0, max=self.max_depth)
        mask = (gt > 1e-4).to(pred.dtype)
        n = mask.sum(dim=[1, 2, 3]) + 1e-8
        d = pred - gt
        l1 = ((d.abs() * mask).sum(dim=[1, 2, 3]) / n).sum()
        l2 = ((d * d * mask).sum(dim=[1, 2, 3]) / n).sum()
        return self.w_l1 * l1 + self.w_l2 * l2


def train_step(net, loss_fn, optimizer, sample):
    """One iteration of the reference's training loop (main.py:213-228, apex O0 = plain fp32):
    zero_grad, forward, loss / batch size, backward, optimizer step.  `net` may be DDP-wrapped."""
    optimizer.zero_grad(set_to_none=True)
    output = net(sample)
    loss = loss_fn(output["pred"], sample["gt"]) / sample["gt"].shape[0]
    loss.backward()
    optimizer.step()
    return loss.detach(), output
