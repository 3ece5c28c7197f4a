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
unmodified reference when /root/reference is present).

The fork's two additions are covered too (its DEFAULT configuration has both on, src/config.py:225-232):
``use_S2D`` replaces the sparse-depth stem by a min/max-pool pyramid (nlspnmodel.py:406-462; stock torch
layers), and ``use_GRU`` re-estimates the affinities between iterations with a ConvGRU (nlspnmodel.py:123-147,
228-234,365-373,386-403).  With changing affinities the T-iteration kernels do not apply; the loop then runs
one fused native step per iteration (``nlspn_step``: nlspn_step_fwd / nlspn_step_bwd in the C ABI --
deformable or, with ``offset=False`` as in the fork's default, fixed-local), with the small GRU convolutions
and the per-iteration affinity normalisation in stock torch ops.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .nlspn import NLSPN

__all__ = ["NLSPNModel", "NLSPNLoss", "train_step", "ConvGRU", "S2D"]


def _cbr(cin, cout, stride=1, bn=True, relu=True, zero_init=False, kernel=3):
    """kxk conv [+BN] [+ReLU] as an nn.Sequential with the reference's child indices (common.py:45-67)."""
    conv = nn.Conv2d(cin, cout, kernel, stride, (kernel - 1) // 2, bias=not bn)
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


def _up(cin, cout, bn=True, relu=True, zero_init=False):
    """stride-2 3x3 transposed conv [+BN] [+ReLU] (common.py:70-91 as called at nlspnmodel.py:63-67,142-146)."""
    convt = nn.ConvTranspose2d(cin, cout, 3, 2, 1, 1, bias=not bn)
    if zero_init:
        nn.init.zeros_(convt.weight)
        if convt.bias is not None:
            nn.init.zeros_(convt.bias)
    mods = [convt]
    if bn:
        mods.append(nn.BatchNorm2d(cout))
    if relu:
        mods.append(nn.ReLU(inplace=True))
    return nn.Sequential(*mods)


class ConvGRU(nn.Module):
    """Convolutional GRU cell that refines the affinity feature between iterations (nlspnmodel.py:386-403);
    parameter names convz / convr / convq as in the reference."""

    def __init__(self, hidden, inp):
        super().__init__()
        self.convz = nn.Conv2d(hidden + inp, hidden, 3, padding=1)
        self.convr = nn.Conv2d(hidden + inp, hidden, 3, padding=1)
        self.convq = nn.Conv2d(hidden + inp, hidden, 3, padding=1)

    def forward(self, h, x):
        hx = torch.cat((h, x), 1)
        z = torch.sigmoid(self.convz(hx))
        r = torch.sigmoid(self.convr(hx))
        q = torch.tanh(self.convq(torch.cat((r * h, x), 1)))
        return (1 - z) * h + z * q


class S2D(nn.Module):
    """Sparse-to-dense stem (nlspnmodel.py:406-462): min-pools 3/5/7/9 over the valid (non-zero) depths,
    max-pools 11/13, two 1x1 convolutions over the 6-channel pyramid, then a 3x3 convolution with the raw
    depth appended."""

    MIN_POOLS = (3, 5, 7, 9)
    MAX_POOLS = (11, 13)

    def __init__(self):
        super().__init__()
        n = len(self.MIN_POOLS) + len(self.MAX_POOLS)
        self.pool_convs = nn.Sequential(_cbr(n, 8, bn=False, kernel=1), _cbr(8, 16, bn=False, kernel=1))
        self.conv = _cbr(16 + 1, 32, bn=False)

    def forward(self, dep):
        pyr = []
        neg = torch.where(dep == 0, torch.full_like(dep, -999.0), -dep)     # zeros flagged so they never win
        for k in self.MIN_POOLS:
            z = -nn.functional.max_pool2d(neg, k, 1, k // 2)
            pyr.append(torch.where(z == 999, torch.zeros_like(dep), z))
        for k in self.MAX_POOLS:
            pyr.append(nn.functional.max_pool2d(dep, k, 1, k // 2))
        feat = self.pool_convs(torch.cat(pyr, 1))
        return self.conv(torch.cat((feat, dep), 1))


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
        self.use_GRU, self.use_S2D = bool(opt("use_GRU", False)), bool(opt("use_S2D", False))
        # The three final head convolutions (:69-86) as ONE tcgen05 implicit GEMM (heads.py, kernels_head.cuh) instead of
        # three torch.cat + three cuDNN convolutions.  'auto' (default): on CUDA, when the configuration has all three
        # heads with 3N guidance channels and PyTorch's cuDNN TF32 switch is on (the kernel computes in TF32, which is
        # what cuDNN does for these layers then); True: whenever the tensors qualify; False: always the stock layers.
        self.fused_heads = opt("fused_heads", "auto")
        # ... and, when no backward can follow (torch.no_grad / eval inference), with the propagation's prologue as that
        # GEMM's epilogue: the 3N-channel guidance tensor is never written (heads.fused_heads_prologue).  False: never.
        self.fused_prologue = opt("fused_prologue", "auto")
        if self.use_GRU and (self.conf_mode != "premul" or self.blend != "post"):
            raise NotImplementedError("use_GRU is a fork feature: fork semantics (conf_mode='premul', blend='post') only")
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
        if self.use_GRU:                                                   # :123-147
            hid, inp = int(opt("GRU_hidden_dim", 16)), int(opt("GRU_input_dim", 16))
            self.patch_height, self.patch_width = int(opt("patch_height", 228)), int(opt("patch_width", 304))
            self.GRU = ConvGRU(hid, inp)
            self.encode_aff = nn.Sequential(_cbr(N + 1, 16, 2, bn=False), _cbr(16, 2 * hid, 2, bn=False),
                                            _cbr(2 * hid, hid, 2, bn=False, relu=False), nn.Tanh())
            self.encode_dep = nn.Sequential(_cbr(1, 16, 2, bn=False), _cbr(16, 2 * inp, 2, bn=False),
                                            _cbr(2 * inp, inp, 2, bn=False))
            self.decode_aff = nn.Sequential(_up(hid, 2 * hid, bn=False), _up(2 * hid, 16, bn=False),
                                            _up(16, N, bn=False, relu=False, zero_init=bool(opt("zero_init_aff", False))))
        if self.use_S2D:
            self.S2D = S2D()
        # one optimiser group holding every trainable parameter (:152-161)
        self.param_groups = [{"params": [p for p in self.parameters() if p.requires_grad],
                              "lr": float(opt("lr", 1e-3))}]

    def heads(self, rgb, dep, _defer_heads=False):
        """Encoder-decoder up to the three head outputs (nlspnmodel.py:271-315):
        -> (pred_init [B,1,H,W], guidance [B,3N,H,W] (or [B,N,H,W] without offsets), confidence|None).
        ``_defer_heads``: return the head GEMM's arguments instead of running it (forward() fuses it with the
        propagation's prologue), or None when the fused heads do not apply."""
        fe1 = torch.cat((self.conv1_rgb(rgb), self.S2D(dep) if self.use_S2D else self.conv1_dep(dep)), 1)
        fe2 = self.conv2(fe1)
        fe3 = self.conv3(fe2)
        fe4 = self.conv4(fe3)
        fd4 = self.dec4(self.conv5(fe4))
        fd3 = self.dec3(_crop_cat(fd4, fe4))
        fd2 = self.dec2(_crop_cat(fd3, fe3))
        trunk = _crop_cat(fd2, fe2)
        if self._use_fused_heads(fe1):
            from . import heads as H_
            crop = lambda t: t[:, :, :fe1.shape[2], :fe1.shape[3]]
            head_in = (crop(self.id_dec1(trunk)), crop(self.off_aff_dec1(trunk)), crop(self.cf_dec1(trunk)), fe1,
                       self.id_dec0[0].weight, self.id_dec0[0].bias, self.off_aff_dec0[0].weight,
                       self.off_aff_dec0[0].bias, self.cf_dec0[0].weight, self.cf_dec0[0].bias)
            if _defer_heads:
                return head_in
            return H_.fused_heads(*head_in, self.prop_kernel)
        if _defer_heads:
            return None
        pred_init = self.id_dec0(_crop_cat(self.id_dec1(trunk), fe1))
        guidance = self.off_aff_dec0(_crop_cat(self.off_aff_dec1(trunk), fe1))
        confidence = self.cf_dec0(_crop_cat(self.cf_dec1(trunk), fe1)) if self.conf_prop else None
        return pred_init, guidance, confidence

    def _use_fused_heads(self, fe1):
        if self.fused_heads is False or not (self.conf_prop and self.offset):
            return False
        ok = fe1.is_cuda and fe1.dtype == torch.float32 and fe1.shape[1] == 64
        if self.fused_heads == "auto":
            return ok and torch.backends.cudnn.allow_tf32
        if not ok:
            raise RuntimeError("fused_heads=True needs CUDA float32 tensors")
        return True

    # ---- inference: heads + prologue in one kernel, `guidance` never materialised ------------------------
    def _use_fused_prologue(self, rgb):
        """The head GEMM can run the prologue as its epilogue when no backward will follow (the backward re-derives the
        normalisation from `guidance`), in the fork's default semantics, for K = 3, 5 and W % 4 == 0."""
        if self.fused_prologue is False or self.fused_heads is False or self.use_GRU or torch.is_grad_enabled():
            return False
        if not (self.conf_prop and self.offset and self.conf_mode == "premul" and self.blend == "post"):
            return False
        if not (rgb.is_cuda and rgb.dtype == torch.float32):
            return False
        if self.fused_heads == "auto" and not torch.backends.cudnn.allow_tf32:
            return False
        from . import heads as H_
        return H_.prologue_supported(rgb.shape[3], self.prop_kernel)

    def _forward_fused_prologue(self, rgb, dep):
        from . import heads as H_, functional as F_
        head_in = self.heads(rgb, dep, _defer_heads=True)
        if head_in is None:
            return None
        B, _, H, W = head_in[3].shape
        T, K = self.prop_time, self.prop_kernel
        fix = dep if self.preserve_input else None
        src = torch.empty((min(T, 2), B, 1, H, W), device=rgb.device, dtype=torch.float32)
        list_feat = torch.empty((T, B, 1, H, W), device=rgb.device, dtype=torch.float32)
        o = H_.fused_heads_prologue(*head_in, fix, self.aff_scale_const, K, self.affinity, self.preserve_input,
                                    self.always_clip, conf_prop=True, src0=src[0])
        F_.propagate_fwd(o["offset"], o["aff"], o["conf_fixed"], fix, src, list_feat, K, T, self.preserve_input, self.always_clip)
        list_feat = [list_feat[t] for t in range(T)]
        feat_result = list_feat[-1]
        pred = feat_result if self.always_clip else torch.clamp(feat_result, min=0)      # :375-377
        return {"pred": pred, "pred_init": o["pred_init"], "pred_inter": list_feat, "offset": o["offset"], "aff": o["aff"],
                "gamma": self.aff_scale_const.data, "confidence": o["conf_fixed"]}

    # ---- the fork's GRU mode: one native fused step per iteration ---------------------------------------
    def _normalize_affinity(self, raw):
        """nlspnmodel.py:179-201 + :261-269 in torch ops (the GRU mode re-normalises every iteration; the
        fused path does this inside prologue_fwd_kernel)."""
        a = raw
        if self.affinity == "TC":
            a = torch.tanh(a) / self.aff_scale_const
        elif self.affinity == "TGASS":
            a = torch.tanh(a) / (self.aff_scale_const + 1e-8)
        s = a.abs().sum(1, keepdim=True) + 1e-4
        if self.affinity in ("ASS", "TGASS"):
            s = torch.where(s < 1.0, torch.ones_like(s), s)       # the in-place masked assignment of :194
        if self.affinity in ("AS", "ASS", "TGASS"):
            a = a / s
        ref = 1.0 - a.sum(1, keepdim=True)
        return torch.cat((a[:, :self.idx_ref], ref, a[:, self.idx_ref:]), 1)

    def _insert_zero_offset(self, off):
        B, _, H, W = off.shape
        o = off.view(B, self.num_neighbors, 2, H, W)
        z = torch.zeros((B, 1, 2, H, W), dtype=off.dtype, device=off.device)
        return torch.cat((o[:, :self.idx_ref], z, o[:, self.idx_ref:]), 1).view(B, -1, H, W)

    def _forward_gru(self, pred_init, guidance, confidence, dep, step_impl="fused"):
        """nlspnmodel.py:317-381 with use_GRU: the affinities are re-estimated after every iteration but the
        last (:365-373), so each iteration is one ``nlspn_step`` (premultiplied gather + blend + clip).
        step_impl='dcn' runs the reference's own statements (:350-361) over the B1 drop-in operator
        (``dcn.ModulatedDeformConvFunction``) instead -- the path the unmodified fork takes after
        ``dcn.install_as_DCN()``; offsets only."""
        from .nlspn import nlspn_step
        from .dcn import ModulatedDeformConvFunction
        N = self.num_neighbors
        if self.offset:
            off = self._insert_zero_offset(guidance[:, :2 * N]).contiguous()
            aff = self._normalize_affinity(guidance[:, 2 * N:])
        else:
            off, aff = None, self._normalize_affinity(guidance)
        mask_fix = None
        if self.preserve_input:                                                      # :328-334
            mask_fix = (torch.sum(dep > 0.0, dim=1, keepdim=True).detach() > 0.0).type_as(dep)
            if confidence is not None:
                confidence = (1.0 - mask_fix) * confidence + mask_fix
        x = pred_init
        if self.preserve_input:                                                      # :341-348
            x = (1.0 - mask_fix) * x + mask_fix * dep
        if self.always_clip:
            x = torch.clamp(x, min=0)
        src = x * confidence if confidence is not None else x
        list_pred = []
        aff_feat = None
        for k in range(1, self.prop_time + 1):
            if step_impl == "dcn":
                out = ModulatedDeformConvFunction.apply(src, off, aff, self.w, self.b, 1, (self.prop_kernel - 1) // 2,
                                                        1, self.ch_f, 1, 64)              # :205-208
                if self.preserve_input:
                    out = (1.0 - mask_fix) * out + mask_fix * dep                         # :355-357
                if self.always_clip:
                    out = torch.clamp(out, min=0)                                         # :359-361
                src_next = out * confidence if confidence is not None else None           # :351 of the next trip
            else:
                out, src_next = nlspn_step(src, off, aff, confidence, dep if self.preserve_input else None,
                                           self.prop_kernel, self.preserve_input, self.always_clip)
            list_pred.append(out)
            src = src_next if src_next is not None else out
            if k < self.prop_time:                                                   # :365-373
                dep_feat = self.encode_dep(out / self.max_depth)
                if k == 1:
                    aff_feat = self.encode_aff(aff)
                aff_feat = self.GRU(h=aff_feat, x=dep_feat)
                raw = self.decode_aff(aff_feat)[:, :, :self.patch_height, :self.patch_width]      # _clip_as, :237-250
                aff = self._normalize_affinity(raw)
        out = list_pred[-1]
        pred = out if self.always_clip else torch.clamp(out, min=0)
        return {"pred": pred, "pred_init": pred_init, "pred_inter": list_pred, "offset": off, "aff": aff,
                "gamma": self.aff_scale_const.data, "confidence": confidence}

    def forward(self, sample):
        rgb, dep = sample["rgb"], sample["dep"]
        if self._use_fused_prologue(rgb):
            out = self._forward_fused_prologue(rgb, dep)
            if out is not None:
                return out
        pred_init, guidance, confidence = self.heads(rgb, dep)
        if self.use_GRU:
            return self._forward_gru(pred_init, guidance, confidence, dep)
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
