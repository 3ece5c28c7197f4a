"""f1: the full NLSPNModel -- encoder-decoder + heads in stock torch layers around the
fused propagation.  CPU part: state-dict compatibility with, and head outputs identical to, the
UNMODIFIED reference model (only where /root/reference is mounted; the GPU box skips it).
GPU part: the whole model against heads + the reference's own CUDA kernels, one training step, loss
semantics."""
import os

import numpy as np
import pytest
import torch

HAVE_REF = os.path.isdir("/root/reference/src/model")


def _sample(B, H, W, seed=0, dev="cpu"):
    g = torch.Generator().manual_seed(seed)
    gt = 0.5 + 9.0 * torch.rand(B, 1, H, W, generator=g)
    dep = gt * (torch.rand(B, 1, H, W, generator=g) < 0.02).float()
    rgb = torch.randn(B, 3, H, W, generator=g)
    return {"rgb": rgb.to(dev), "dep": dep.to(dev), "gt": gt.to(dev)}


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
@pytest.mark.parametrize("network,K", [("resnet34", 3), ("resnet18", 5)])
def test_state_dict_and_heads_identical_to_reference(network, K):
    from oracle import ref_harness
    from nlspn_eccv20_b200.model import NLSPNModel
    torch.manual_seed(0)
    ref = ref_harness.build_reference_model(network=network, prop_kernel=K, prop_time=2).eval()
    ours = NLSPNModel(ref_harness.reference_args(network=network, prop_kernel=K, prop_time=2)).eval()
    sd = ref.state_dict()
    assert list(sd.keys()) == list(ours.state_dict().keys())            # same names, same order
    assert [tuple(v.shape) for v in sd.values()] == [tuple(v.shape) for v in ours.state_dict().values()]
    ours.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())
    assert [n for n, p in ours.named_parameters() if p.requires_grad] == \
           [n for n, p in ref.named_parameters() if p.requires_grad]
    s = _sample(1, 45, 61)                                               # odd size: exercises the crop-concat
    cap = {}
    hooks = [ref.id_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("init", o)),
             ref.off_aff_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("guid", o)),
             ref.cf_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("conf", o))]
    with torch.no_grad():
        ref({"rgb": s["rgb"], "dep": s["dep"]})
        pi, gd, cf = ours.heads(s["rgb"], s["dep"])
    for h in hooks:
        h.remove()
    assert torch.equal(pi, cap["init"]) and torch.equal(gd, cap["guid"]) and torch.equal(cf, cap["conf"])


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not mounted")
def test_loss_identical_to_reference():
    import sys
    from argparse import Namespace
    sys.path.insert(0, "/root/reference/src")
    from loss.submodule.l1loss import L1Loss
    from loss.submodule.l2loss import L2Loss
    from nlspn_eccv20_b200.model import NLSPNLoss
    g = torch.Generator().manual_seed(1)
    pred = 12.0 * torch.rand(3, 1, 20, 30, generator=g) - 1.0
    gt = 11.0 * torch.rand(3, 1, 20, 30, generator=g) * (torch.rand(3, 1, 20, 30, generator=g) < 0.7)
    a = Namespace(max_depth=10.0)
    ref = 1.0 * L1Loss(a)(pred, gt) + 1.0 * L2Loss(a)(pred, gt)
    assert torch.allclose(NLSPNLoss(10.0)(pred, gt), ref, rtol=1e-6, atol=0)


def test_model_refuses_gru_and_s2d():
    from nlspn_eccv20_b200.model import NLSPNModel
    with pytest.raises(NotImplementedError):
        NLSPNModel(use_GRU=True)
    with pytest.raises(NotImplementedError):
        NLSPNModel(network="resnet50")


@pytest.mark.gpu
def test_full_model_against_reference_cuda_kernels_and_one_train_step():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import ref_cuda
    from nlspn_eccv20_b200.model import NLSPNModel, NLSPNLoss, train_step
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = NLSPNModel(network="resnet18", prop_kernel=3, prop_time=18, max_depth=10.0).to(dev).train()
    s = _sample(2, 76, 100, dev=dev)
    out = net(s)
    assert out["pred"].shape == s["dep"].shape and len(out["pred_inter"]) == 18
    fixed = s["dep"] > 0
    assert torch.equal(out["pred"][fixed], s["dep"][fixed])               # input preservation through the whole model
    assert float(out["pred"].min()) >= 0.0
    assert torch.equal(out["confidence"][fixed], torch.ones_like(out["confidence"][fixed]))
    if ref_cuda.available():
        # same heads, propagation by the reference's own CUDA kernels: forward and every parameter gradient
        loss_fn = NLSPNLoss(10.0)
        net.zero_grad()
        loss_fn(out["pred"], s["gt"]).backward()
        ours = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
        net.zero_grad()
        pi, gd, cf = net.heads(s["rgb"], s["dep"])
        r = ref_cuda.propagate(pi, gd, cf, s["dep"], net.aff_scale_const, 3, 18)
        pred_ref = torch.clamp(r["feat_result"], min=0)
        assert (out["pred"] - pred_ref).abs().max() <= 1e-4
        loss_fn(pred_ref, s["gt"]).backward()
        for n, p in net.named_parameters():
            if p.grad is None:
                continue
            scale = float(p.grad.abs().max().clamp_min(1e-12))
            assert float((ours[n] - p.grad).abs().max()) <= 2e-3 * scale, n   # dense fp32 convs (TF32 off) both sides
    opt = torch.optim.Adam(net.param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    before = net.aff_scale_const.detach().clone()
    l0, _ = train_step(net, NLSPNLoss(10.0), opt, s)
    for _ in range(5):
        l1, _ = train_step(net, NLSPNLoss(10.0), opt, s)
    assert torch.isfinite(l1) and float(l1) < float(l0)                  # it trains
    assert not torch.equal(before, net.aff_scale_const.detach())         # gamma receives gradient through the fused op
