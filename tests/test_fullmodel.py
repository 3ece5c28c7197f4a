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
@pytest.mark.parametrize("offset", [True, False])
def test_fork_default_gru_s2d_state_dict_and_heads_identical_to_reference(offset):
    """The fork's default configuration (use_GRU, use_S2D on; src/config.py:225-232): same parameter names /
    order / shapes incl. GRU, encode_*, decode_aff, S2D; the S2D stem and the heads are bit-identical; the
    per-iteration affinity normalisation and offset insertion restated in model.py equal the reference's methods."""
    from oracle import ref_harness
    from nlspn_eccv20_b200.model import NLSPNModel
    torch.manual_seed(0)
    kw = dict(network="resnet18", prop_kernel=3, prop_time=3, use_GRU=True, use_S2D=True, offset=offset,
              patch_height=40, patch_width=56)
    ref = ref_harness.build_reference_model(**kw).eval()
    ours = NLSPNModel(ref_harness.reference_args(**kw)).eval()
    sd = ref.state_dict()
    assert list(sd.keys()) == list(ours.state_dict().keys())
    assert [tuple(v.shape) for v in sd.values()] == [tuple(v.shape) for v in ours.state_dict().values()]
    ours.load_state_dict(sd, strict=True)
    assert [n for n, p in ours.named_parameters() if p.requires_grad] == \
           [n for n, p in ref.named_parameters() if p.requires_grad]
    s = _sample(1, 40, 56)
    cap = {}
    hooks = [ref.id_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("init", o)),
             ref.off_aff_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("guid", o)),
             ref.cf_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("conf", o)),
             ref.S2D.register_forward_hook(lambda m, i, o: cap.__setitem__("s2d", o))]
    with torch.no_grad():
        ref({"rgb": s["rgb"], "dep": s["dep"]})
        pi, gd, cf = ours.heads(s["rgb"], s["dep"])
        assert torch.equal(ours.S2D(s["dep"]), cap["s2d"])
        assert torch.equal(pi, cap["init"]) and torch.equal(gd, cap["guid"]) and torch.equal(cf, cap["conf"])
        raw = torch.randn(2, 8, 9, 11)
        assert torch.equal(ours._normalize_affinity(raw), ref._affinity_normalization(raw))
        off = torch.randn(2, 16, 9, 11)
        assert torch.equal(ours._insert_zero_offset(off), ref._off_insert(off))
        h, x = torch.randn(1, 16, 5, 7), torch.randn(1, 16, 5, 7)
        assert torch.equal(ours.GRU(h=h, x=x), ref.GRU(h=h, x=x))
    for h in hooks:
        h.remove()


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


def test_model_refuses_what_it_does_not_implement():
    from nlspn_eccv20_b200.model import NLSPNModel
    with pytest.raises(NotImplementedError):
        NLSPNModel(use_GRU=True, conf_mode="sampled")      # GRU is a fork feature: fork semantics only
    with pytest.raises(NotImplementedError):
        NLSPNModel(network="resnet50")


def _gru_model_from_fixture(g, dev):
    from nlspn_eccv20_b200.model import NLSPNModel
    H, W = g["in_feat_init"].shape[2:]
    net = NLSPNModel(network="resnet18", prop_kernel=int(g["meta_K"]), prop_time=int(g["meta_T"]),
                     offset=bool(int(g["meta_use_offset"])), use_GRU=True, use_S2D=True, patch_height=H,
                     patch_width=W, max_depth=float(g["meta_max_depth"])).to(dev)
    sd = {k[len("param_"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param_")}
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected and all(m.split(".")[0] not in ("GRU", "encode_aff", "encode_dep", "decode_aff") for m in missing)
    with torch.no_grad():
        net.aff_scale_const.fill_(float(g["meta_gamma"]))
    return net


@pytest.mark.gpu
@pytest.mark.parametrize("name,impl", [("gru_offset_k3_t4", "fused"), ("gru_offset_k3_t4", "dcn"),
                                       ("gru_fixedlocal_k3_t4", "fused")])
def test_fork_gru_loop_matches_reference_fixture(name, impl):
    """the fork's GRU loop (affinities re-estimated between iterations, nlspnmodel.py:365-373).
    The fixture is a forward + backward of the UNMODIFIED reference ``NLSPNModel(use_GRU=True, use_S2D=True)``
    (oracle/gen_golden.py::case_gru).  'fused': one nlspn_step per iteration; 'dcn': the reference's own loop
    statements over the B1 drop-in ``dcn.ModulatedDeformConvFunction`` (what the unmodified fork runs after
    ``dcn.install_as_DCN()``).  Forward 1e-4 m per intermediate state, gradients 1e-4 relative (offset gradients by
    outlier fraction), incl. gamma and the GRU-side parameters."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from conftest import load_golden

    def loss_fn(out, gt):      # the fixed loss of oracle/gen_golden.py (final + intermediate states)
        pred = out["pred"]
        l = (pred - gt).abs().mean() + ((pred - gt) ** 2).mean()
        for i, p in enumerate(out["pred_inter"][:-1]):
            l = l + 0.05 * ((i % 3) + 1) * (p * torch.cos(gt * (i + 1))).mean()
        return l
    dev = torch.device("cuda:0")
    g = load_golden(name)
    net = _gru_model_from_fixture(g, dev)
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    fi, gd, cf = (t(k).requires_grad_(True) for k in ("in_feat_init", "in_guidance", "in_confidence"))
    out = net._forward_gru(fi, gd, cf, t("in_feat_fix"), step_impl=impl)
    lf = torch.stack(out["pred_inter"], 0)
    assert (lf - t("out_list_feat")).abs().max() <= 1e-4
    assert (out["pred"] - t("out_pred")).abs().max() <= 1e-4
    assert (out["aff"] - t("out_aff")).abs().max() <= 1e-5
    if int(g["meta_use_offset"]):
        assert torch.equal(out["offset"], t("out_offset"))
    assert torch.equal(out["confidence"], t("out_conf_fixed"))
    loss = loss_fn(out, t("in_gt"))
    assert abs(float(loss) - float(g["out_loss"])) <= 1e-5 * max(1.0, abs(float(g["out_loss"])))
    loss.backward()
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
    assert rel(fi.grad, t("out_g_feat_init")) < 1e-4
    assert rel(cf.grad, t("out_g_confidence")) < 1e-4
    N = int(g["meta_K"]) ** 2 - 1
    if int(g["meta_use_offset"]):
        assert rel(gd.grad[:, 2 * N:], t("out_g_guidance")[:, 2 * N:]) < 2e-4
        d = (gd.grad[:, :2 * N] - t("out_g_guidance")[:, :2 * N]).abs()
        assert float((d > 1e-4 * t("out_g_guidance")[:, :2 * N].abs().max()).float().mean()) < 2e-3
    else:
        assert rel(gd.grad, t("out_g_guidance")) < 2e-4
    ref_g = float(g["out_g_gamma"].reshape(-1)[0])
    assert abs(float(net.aff_scale_const.grad) - ref_g) <= 2e-4 * max(abs(ref_g), 1e-6)
    for n, p in net.named_parameters():
        if "pgrad_" + n in g:
            assert rel(p.grad, t("pgrad_" + n)) < 5e-4, n


@pytest.mark.gpu
def test_fork_default_configuration_trains():
    """use_GRU + use_S2D + offset=False + prop_time=12 (the fork's command-line defaults): one model, a few steps."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nlspn_eccv20_b200.model import NLSPNModel, NLSPNLoss, train_step
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = NLSPNModel(network="resnet18", prop_kernel=3, prop_time=12, offset=False, use_GRU=True, use_S2D=True,
                     patch_height=64, patch_width=96, max_depth=10.0).to(dev).train()
    s = _sample(2, 64, 96, dev=dev)
    opt = torch.optim.Adam(net.param_groups, lr=1e-3)
    l0, out = train_step(net, NLSPNLoss(10.0), opt, s)
    assert out["offset"] is None and len(out["pred_inter"]) == 12
    fixed = s["dep"] > 0
    assert torch.equal(out["pred"][fixed], s["dep"][fixed])
    for _ in range(5):
        l1, _ = train_step(net, NLSPNLoss(10.0), opt, s)
    assert torch.isfinite(l1) and float(l1) < float(l0)


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
