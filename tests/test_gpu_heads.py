"""SURVEY 8f row f3 (first half): the three final head convolutions as one tcgen05 implicit GEMM
(nlspn_eccv20_b200.heads, csrc/kernels_head.cuh) against the stock torch layers of nlspnmodel.py:69-86,297,301,313."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(B, H, W, K, seed, dev):
    g = torch.Generator().manual_seed(seed)
    N3 = 3 * (K * K - 1)
    x = [torch.randn(B, 64, H, W, generator=g).to(dev) for _ in range(4)]
    s = (128 * 9) ** -0.5
    w = [(s * torch.randn(n, 128, 3, 3, generator=g)).to(dev) for n in (1, N3, 1)]
    b = [(0.1 * torch.randn(n, generator=g)).to(dev) for n in (1, N3, 1)]
    return x, w, b


@pytest.mark.parametrize("B,H,W,K", [(1, 9, 40, 3), (2, 37, 131, 3), (1, 20, 300, 5), (1, 11, 130, 7), (2, 64, 256, 3)])
def test_fused_heads_match_fp32_and_tf32_convolutions(B, H, W, K):
    """Values: within TF32 rounding of the fp32 layers (|d| <= 4e-3 of the output scale; K = 1152 products of 10-bit
    mantissas) and as close to cuDNN's own TF32 result as that is to fp32.  Ragged widths (W % 128 != 0), image
    borders (zero padding), all three activations."""
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    x, w, b = _case(B, H, W, K, 100 + K, dev)
    ours = heads.fused_heads(x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2], K)
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        ref32 = heads.reference_heads(x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])
        torch.backends.cudnn.allow_tf32 = True
        reftf = heads.reference_heads(x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for o, r32, rtf, name in zip(ours, ref32, reftf, ("pred_init", "guidance", "confidence")):
        assert o.shape == r32.shape, name
        scale = float(r32.abs().max().clamp_min(1.0))
        assert float((o - r32).abs().max()) <= 4e-3 * scale, name
        assert float((o - r32).abs().max()) <= 2.0 * float((rtf - r32).abs().max()) + 1e-3 * scale, name


def test_fused_heads_backward_is_the_stock_layers_backward():
    """The backward is stock torch from the saved inputs: gradients equal autograd's through the stock layers fed with
    the same upstream gradients (activation masks come from OUR forward values, hence the tolerance)."""
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    K = 3
    x, w, b = _case(2, 24, 72, K, 7, dev)
    leaves = [t.clone().requires_grad_(True) for t in x + w + b]
    o = heads.fused_heads(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[7], leaves[5], leaves[8], leaves[6], leaves[9], K)
    g = [torch.randn_like(t) for t in o]
    torch.autograd.backward(o, g)
    leaves2 = [t.clone().requires_grad_(True) for t in x + w + b]
    r = heads.reference_heads(leaves2[0], leaves2[1], leaves2[2], leaves2[3], leaves2[4], leaves2[7], leaves2[5], leaves2[8], leaves2[6], leaves2[9])
    torch.autograd.backward(r, g)
    for a, c in zip(leaves, leaves2):
        s = float(c.grad.abs().max().clamp_min(1e-6))
        assert float((a.grad - c.grad).abs().max()) <= 2e-2 * s


def test_fused_heads_reject_what_they_do_not_implement():
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    x, w, b = _case(1, 8, 16, 3, 1, dev)
    with pytest.raises(RuntimeError):
        heads.fused_heads(x[0][:, :32], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2], 3)
    with pytest.raises(RuntimeError):
        heads.fused_heads(x[0].cpu(), x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2], 3)


def test_model_with_fused_heads_matches_stock_heads_and_trains():
    """NLSPNModel(fused_heads=True) against the same weights with fused_heads=False: head outputs within TF32
    rounding, final prediction close, and a few Adam steps reduce the loss with gradients reaching the head weights."""
    from nlspn_eccv20_b200.model import NLSPNModel, NLSPNLoss, train_step
    from nlspn_eccv20_b200.synth import make_inputs
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    a = NLSPNModel(network="resnet18", prop_kernel=3, prop_time=6, max_depth=10.0, fused_heads=True).to(dev).eval()
    b = NLSPNModel(network="resnet18", prop_kernel=3, prop_time=6, max_depth=10.0, fused_heads=False).to(dev).eval()
    b.load_state_dict(a.state_dict())
    d = make_inputs(2, 61, 84, 3, seed=5, device=dev)
    s = {"rgb": torch.randn(2, 3, 61, 84, device=dev), "dep": d["feat_fix"], "gt": d["gt"]}
    with torch.no_grad():
        ha, hb = a.heads(s["rgb"], s["dep"]), b.heads(s["rgb"], s["dep"])
        for x, y, name in zip(ha, hb, ("pred_init", "guidance", "confidence")):
            scale = float(y.abs().max().clamp_min(1.0))
            assert float((x - y).abs().max()) <= 5e-3 * scale, name
        oa, ob = a(s), b(s)
    assert float((oa["pred"] - ob["pred"]).abs().max()) <= 5e-2           # metres, after 6 iterations of a random net
    a.train()
    opt = torch.optim.Adam(a.param_groups, lr=1e-3)
    w0 = a.off_aff_dec0[0].weight.detach().clone()
    l0, _ = train_step(a, NLSPNLoss(10.0), opt, s)
    for _ in range(5):
        l1, _ = train_step(a, NLSPNLoss(10.0), opt, s)
    assert torch.isfinite(l1) and float(l1) < float(l0)
    assert not torch.equal(w0, a.off_aff_dec0[0].weight.detach())
