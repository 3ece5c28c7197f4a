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
