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


def _bias_scale(k, g):
    """Scale for comparing the bias gradient of head k (0 init, 1 guidance, 2 confidence): a sum over all pixels of signed
    terms, so rounding differences of the activations (TF32 forward vs fp32 forward) enter relative to the sum of the
    terms' magnitudes, not to the (cancelling) sum itself."""
    return (1.0, 1.0, 0.25)[k] * float(g[k].abs().sum(dim=(0, 2, 3)).max())


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
    gen = torch.Generator().manual_seed(8)
    g = [torch.randn(t.shape, generator=gen).to(dev) for t in o]
    # pixels whose init pre-activation lies within TF32 rounding of the ReLU kink may take the other branch in the two
    # implementations: no upstream gradient there, so that both backward passes see the same mask
    with torch.no_grad():
        z_init = torch.nn.functional.conv2d(torch.cat((x[0], x[3]), 1), w[0], b[0], 1, 1)
        g[0] = g[0] * (z_init.abs() > 2e-2).to(g[0].dtype)
    torch.autograd.backward(o, g)
    leaves2 = [t.clone().requires_grad_(True) for t in x + w + b]
    r = heads.reference_heads(leaves2[0], leaves2[1], leaves2[2], leaves2[3], leaves2[4], leaves2[7], leaves2[5], leaves2[8], leaves2[6], leaves2[9])
    torch.autograd.backward(r, g)
    for i, (a, c) in enumerate(zip(leaves, leaves2)):
        s = float(c.grad.abs().max().clamp_min(1e-6))
        if i >= 7:                                   # b_id, b_oa, b_cf
            s = max(s, _bias_scale(i - 7, g))
        assert float((a.grad - c.grad).abs().max()) <= 5e-3 * s


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


@pytest.mark.parametrize("B,H,W,K", [(1, 7, 1216, 3), (2, 10, 120, 3), (1, 5, 124, 3), (1, 8, 244, 3), (2, 3, 4, 3),
                                     (1, 1, 8, 3), (1, 2, 12, 3), (2, 67, 132, 3), (1, 6, 360, 5), (1, 4, 128, 5)])
def test_rows_form_matches_the_ninetap_form(B, H, W, K):
    """kernels_head2.cuh (MN-major operands straight from TMA boxes, dx taps on the output side, 120-pixel tiles) against
    kernels_head.cuh (nine re-packed tap tiles) and the fp32 layers: tile seams (W = 120, 124, 244), the last partial
    tile, rows that are not a multiple of the row group, the KITTI width."""
    from nlspn_eccv20_b200 import heads, _lib
    dev = torch.device("cuda:0")
    x, w, b = _case(B, H, W, K, 300 + W, dev)
    args = (x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2], K)
    assert _lib.get_option("heads_rows") == 1 and _lib.get_option("heads_persist") == 1
    rows = heads.fused_heads(*args)
    with _lib.options(heads_persist=0):
        per_tile = heads.fused_heads(*args)
    with _lib.options(heads_ks=1, heads_reuse=0):
        plain = heads.fused_heads(*args)
    for o, t, q in zip(rows, per_tile, plain):     # persistent / one-CTA-per-tile forms, 16- / 8-channel stages, with / without
        assert torch.equal(o, t) and torch.equal(o, q)     # the A-collector hints: the same arithmetic
    with _lib.options(heads_rows=0):
        nine = heads.fused_heads(*args)
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = False
        ref32 = heads.reference_heads(*args[:-1])
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for o, n, r, name in zip(rows, nine, ref32, ("pred_init", "guidance", "confidence")):
        scale = float(r.abs().max().clamp_min(1.0))
        assert float((o - r).abs().max()) <= 4e-3 * scale, name
        assert float((o - n).abs().max()) <= 2e-3 * scale, name       # same TF32 products, another summation order


@pytest.mark.parametrize("affinity,preserve,clip,conf", [("TGASS", True, False, True), ("AS", False, True, True),
                                                         ("TC", True, True, False), ("ASS", True, False, True)])
@pytest.mark.parametrize("B,H,W,K", [(2, 37, 132, 3), (1, 20, 300, 5)])
def test_fused_prologue_equals_heads_then_prologue(affinity, preserve, clip, conf, B, H, W, K):
    """nlspn_heads_prologue_fwd (the prologue as the GEMM's epilogue) against the same GEMM followed by
    nlspn_prologue_fwd on the materialised guidance: offsets bit-identical, the rest to the last bit or two (the same
    expressions compiled into another kernel); with and without writing guidance."""
    from nlspn_eccv20_b200 import heads, functional as F_
    from nlspn_eccv20_b200.synth import make_inputs
    dev = torch.device("cuda:0")
    x, w, b = _case(B, H, W, K, 500 + K, dev)
    fix = make_inputs(B, H, W, K, seed=9, device=dev)["feat_fix"]
    gamma = torch.tensor([0.5 * (K * K - 1)], device=dev)
    args = (x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])
    p, g, c = heads.fused_heads(*args, K)
    off, aff, cfix, src0 = F_.prologue_fwd(g, c if conf else None, p, fix if preserve else None, gamma, K, affinity, preserve, clip)
    for want in (True, False):
        o = heads.fused_heads_prologue(*args, fix if preserve else None, gamma, K, affinity, preserve, clip, conf_prop=conf,
                                       want_guidance=want)
        assert torch.equal(o["pred_init"], p) and torch.equal(o["confidence"], c)
        assert (o["guidance"] is None) != want
        if want:
            assert torch.equal(o["guidance"], g)
        assert torch.equal(o["offset"], off)
        assert float((o["aff"] - aff).abs().max()) <= 2e-7
        assert float((o["src0"] - src0).abs().max()) <= 1e-6 * float(src0.abs().max().clamp_min(1.0))
        if conf:
            assert float((o["conf_fixed"] - cfix).abs().max()) <= 2e-7
        else:
            assert o["conf_fixed"] is None and cfix is None


def test_fused_prologue_rejects_what_it_does_not_implement():
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    x, w, b = _case(1, 8, 18, 3, 1, dev)           # W % 4 != 0
    assert not heads.prologue_supported(18, 3) and not heads.prologue_supported(16, 7) and heads.prologue_supported(16, 3)
    with pytest.raises(RuntimeError):
        heads.fused_heads_prologue(x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2], None, 4.0, 3)


def test_model_inference_with_the_prologue_fused_into_the_heads():
    """NLSPNModel under torch.no_grad(): heads + prologue as one kernel (fused_prologue='auto') against the same model
    with fused_prologue=False (heads kernel, then the propagation's own prologue)."""
    from nlspn_eccv20_b200.model import NLSPNModel
    from nlspn_eccv20_b200.synth import make_inputs
    dev = torch.device("cuda:0")
    torch.manual_seed(4)
    a = NLSPNModel(network="resnet18", prop_kernel=3, prop_time=6, max_depth=10.0, fused_heads=True).to(dev).eval()
    d = make_inputs(2, 60, 84, 3, seed=6, device=dev)
    s = {"rgb": torch.randn(2, 3, 60, 84, device=dev), "dep": d["feat_fix"]}
    with torch.no_grad():
        assert a._use_fused_prologue(s["rgb"])
        oa = a(s)
        a.fused_prologue = False
        assert not a._use_fused_prologue(s["rgb"])
        ob = a(s)
    for k in ("pred", "pred_init", "offset", "aff", "confidence"):
        assert oa[k].shape == ob[k].shape, k
        assert float((oa[k] - ob[k]).abs().max()) <= 1e-5 * float(ob[k].abs().max().clamp_min(1.0)), k
    assert len(oa["pred_inter"]) == len(ob["pred_inter"]) == 6
    a.fused_prologue = "auto"
    assert not a._use_fused_prologue(s["rgb"])            # grad mode: a backward may follow, guidance is needed


def test_heads_random_shapes_against_fp32_layers():
    """Randomized sweep over tile geometry: widths around the 124-pixel tile seams, heights around the 3-row groups,
    batches; K = 3 (three rows per tile) and K = 5 (one row per tile); heads alone and with the fused prologue
    (offsets must be the guidance channels bit for bit)."""
    import random
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    rnd = random.Random(20261019)
    old = torch.backends.cudnn.allow_tf32
    try:
        for case in range(24):
            K = 3 if case % 3 else 5
            B = rnd.choice((1, 1, 2, 3))
            H = rnd.choice((1, 2, 3, 4, 5, 7, 8, 17, 31, 40))
            W = 4 * rnd.choice((1, 2, 29, 30, 31, 32, 33, 61, 62, 63, 93, 94, 100))
            x, w, b = _case(B, H, W, K, 1000 + case, dev)
            args = (x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])
            ours = heads.fused_heads(*args, K)
            torch.backends.cudnn.allow_tf32 = False
            ref32 = heads.reference_heads(*args)
            torch.backends.cudnn.allow_tf32 = old
            for o, r, name in zip(ours, ref32, ("pred_init", "guidance", "confidence")):
                scale = float(r.abs().max().clamp_min(1.0))
                assert float((o - r).abs().max()) <= 4e-3 * scale, (case, K, B, H, W, name)
            fz = heads.fused_heads_prologue(*args, None, 0.5 * (K * K - 1), K, preserve_input=False)
            N = K * K - 1
            g = ours[1]
            ref_off = torch.cat((g[:, :N], torch.zeros_like(g[:, :2]), g[:, N:2 * N]), 1)      # zero pair at the centre tap
            assert torch.equal(fz["offset"], ref_off), (case, K, B, H, W)
            assert float((fz["aff"].sum(1) - 1.0).abs().max()) <= 2e-6, (case, K, B, H, W)
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize("B,H,W,K", [(2, 7, 44, 3), (1, 9, 128, 3), (2, 37, 132, 3), (1, 5, 1216, 3), (1, 20, 300, 5),
                                     (1, 11, 132, 7), (3, 1, 4, 3)])
def test_heads_weight_gradients_on_tcgen05_match_fp32(B, H, W, K):
    """nlspn_heads_grad_prep + nlspn_heads_wgrad (csrc/kernels_head_wgrad.cuh) against what autograd runs for the reference's
    three head layers (nlspnmodel.py:69-86,297,301,313): activation derivatives, concatenation, the shifted copies
    (exact), bias sums and the fp32 weight gradients of the 128-channel layers (TF32 products: <= 3e-3 of the scale)."""
    from nlspn_eccv20_b200 import heads, _lib
    dev = torch.device("cuda:0")
    assert heads.wgrad_supported(W, K)
    x, w, b = _case(B, H, W, K, 11, dev)
    g = torch.Generator().manual_seed(12)
    N3 = 3 * (K * K - 1)
    pred_init = torch.relu(torch.randn(B, 1, H, W, generator=g)).to(dev)
    confidence = torch.sigmoid(torch.randn(B, 1, H, W, generator=g)).to(dev)
    gi, gg, gc = (torch.randn(B, n, H, W, generator=g).to(dev) for n in (1, N3, 1))
    g_shift, g_bias = heads.grad_prep(pred_init, confidence, gi, gg, gc, K)
    g_all = torch.cat((gi * (pred_init > 0), gc * confidence * (1.0 - confidence), gg), 1)      # init, confidence, guidance
    assert float((g_shift[1] - g_all).abs().max()) <= 1e-6
    z = torch.zeros_like(g_all[..., :1])
    assert torch.equal(g_shift[0], torch.cat((g_shift[1][..., 1:], z), -1))          # g[.., x + 1]
    assert torch.equal(g_shift[2], torch.cat((z, g_shift[1][..., :-1]), -1))         # g[.., x - 1]
    bias_ref = g_all.double().sum(dim=(0, 2, 3))
    assert float((g_bias.double() - bias_ref).abs().max()) <= 1e-4 * float(bias_ref.abs().max().clamp_min(1.0))
    dw = heads.weight_grads(x[0], x[1], x[2], x[3], g_shift, K)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        own = (slice(0, 1), slice(2, N3 + 2), slice(1, 2))
        ref = torch.empty_like(dw)
        for k, sl in enumerate(own):
            ref[sl] = torch.nn.grad.conv2d_weight(torch.cat((x[k], x[3]), 1), (sl.stop - sl.start, 128, 3, 3),
                                                  g_all[:, sl].contiguous(), stride=1, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    scale = float(ref.abs().max().clamp_min(1e-6))
    assert float((dw - ref).abs().max()) <= 3e-3 * scale
    # a NULL branch leaves its block zero and the others unchanged; a NULL gradient is a zero gradient
    dw2 = heads.weight_grads(None, x[1], None, x[3], g_shift, K)
    assert float(dw2[:2, :64].abs().max()) == 0.0
    assert float((dw2[2:] - dw[2:]).abs().max()) <= 1e-4 * scale and float((dw2[:2, 64:] - dw[:2, 64:]).abs().max()) <= 1e-4 * scale
    dw3 = heads.weight_grads(x[0], None, None, x[3], g_shift, K)
    assert float(dw3[1:, :64].abs().max()) == 0.0 and float((dw3[0] - dw[0]).abs().max()) <= 1e-4 * scale
    with _lib.options(heads_wgrad_roll=0):           # the per-chunk form
        dw4 = heads.weight_grads(x[0], x[1], x[2], x[3], g_shift, K)
    assert float((dw4 - dw).abs().max()) <= 1e-4 * scale
    s0, b0 = heads.grad_prep(pred_init, confidence, None, gg, None, K)
    assert float(s0[:, :, :2].abs().max()) == 0.0 and torch.equal(s0[1][:, 2:], gg)
    assert float(b0[:2].abs().max()) == 0.0


def test_heads_weight_gradients_reject_what_they_do_not_implement():
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    assert not heads.wgrad_supported(131, 3)
    x, w, b = _case(1, 6, 18, 3, 2, dev)
    a = torch.rand(1, 1, 6, 18, device=dev)
    with pytest.raises(RuntimeError):
        heads.grad_prep(a, a, None, None, None, 3)
    # ... and the autograd Function keeps the stock gradients there
    leaves = [t.clone().requires_grad_(True) for t in x + w + b]
    o = heads.fused_heads(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[7], leaves[5], leaves[8], leaves[6], leaves[9], 3)
    torch.autograd.backward(o, [torch.ones_like(t) for t in o])
    assert all(t.grad is not None and bool(torch.isfinite(t.grad).all()) for t in leaves)


@pytest.mark.parametrize("B,H,W,K", [(2, 7, 44, 3), (1, 1, 4, 3), (1, 33, 1216, 3), (2, 19, 132, 5)])
def test_heads_one_channel_data_gradients_match_fp32(B, H, W, K):
    """nlspn_heads_dgrad_one (fp32 nine-tap stencil) against conv2d_input of the init / confidence layers' own 64-channel
    branches (nlspnmodel.py:69-72,83-86) in fp32."""
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    x, w, b = _case(B, H, W, K, 21, dev)
    g = torch.Generator().manual_seed(22)
    N3 = 3 * (K * K - 1)
    g_all = torch.randn(B, N3 + 2, H, W, generator=g).to(dev)
    d_id, d_cf = heads.dgrad_one(g_all, w[0], w[2], K)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        r_id = torch.nn.grad.conv2d_input((B, 64, H, W), w[0][:, :64].contiguous(), g_all[:, 0:1].contiguous(), stride=1, padding=1)
        r_cf = torch.nn.grad.conv2d_input((B, 64, H, W), w[2][:, :64].contiguous(), g_all[:, 1:2].contiguous(), stride=1, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for a, r in ((d_id, r_id), (d_cf, r_cf)):
        assert float((a - r).abs().max()) <= 1e-5 * float(r.abs().max().clamp_min(1e-6))
    only_cf = heads.dgrad_one(g_all, None, w[2], K)
    assert only_cf[0] is None and torch.equal(only_cf[1], d_cf)


@pytest.mark.parametrize("B,H,W", [(2, 7, 44), (1, 1, 4), (1, 9, 128), (1, 5, 1216), (2, 37, 132), (1, 3, 260)])
def test_heads_wide_data_gradients_on_tcgen05_match_fp32(B, H, W):
    """nlspn_heads_dgrad_wide (csrc/kernels_head_dgrad.cuh) against conv2d_input of the guidance layer's own branch and of
    the fe1 halves of all three layers (nlspnmodel.py:69-86) in fp32: TF32 products over 9 x 26 terms, <= 3e-3 of the scale."""
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    K, N3 = 3, 24
    assert heads.dgrad_supported(W, K) and not heads.dgrad_supported(W, 5)
    x, w, b = _case(B, H, W, K, 31, dev)
    g = torch.Generator().manual_seed(32)
    pred_init = torch.relu(torch.randn(B, 1, H, W, generator=g)).to(dev)
    confidence = torch.sigmoid(torch.randn(B, 1, H, W, generator=g)).to(dev)
    gi, gg, gc = (torch.randn(B, n, H, W, generator=g).to(dev) for n in (1, N3, 1))
    g_shift, _ = heads.grad_prep(pred_init, confidence, gi, gg, gc, K)
    d_oa, d_fe = heads.dgrad_wide(g_shift, w[0], w[1], w[2], K)
    g_all = g_shift[1]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        r_oa = torch.nn.grad.conv2d_input((B, 64, H, W), w[1][:, :64].contiguous(), g_all[:, 2:].contiguous(), stride=1, padding=1)
        w_fe = torch.cat((w[0][:, 64:], w[2][:, 64:], w[1][:, 64:]), 0).contiguous()
        r_fe = torch.nn.grad.conv2d_input((B, 64, H, W), w_fe, g_all.contiguous(), stride=1, padding=1)
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for a, r, name in ((d_oa, r_oa, "oa"), (d_fe, r_fe, "fe1")):
        assert float((a - r).abs().max()) <= 3e-3 * float(r.abs().max().clamp_min(1e-6)), name
    only_fe = heads.dgrad_wide(g_shift, w[0], w[1], w[2], K, want_oa=False)
    assert only_fe[0] is None and torch.equal(only_fe[1], d_fe)


def test_heads_gradient_kernels_random_shapes_against_fp32_autograd():
    """Randomized sweep of the whole native backward (grad_prep, wgrad, dgrad_one, dgrad_wide through FusedHeadsFunction)
    against fp32 autograd through the stock layers: widths around the 32-pixel strip and 128-pixel tile seams, heights around
    the rolling form's row segments (<= 32 rows, two lead-in rows), batches.  The activation masks come from OUR forward
    values in both passes (no upstream gradient within TF32 rounding of the ReLU kink)."""
    import random
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    rnd = random.Random(20261020)
    old = torch.backends.cudnn.allow_tf32
    names = ["id_fd1", "oa_fd1", "cf_fd1", "fe1", "w_id", "w_oa", "w_cf", "b_id", "b_oa", "b_cf"]
    try:
        for case in range(16):
            K = 3 if case % 4 else 5
            B = rnd.choice((1, 1, 2, 3))
            H = rnd.choice((1, 2, 3, 5, 31, 32, 33, 34, 63, 65, 70))
            W = 4 * rnd.choice((1, 2, 7, 8, 9, 31, 32, 33, 63, 64, 65, 97))
            x, w, b = _case(B, H, W, K, 2000 + case, dev)
            order = (0, 1, 2, 3, 4, 7, 5, 8, 6, 9)                     # (x.., w_id, b_id, w_oa, b_oa, w_cf, b_cf)
            leaves = [t.clone().requires_grad_(True) for t in x + w + b]
            o = heads.fused_heads(*[leaves[i] for i in order], K)
            gen = torch.Generator().manual_seed(3000 + case)
            g = [torch.randn(t.shape, generator=gen).to(dev) for t in o]
            with torch.no_grad():
                z = torch.nn.functional.conv2d(torch.cat((x[0], x[3]), 1), w[0], b[0], 1, 1)
                g[0] = g[0] * (z.abs() > 2e-2).to(g[0].dtype)
            torch.autograd.backward(o, g)
            torch.backends.cudnn.allow_tf32 = False
            ref = [t.clone().requires_grad_(True) for t in x + w + b]
            r = heads.reference_heads(*[ref[i] for i in order])
            torch.autograd.backward(r, g)
            torch.backends.cudnn.allow_tf32 = old
            for i, (a, c, name) in enumerate(zip(leaves, ref, names)):
                s = float(c.grad.abs().max().clamp_min(1e-6))
                if i >= 7:
                    s = max(s, _bias_scale(i - 7, g))
                assert float((a.grad - c.grad).abs().max()) <= 4e-3 * s + 1e-6, (case, K, B, H, W, name)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_heads_backward_under_deterministic_algorithms_repeats_exactly():
    """The native weight / bias gradients add split-K partials with fp32 atomics; under
    torch.use_deterministic_algorithms(True) FusedHeadsFunction takes the stock (deterministic) gradient path instead:
    two backward passes are bit-identical, and equal the native path within TF32 rounding."""
    from nlspn_eccv20_b200 import heads
    dev = torch.device("cuda:0")
    K = 3
    x, w, b = _case(2, 40, 96, K, 41, dev)
    order = (0, 1, 2, 3, 4, 7, 5, 8, 6, 9)
    gen = torch.Generator().manual_seed(42)
    gu = None

    def run():
        nonlocal gu
        leaves = [t.clone().requires_grad_(True) for t in x + w + b]
        o = heads.fused_heads(*[leaves[i] for i in order], K)
        if gu is None:
            gu = [torch.randn(t.shape, generator=gen).to(dev) for t in o]
        torch.autograd.backward(o, gu)
        return [t.grad for t in leaves]

    native = run()
    old = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True)
    try:
        a, c = run(), run()
    finally:
        torch.use_deterministic_algorithms(old)
    for p, q, n in zip(a, c, native):
        assert torch.equal(p, q)
        assert float((p - n).abs().max()) <= 5e-3 * float(p.abs().max().clamp_min(1e-6)) + 1e-5
