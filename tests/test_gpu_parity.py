"""GPU (B200): the CUDA path, called through the C ABI, against (1) the golden vectors the
unmodified reference produced, (2) the C oracle on seeded inputs, (3) size-independent
properties at BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): corner indices exact; fp32 depth within 1e-4 m absolute
after 18 iterations on the stable set; RMSE/MAE identical to 1e-5; gradients: relative 1e-4
of the tensor's max, offset gradients by outlier fraction (floor() flips).
"""
import numpy as np
import pytest
import torch

from conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu

PATHS = golden_names("path_") + golden_names("fullmodel_")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nlspn_eccv20_b200 import _lib
    _lib.load()  # raises if the extension is missing: no silent fallback
    return torch.device("cuda:0")


def _meta(g):
    return dict(K=int(g["meta_K"]), T=int(g["meta_T"]), affinity=str(g["meta_affinity"]),
                preserve=bool(int(g["meta_preserve"])), use_conf=bool(int(g["meta_use_conf"])),
                always_clip=bool(int(g.get("meta_always_clip", 0))), gamma=float(g["meta_gamma"]),
                use_offset=bool(int(g.get("meta_use_offset", 1))))


def _run_module(g, dev, grad=False):
    from nlspn_eccv20_b200 import NLSPN
    m = _meta(g)
    mod = NLSPN(prop_kernel=m["K"], prop_time=m["T"], affinity=m["affinity"],
                conf_prop=m["use_conf"], preserve_input=m["preserve"],
                always_clip=m["always_clip"], offset=m["use_offset"]).to(dev)
    with torch.no_grad():
        mod.aff_scale_const.fill_(m["gamma"])
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    fi, gd = t("in_feat_init").requires_grad_(grad), t("in_guidance").requires_grad_(grad)
    cf = t("in_confidence").requires_grad_(grad) if m["use_conf"] else None
    out = mod(fi, gd, cf, t("in_feat_fix"))
    return mod, (fi, gd, cf), out, m


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("name", PATHS)
def test_forward_matches_reference_golden(dev, name):
    g = load_golden(name)
    mod, _, (feat_result, list_feat, offset, aff, gamma), m = _run_module(g, dev)
    if m["use_offset"]:
        assert torch.equal(offset.detach().cpu(), torch.from_numpy(g["out_offset"]))      # offsets exact
    else:
        assert offset is None                                                              # nlspnmodel.py:307
    np.testing.assert_allclose(aff.detach().cpu().numpy(), g["out_aff"], rtol=0, atol=2e-6)
    lf = torch.stack(list_feat, 0).detach().cpu().numpy()
    scale = max(1.0, float(np.abs(g["out_list_feat"]).max()))
    tol = 1e-4 if "signed" not in name and "clip" not in name else 1e-5 * scale
    assert np.abs(lf - g["out_list_feat"]).max() <= tol
    assert np.abs(feat_result.detach().cpu().numpy() - g["out_feat_result"]).max() <= tol
    assert abs(float(gamma) - m["gamma"]) < 1e-6
    if m["preserve"]:
        fix = g["in_feat_fix"] > 0
        for t in range(m["T"]):
            assert np.array_equal(lf[t][fix], g["in_feat_fix"][fix])               # nlspnmodel.py:357
    # RMSE / MAE identical to 1e-5 (src/metric/nlspnmetric.py:53-60)
    from nlspn_eccv20_b200.synth import rmse_mae
    gt = torch.from_numpy(g["in_gt"])
    a = rmse_mae(torch.clamp(feat_result.detach().cpu(), min=0), gt)
    b = rmse_mae(torch.from_numpy(g["out_pred"]), gt)
    assert abs(a[0] - b[0]) <= 1e-5 and abs(a[1] - b[1]) <= 1e-5


@pytest.mark.parametrize("form", ["default", "red", "gather"])
@pytest.mark.parametrize("name", [n for n in PATHS if "fullmodel" not in n])
def test_backward_matches_reference_autograd(dev, nlspn_opt, name, form):
    """`form`: pass A of the backward as the library chooses it (K = 3: the CTA-local transpose, kernels_local.cuh),
    forced to the RED scatter (kernels_v2.cuh) or to the tabulated gather (kernels_gather.cuh) -- every mode of
    the golden set (affinity modes, no confidence, no preserve_input, always_clip, K = 5) goes through all three."""
    if form == "gather":
        nlspn_opt(state_gather=1)
    elif form == "red":
        nlspn_opt(state_gather=0, state_local=0)
    g = load_golden(name)
    mod, (fi, gd, cf), (feat_result, list_feat, offset, aff, _), m = _run_module(g, dev, grad=True)
    gl = torch.from_numpy(g["out_g_list"]).to(dev)
    torch.autograd.backward(list_feat, [gl[t] for t in range(m["T"])])
    N = m["K"] ** 2 - 1
    assert _rel(fi.grad.cpu().numpy(), g["out_g_feat_init"]) < 1e-4
    gg = gd.grad.cpu().numpy()
    if not m["use_offset"]:
        assert _rel(gg, g["out_g_guidance"]) < 2e-4
        if m["use_conf"]:
            assert _rel(cf.grad.cpu().numpy(), g["out_g_confidence"]) < 1e-4
        return
    assert _rel(gg[:, 2 * N:], g["out_g_guidance"][:, 2 * N:]) < 2e-4
    if m["use_conf"]:
        assert _rel(cf.grad.cpu().numpy(), g["out_g_confidence"]) < 1e-4
    if "out_g_gamma" in g:
        ref = float(np.asarray(g["out_g_gamma"]).reshape(-1)[0])
        assert abs(float(mod.aff_scale_const.grad) - ref) <= 2e-4 * max(abs(ref), 1e-6)
    d = np.abs(gg[:, :2 * N] - g["out_g_guidance"][:, :2 * N])
    s = np.abs(g["out_g_guidance"][:, :2 * N]).max()
    assert (d > 1e-4 * s).mean() < 1e-3


@pytest.mark.parametrize("name", golden_names("dcn_"))
def test_dcn_dropin_matches_reference_function(dev, name):
    """Boundary B1: same call the reference makes (nlspnmodel.py:205-208), non-trivial w and b."""
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction
    g = load_golden(name)
    K = int(g["meta_K"])
    t = lambda k: torch.from_numpy(g[k]).to(dev).requires_grad_(True)
    x, off, msk, w, b = t("in_x"), t("in_off"), t("in_msk"), t("in_w"), t("in_b")
    y = ModulatedDeformConvFunction.apply(x, off, msk, w, b, 1, (K - 1) // 2, 1, 1, 1, 64)
    np.testing.assert_allclose(y.detach().cpu().numpy(), g["out_y"], rtol=0, atol=2e-5)
    y.backward(torch.from_numpy(g["in_gout"]).to(dev))
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["out_gx"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(msk.grad.cpu().numpy(), g["out_gmsk"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(off.grad.cpu().numpy(), g["out_goff"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(w.grad.cpu().numpy(), g["out_gw"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(b.grad.cpu().numpy(), g["out_gb"], rtol=1e-4, atol=1e-4)


def test_dcn_domain_errors(dev):
    import nlspn_eccv20_b200.dcn as DCN
    x = torch.zeros(1, 1, 4, 4, device=dev)
    w, b = torch.ones(1, 1, 3, 3, device=dev), torch.zeros(1, device=dev)
    off, msk = torch.zeros(1, 18, 4, 4, device=dev), torch.ones(1, 9, 4, 4, device=dev)
    with pytest.raises(RuntimeError):
        DCN.modulated_deform_conv_forward(x, w, b, off, msk, 3, 3, 2, 2, 1, 1, 1, 1, 1, 1, 64)
    with pytest.raises(RuntimeError, match="contiguous"):
        DCN.modulated_deform_conv_forward(x.expand(1, 1, 4, 4).transpose(2, 3), w, b, off, msk,
                                          3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    with pytest.raises(RuntimeError, match="CPU"):
        DCN.modulated_deform_conv_forward(x.cpu(), w, b, off, msk, 3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    # known answers of the reference's own DCN test (deformconv/test.py:69-110,142-181):
    # zero offsets + identity kernel + mask 0.5  =>  2*out == in
    wi = torch.zeros(1, 1, 3, 3, device=dev); wi[0, 0, 1, 1] = 1.0
    xin = torch.rand(2, 1, 6, 7, device=dev)
    out = DCN.modulated_deform_conv_forward(xin, wi, torch.zeros(1, device=dev),
                                            torch.zeros(2, 18, 6, 7, device=dev),
                                            0.5 * torch.ones(2, 9, 6, 7, device=dev),
                                            3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    assert float((2 * out - xin).abs().max()) < 1e-10
    # zero offsets, mask 1 == nn.Conv2d
    wr = torch.randn(1, 1, 3, 3, device=dev); br = torch.randn(1, device=dev)
    out = DCN.modulated_deform_conv_forward(xin, wr, br, torch.zeros(2, 18, 6, 7, device=dev),
                                            torch.ones(2, 9, 6, 7, device=dev),
                                            3, 3, 1, 1, 1, 1, 1, 1, 1, 1, 64)
    ref = torch.nn.functional.conv2d(xin.cpu(), wr.cpu(), br.cpu(), padding=1)
    assert float((out.cpu() - ref).abs().max()) < 1e-5


@pytest.mark.parametrize("K,B,H,W", [(3, 2, 37, 53), (5, 1, 30, 41), (7, 1, 20, 25)])
def test_corner_indices_exact(dev, oracle, K, B, H, W):
    from nlspn_eccv20_b200 import functional as F_
    g = torch.Generator().manual_seed(11 + K)
    off = 3.0 * torch.randn(B, 2 * K * K, H, W, generator=g)
    off[0, 0, 0, :] = 0.0            # coordinate exactly -1
    off[0, 1, :, 0] = 0.0
    off[0, 2, 3, :] = 1.0            # exact integers
    off[0, 3, :, 5] = -2.0
    off[0, 4:6, 7, :] *= 10.0        # far out of range
    ref = oracle.dcn_debug_indices(off.numpy(), K)
    got = F_.debug_indices(off.to(dev), K).cpu().numpy()
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("K,T,B,H,W,use_conf", [(3, 18, 2, 61, 83, True), (5, 6, 1, 40, 56, True),
                                                (3, 4, 1, 33, 47, False), (7, 3, 1, 24, 31, True)])
def test_against_oracle_seeded(dev, oracle, K, T, B, H, W, use_conf):
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    inp = make_inputs(B, H, W, K, seed=100 + K + T, conf_mean=3.0)
    gamma = 0.5 * (K * K - 1)
    conf = inp["confidence"] if use_conf else None
    ref = oracle.nlspn_forward(inp["feat_init"].numpy(), inp["guidance"].numpy(),
                               conf.numpy() if use_conf else None, inp["feat_fix"].numpy(), gamma, K, T)
    mod = NLSPN(prop_kernel=K, prop_time=T, conf_prop=use_conf).to(dev)
    fi = inp["feat_init"].to(dev).requires_grad_(True)
    gd = inp["guidance"].to(dev).requires_grad_(True)
    cf = conf.to(dev).requires_grad_(True) if use_conf else None
    feat_result, list_feat, offset, aff, _ = mod(fi, gd, cf, inp["feat_fix"].to(dev))
    lf = torch.stack(list_feat, 0).detach().cpu().numpy()
    assert np.abs(lf - ref["list_feat"]).max() <= 1e-4
    assert np.array_equal(offset.detach().cpu().numpy(), ref["offset"])
    np.testing.assert_allclose(aff.detach().cpu().numpy(), ref["aff"], rtol=0, atol=1e-6)
    # backward with a gradient on the final state and on one intermediate state
    gen = torch.Generator().manual_seed(5)
    g_list = np.zeros((T, B, 1, H, W), np.float32)
    g_list[-1] = torch.randn(B, 1, H, W, generator=gen).numpy()
    g_list[T // 2] = torch.randn(B, 1, H, W, generator=gen).numpy()
    rgi, rgg, rgc, rgam = oracle.nlspn_backward(inp["feat_init"].numpy(), inp["guidance"].numpy(),
                                                conf.numpy() if use_conf else None,
                                                inp["feat_fix"].numpy(), gamma, K, ref, g_list)
    torch.autograd.backward([list_feat[-1], list_feat[T // 2]],
                            [torch.from_numpy(g_list[-1]).to(dev), torch.from_numpy(g_list[T // 2]).to(dev)])
    N = K * K - 1
    assert _rel(fi.grad.cpu().numpy(), rgi) < 1e-4
    assert _rel(gd.grad.cpu().numpy()[:, 2 * N:], rgg[:, 2 * N:]) < 2e-4
    if use_conf:
        assert _rel(cf.grad.cpu().numpy(), rgc) < 1e-4
    assert abs(float(mod.aff_scale_const.grad) - rgam) <= 2e-4 * max(abs(rgam), 1e-6)
    d = np.abs(gd.grad.cpu().numpy()[:, :2 * N] - rgg[:, :2 * N])
    assert (d > 1e-4 * np.abs(rgg[:, :2 * N]).max()).mean() < 1e-3


def test_inference_pingpong_equals_training_path(dev):
    """S=2 ping-pong (no grad) and S=T (grad) must give bit-identical states."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    inp = make_inputs(2, 45, 67, 3, seed=3, device=dev)
    mod = NLSPN(prop_kernel=3, prop_time=7).to(dev)
    with torch.no_grad():
        a = mod(inp["feat_init"], inp["guidance"], inp["confidence"], inp["feat_fix"])
    b = mod(inp["feat_init"].clone().requires_grad_(True), inp["guidance"], inp["confidence"], inp["feat_fix"])
    assert torch.equal(a[0], b[0].detach())
    for x, y in zip(a[1], b[1]):
        assert torch.equal(x, y.detach())


@pytest.mark.parametrize("name,B,K,T", [("nyu", 12, 3, 18), ("kitti", 2, 3, 18), ("kitti", 1, 5, 36)])
def test_full_size_properties(dev, name, B, K, T):
    """BASELINE.json sizes: properties that need no oracle run (SURVEY 8c known answers)."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import workload
    inp = workload(name, B, K, device=dev, conf_mean=3.0)
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    with torch.no_grad():
        feat_result, list_feat, offset, aff, _ = mod(inp["feat_init"], inp["guidance"],
                                                     inp["confidence"], inp["feat_fix"])
    ref = (K * K - 1) // 2
    assert not offset[:, 2 * ref:2 * ref + 2].any()                       # centre offset == 0
    assert float((aff.sum(1) - 1).abs().max()) < 1e-5                     # sum_k aff_k == 1
    nb = torch.cat([aff[:, :ref], aff[:, ref + 1:]], 1)
    assert float(nb.abs().sum(1).max()) <= 1.0 + 1e-5                     # sum |a_k| <= 1 (TGASS)
    fix = inp["feat_fix"] > 0
    for x in list_feat:
        assert torch.equal(x[fix], inp["feat_fix"][fix])                  # fixed pixels exact
        assert torch.isfinite(x).all()
    md = float(inp["gt"].max()) * 1.2
    assert float(feat_result.max()) <= md and float(feat_result.min()) >= 0.0   # convex update
    # zero guidance => x_t = blend(x_{t-1} * c): closed form (SURVEY 3.2 known answer)
    with torch.no_grad():
        z = mod(inp["feat_init"], torch.zeros_like(inp["guidance"]), inp["confidence"], inp["feat_fix"])
    m = fix.float()
    c = (1 - m) * inp["confidence"] + m
    x = (1 - m) * inp["feat_init"] + m * inp["feat_fix"]
    for t in range(T):
        x = (1 - m) * (x * c) + m * inp["feat_fix"]
        assert float((z[1][t] - x).abs().max()) <= 1e-5
    # linearity of one step in the state (the operator is linear): G(a*x) == a*G(x)
    import nlspn_eccv20_b200.dcn as DCN
    w, b0 = torch.ones(1, 1, K, K, device=dev), torch.zeros(1, device=dev)
    p = (K - 1) // 2
    y1 = DCN.modulated_deform_conv_forward(inp["feat_init"], w, b0, offset, aff, K, K, 1, 1, p, p, 1, 1, 1, 1, 64)
    y2 = DCN.modulated_deform_conv_forward(2 * inp["feat_init"], w, b0, offset, aff, K, K, 1, 1, p, p, 1, 1, 1, 1, 64)
    assert float((y2 - 2 * y1).abs().max()) <= 1e-4


def test_fused_path_equals_unfused_reference_statements(dev):
    """The unmodified statement sequence of nlspnmodel.py:323-377 written with torch ops on the
    GPU, using OUR DCN drop-in for the gather, must equal the fused module."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction as Fn
    from nlspn_eccv20_b200.synth import make_inputs
    K, T = 3, 6
    inp = make_inputs(2, 40, 52, K, seed=9, device=dev, conf_mean=3.0)
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    with torch.no_grad():
        fr, lf, offset, aff, _ = mod(inp["feat_init"], inp["guidance"], inp["confidence"], inp["feat_fix"])
        dep = inp["feat_fix"]
        m = (dep > 0).float()
        c = (1 - m) * inp["confidence"] + m
        x = (1 - m) * inp["feat_init"] + m * dep
        for t in range(T):
            x = Fn.apply(x * c, offset, aff, mod.w, mod.b, 1, 1, 1, 1, 1, 64)
            x = (1 - m) * x + m * dep
            assert float((x - lf[t]).abs().max()) <= 1e-5


@pytest.mark.parametrize("K,T,use_conf", [(3, 6, True), (5, 3, True), (3, 4, False), (7, 2, True)])
@pytest.mark.parametrize("form", ["red", "local", "gather", "gather-compact"])
def test_two_pass_backward_equals_per_iteration_backward(dev, nlspn_opt, K, T, use_conf, form):
    """Both forms of pass A -- the REDx4 scatter (kernels_v2.cuh; default for K = 3) and the tabulated
    gather (kernels_gather.cuh; default for K >= 5) -- with pass B in registers, against the
    per-iteration formulation (accumulator RMW + scalar atomics) on the same saved tensors."""
    from nlspn_eccv20_b200 import functional as F_
    from nlspn_eccv20_b200.synth import make_inputs
    nlspn_opt(state_gather=0 if form in ("red", "local") else 1, gather_compact=1 if form == "gather-compact" else 0,
              state_local=1 if form == "local" else 0)
    B, H, W = (2, 38, 45) if form != "local" else (2, 38, 44)     # the local form needs W % 4 == 0 (TMA row pitch)
    inp = make_inputs(B, H, W, K, seed=77 + K, device=dev, conf_mean=2.0)
    gamma = 0.5 * (K * K - 1)
    conf = inp["confidence"] if use_conf else None
    offset, aff, cfx, src0 = F_.prologue_fwd(inp["guidance"], conf, inp["feat_init"], inp["feat_fix"], gamma, K)
    S = T if use_conf else 1
    src = torch.empty((S, B, 1, H, W), device=dev)
    src[0].copy_(src0)
    lf = torch.empty((T, B, 1, H, W), device=dev)
    F_.propagate_fwd(offset, aff, cfx, inp["feat_fix"], src, lf, K, T)
    g = torch.Generator().manual_seed(1)
    g_list = [torch.randn(B, 1, H, W, generator=g).to(dev) if t % 2 == 1 or t == T - 1 else None for t in range(T)]
    goe = torch.randn(offset.shape, generator=g).to(dev)
    gae = torch.randn(aff.shape, generator=g).to(dev)
    args = (inp["guidance"], inp["feat_init"], inp["feat_fix"], offset, aff, cfx, src, lf, g_list, gamma, K, T)
    a = F_.backward(*args, g_offset_ext=goe, g_aff_ext=gae)
    b = F_.backward(*args, g_offset_ext=goe, g_aff_ext=gae, per_iteration=True)
    for x, y, name in zip(a, b, ["g_init", "g_guidance", "g_conf", "g_gamma"]):
        if x is None:
            assert y is None
            continue
        s = float(y.abs().max())
        assert float((x - y).abs().max()) <= 2e-5 * max(s, 1e-20), name


@pytest.mark.parametrize("K,T,B,H,W,use_conf,always_clip", [(3, 8, 2, 64, 96, True, False), (5, 5, 2, 48, 61, True, False),
                                                             (3, 4, 1, 37, 45, False, True), (7, 2, 1, 30, 41, True, False)])
def test_deterministic_backward_is_bit_identical_and_matches_default(dev, K, T, B, H, W, use_conf, always_clip):
    """NLSPN_FLAG_DETERMINISTIC (kernels_det.cuh): ten backward calls on the same saved tensors give bit-identical
    gradients (the reference's own backward does not: deformconv/test.py:627-631), and they agree with the default
    (atomic) form within the summation-order tolerance."""
    from nlspn_eccv20_b200 import functional as F_
    from nlspn_eccv20_b200.synth import make_inputs
    inp = make_inputs(B, H, W, K, seed=900 + K, device=dev, conf_mean=2.0)
    gamma = 0.5 * (K * K - 1)
    conf = inp["confidence"] if use_conf else None
    offset, aff, cfx, src0 = F_.prologue_fwd(inp["guidance"], conf, inp["feat_init"], inp["feat_fix"], gamma, K)
    S = T if use_conf else 1
    src = torch.empty((S, B, 1, H, W), device=dev)
    src[0].copy_(src0)
    lf = torch.empty((T, B, 1, H, W), device=dev)
    F_.propagate_fwd(offset, aff, cfx, inp["feat_fix"], src, lf, K, T, always_clip=always_clip)
    g = torch.Generator().manual_seed(2)
    g_list = [torch.randn(B, 1, H, W, generator=g).to(dev) if t % 2 == 0 or t == T - 1 else None for t in range(T)]
    goe = torch.randn(offset.shape, generator=g).to(dev)
    gae = torch.randn(aff.shape, generator=g).to(dev)
    args = (inp["guidance"], inp["feat_init"], inp["feat_fix"], offset, aff, cfx, src, lf, g_list, gamma, K, T)
    kw = dict(g_offset_ext=goe, g_aff_ext=gae, always_clip=always_clip)
    first = F_.backward(*args, deterministic=True, **kw)
    for _ in range(9):
        again = F_.backward(*args, deterministic=True, **kw)
        for x, y, name in zip(first, again, ["g_init", "g_guidance", "g_conf", "g_gamma"]):
            assert (x is None) == (y is None)
            if x is not None:
                assert torch.equal(x, y), name
    ref = F_.backward(*args, **kw)
    for x, y, name in zip(first, ref, ["g_init", "g_guidance", "g_conf", "g_gamma"]):
        if x is None:
            assert y is None
            continue
        s = float(y.abs().max())
        assert float((x - y).abs().max()) <= 2e-5 * max(s, 1e-20), name


def test_deterministic_module_backward_repeats_exactly(dev):
    """The same through the module: NLSPN(deterministic=True), loss.backward() twice -> identical .grad tensors,
    gamma's included."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    K, T = 3, 6
    mod = NLSPN(prop_kernel=K, prop_time=T, deterministic=True).to(dev)
    inp = make_inputs(2, 40, 52, K, seed=41, device=dev, conf_mean=2.0)
    got = []
    for _ in range(3):
        fi, gd, cf = (inp[k].clone().requires_grad_(True) for k in ("feat_init", "guidance", "confidence"))
        mod.aff_scale_const.grad = None
        out = mod(fi, gd, cf, inp["feat_fix"])[0]
        (out.clamp(min=0) - inp["gt"]).abs().sum().backward()
        got.append([fi.grad.clone(), gd.grad.clone(), cf.grad.clone(), mod.aff_scale_const.grad.clone()])
    for other in got[1:]:
        for x, y in zip(got[0], other):
            assert torch.equal(x, y)


def test_non_default_stream_and_concurrent_threads(dev):
    """Boundary B1/B2 threading contract: work is enqueued on the caller's CURRENT
    stream, and two host threads (the reference's test() uses nn.DataParallel worker threads,
    src/main.py:366) may call concurrently.  Results must equal the serial default-stream run."""
    import threading
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    K, T = 3, 5
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    inps = [make_inputs(2, 36, 44, K, seed=500 + i, device=dev, conf_mean=2.0) for i in range(2)]

    def run(inp):
        fi = inp["feat_init"].clone().requires_grad_(True)
        out = mod(fi, inp["guidance"], inp["confidence"], inp["feat_fix"])
        out[0].square().sum().backward()
        return out[0].detach().clone(), fi.grad.clone()

    serial = [run(i) for i in inps]
    torch.cuda.synchronize()
    results = [None, None]

    def worker(i):
        s = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(s):
            results[i] = run(inps[i])
        s.synchronize()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for (a, ga), (b, gb) in zip(serial, results):
        assert torch.equal(a, b)                                       # forward is deterministic
        assert float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max())   # atomics: order only


def test_unusual_inputs(dev):
    """Non-contiguous / expanded inputs, offsets far outside the image, inf-free zeros."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    K, T = 3, 4
    inp = make_inputs(1, 32, 40, K, seed=8, device=dev)
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    with torch.no_grad():
        ref = mod(inp["feat_init"], inp["guidance"], inp["confidence"], inp["feat_fix"])[0]
        # channels-last style non-contiguous guidance must give the same answer
        g_nc = inp["guidance"].permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        assert not g_nc.is_contiguous()
        assert torch.equal(mod(inp["feat_init"], g_nc, inp["confidence"], inp["feat_fix"])[0], ref)
        # every offset far outside: only the centre tap survives => x_t = blend(a_ref * x_{t-1} * c)
        g_far = inp["guidance"].clone()
        g_far[:, :2 * (K * K - 1)] = 1.0e4
        out = mod(inp["feat_init"], g_far, inp["confidence"], inp["feat_fix"])
        aff = out[3]
        m = (inp["feat_fix"] > 0).float()
        c = (1 - m) * inp["confidence"] + m
        x = (1 - m) * inp["feat_init"] + m * inp["feat_fix"]
        for t in range(T):
            x = (1 - m) * (aff[:, 4:5] * (x * c)) + m * inp["feat_fix"]
            assert float((out[1][t] - x).abs().max()) <= 1e-6
    with pytest.raises(RuntimeError):
        mod(inp["feat_init"], inp["guidance"][:, :-1], inp["confidence"], inp["feat_fix"])
    with pytest.raises(RuntimeError):
        mod(inp["feat_init"].double(), inp["guidance"], inp["confidence"], inp["feat_fix"])


def test_persistent_forward_equals_per_iteration_forward(dev, nlspn_opt):
    """One NYU-sized frame takes the persistent cooperative kernel (geometry in registers, grid
    barrier per iteration); NLSPN_PERSIST=0 forces the per-iteration kernels.  Same arithmetic."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import workload
    inp = workload("nyu", 1, 3, device=dev, conf_mean=3.0)
    mod = NLSPN(prop_kernel=3, prop_time=18).to(dev)
    with torch.no_grad():
        a = mod(inp["feat_init"], inp["guidance"], inp["confidence"], inp["feat_fix"])
        nlspn_opt(persist=0)
        b = mod(inp["feat_init"], inp["guidance"], inp["confidence"], inp["feat_fix"])
    for x, y in zip(a[1], b[1]):
        assert float((x - y).abs().max()) <= 2e-6
    # and with gradients (all src planes kept): backward consumes what the persistent kernel saved
    nlspn_opt(persist=-1)
    fi = inp["feat_init"].clone().requires_grad_(True)
    out = mod(fi, inp["guidance"], inp["confidence"], inp["feat_fix"])
    out[0].sum().backward()
    g1 = fi.grad.clone()
    nlspn_opt(persist=0)
    fi2 = inp["feat_init"].clone().requires_grad_(True)
    mod(fi2, inp["guidance"], inp["confidence"], inp["feat_fix"])[0].sum().backward()
    assert float((g1 - fi2.grad).abs().max()) <= 1e-5 * float(g1.abs().max())


@pytest.mark.parametrize("form", ["red", "gather"])
@pytest.mark.parametrize("K,T,legacy,B,H,W", [(3, 6, False, 2, 28, 36), (3, 4, True, 1, 21, 27), (5, 3, False, 1, 24, 32)])
def test_upstream_semantics_against_torchvision_restatement(dev, nlspn_opt, K, T, legacy, B, H, W, form):
    """conf_mode='sampled' + blend='pre' (+legacy): the UPSTREAM semantics of the north-star prose.
    PARITY UNPINNED -- no upstream code is in the reference tree; the referee is the restatement in
    oracle/torchvision_port.py (CPU, autograd).  Forward 1e-4 m, gradients 1e-4 relative."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    from oracle import torchvision_port as TP
    nlspn_opt(state_gather=1 if form == "gather" else 0)     # both forms of pass A
    inp = make_inputs(B, H, W, K, seed=900 + K + T, conf_mean=1.0)
    gamma = 0.5 * (K * K - 1)
    # CPU referee
    fi = inp["feat_init"].clone().requires_grad_(True)
    gd = inp["guidance"].clone().requires_grad_(True)
    cf = inp["confidence"].clone().requires_grad_(True)
    gam = torch.tensor([gamma], requires_grad=True)
    ref = TP.propagate_upstream(fi, gd, cf, inp["feat_fix"], gam, K, T, legacy=legacy)
    gen = torch.Generator().manual_seed(1)
    g_last = torch.randn(B, 1, H, W, generator=gen)
    g_mid = torch.randn(B, 1, H, W, generator=gen)
    torch.autograd.backward([ref["list_feat"][-1], ref["list_feat"][T // 2]], [g_last, g_mid])
    # GPU
    mod = NLSPN(prop_kernel=K, prop_time=T, conf_mode="sampled", blend="pre", legacy=legacy).to(dev)
    fi2 = inp["feat_init"].to(dev).requires_grad_(True)
    gd2 = inp["guidance"].to(dev).requires_grad_(True)
    cf2 = inp["confidence"].to(dev).requires_grad_(True)
    feat_result, list_feat, offset, aff, _ = mod(fi2, gd2, cf2, inp["feat_fix"].to(dev))
    lf = torch.stack(list_feat, 0).detach().cpu()
    assert float((lf - torch.stack(ref["list_feat"], 0).detach()).abs().max()) <= 1e-4
    assert float((aff.detach().cpu() - ref["aff"].detach()).abs().max()) <= 2e-6
    torch.autograd.backward([list_feat[-1], list_feat[T // 2]], [g_last.to(dev), g_mid.to(dev)])
    N = K * K - 1
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))
    assert rel(fi2.grad.cpu(), fi.grad) < 1e-4
    assert rel(cf2.grad.cpu(), cf.grad) < 1e-4
    assert rel(gd2.grad.cpu()[:, 2 * N:], gd.grad[:, 2 * N:]) < 2e-4
    assert abs(float(mod.aff_scale_const.grad) - float(gam.grad)) <= 2e-4 * abs(float(gam.grad))
    d = (gd2.grad.cpu()[:, :2 * N] - gd.grad[:, :2 * N]).abs()
    assert float((d > 1e-4 * gd.grad[:, :2 * N].abs().max()).float().mean()) < 1e-3


@pytest.mark.parametrize("shape,K", [((1, 228, 304), 3), ((2, 352, 1216), 3), ((1, 61, 84), 5)])
def test_cuda_graph_replay_is_bit_identical_to_eager(dev, shape, K):
    """GraphedNLSPN: one capture, replays with new inputs; persistent (one NYU frame), tiled (KITTI) and
    the K=5 paths.  The forward has no atomics, so replay == eager bit for bit."""
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    B, H, W = shape
    mod = NLSPN(prop_kernel=K, prop_time=18).to(dev)
    a = make_inputs(B, H, W, K, seed=1, conf_mean=3.0, device=dev)
    b = make_inputs(B, H, W, K, seed=2, conf_mean=3.0, device=dev)
    args = lambda d: (d["feat_init"], d["guidance"], d["confidence"], d["feat_fix"])
    g = mod.graphed(*args(a))
    for d in (a, b, a):
        with torch.no_grad():
            eager = mod(*args(d))
        out = g(*args(d))
        assert torch.equal(out[0], eager[0])
        assert all(torch.equal(x, y) for x, y in zip(out[1], eager[1]))
        assert torch.equal(out[2], eager[2]) and torch.equal(out[3], eager[3])
    with pytest.raises(RuntimeError):
        g(a["feat_init"][:, :, :-1], a["guidance"][:, :, :-1], a["confidence"][:, :, :-1], a["feat_fix"][:, :, :-1])


def test_gather_form_overflowing_blocks(dev, nlspn_opt):
    """Gather-form pass A with many footprints converging on the same 2x2 block (all taps of a whole region
    point at one pixel: far more than CAP entries per block, so the overflow path carries most of the
    gradient) against the RED-form pass A and the per-iteration backward."""
    from nlspn_eccv20_b200 import functional as F_
    from nlspn_eccv20_b200.synth import make_inputs
    K, T, B, H, W = 3, 4, 1, 24, 32
    inp = make_inputs(B, H, W, K, seed=5, device=dev, conf_mean=2.0)
    gd = inp["guidance"].clone()
    hh = torch.arange(H, device=dev).view(1, 1, H, 1).float()
    ww = torch.arange(W, device=dev).view(1, 1, 1, W).float()
    N = K * K - 1
    for n in range(N):          # every tap of every pixel samples around pixel (10.3, 17.6)
        t = n if n < N // 2 else n + 1
        gd[:, 2 * n] = 10.3 - (hh - 1 + t // K)
        gd[:, 2 * n + 1] = 17.6 - (ww - 1 + t % K)
    gamma = 0.5 * N
    offset, aff, cfx, src0 = F_.prologue_fwd(gd, inp["confidence"], inp["feat_init"], inp["feat_fix"], gamma, K)
    src = torch.empty((T, B, 1, H, W), device=dev)
    src[0].copy_(src0)
    lf = torch.empty((T, B, 1, H, W), device=dev)
    F_.propagate_fwd(offset, aff, cfx, inp["feat_fix"], src, lf, K, T)
    g = torch.Generator().manual_seed(2)
    g_list = [torch.randn(B, 1, H, W, generator=g).to(dev) for _ in range(T)]
    args = (gd, inp["feat_init"], inp["feat_fix"], offset, aff, cfx, src, lf, g_list, gamma, K, T)
    ref = F_.backward(*args, per_iteration=True)
    for form, compact in (("0", "0"), ("1", "0"), ("1", "1")):
        nlspn_opt(state_gather=int(form), gather_compact=int(compact))
        out = F_.backward(*args)
        for x, y, name in zip(out, ref, ["g_init", "g_guidance", "g_conf", "g_gamma"]):
            s = float(y.abs().max())
            assert float((x - y).abs().max()) <= 5e-5 * max(s, 1e-20), (form, name)
