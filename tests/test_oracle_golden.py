"""CPU: pins the C oracle (oracle/nlspn_oracle.c) against the golden vectors that the
UNMODIFIED reference produced (oracle/gen_golden.py).  Tolerances: the referee is the
reference's fp32 run; the oracle restates the same fp32 arithmetic, differing only in
summation order / libm tanh, so forward agrees to a few ulp of the depth scale."""
import numpy as np
import pytest

from conftest import golden_names, load_golden

ALL_PATHS = golden_names("path_") + golden_names("fullmodel_")
PATHS = [n for n in ALL_PATHS if "fixedlocal" not in n]     # the C oracle restates the deformable path only


def _meta(g):
    return dict(K=int(g["meta_K"]), T=int(g["meta_T"]), affinity=str(g["meta_affinity"]),
                preserve=bool(int(g["meta_preserve"])), use_conf=bool(int(g["meta_use_conf"])),
                always_clip=bool(int(g.get("meta_always_clip", 0))), gamma=float(g["meta_gamma"]),
                use_offset=bool(int(g.get("meta_use_offset", 1))))


def _fwd(oracle, g, dtype=np.float32):
    m = _meta(g)
    conf = g["in_confidence"].astype(dtype) if m["use_conf"] else None
    return oracle.nlspn_forward(g["in_feat_init"].astype(dtype), g["in_guidance"].astype(dtype), conf,
                                g["in_feat_fix"].astype(dtype), m["gamma"], m["K"], m["T"],
                                m["affinity"], m["preserve"], m["always_clip"]), m


@pytest.mark.parametrize("name", PATHS)
def test_forward_matches_reference(oracle, name):
    g = load_golden(name)
    out, m = _fwd(oracle, g)
    # offsets are copies: exact
    assert np.array_equal(out["offset"], g["out_offset"])
    # centre pair identically zero (nlspnmodel.py:256)
    ref = (m["K"] ** 2 - 1) // 2
    assert not out["offset"][:, 2 * ref:2 * ref + 2].any()
    np.testing.assert_allclose(out["aff"], g["out_aff"], rtol=0, atol=2e-6)
    if m["use_conf"]:
        assert np.array_equal(out["confidence"], g["out_conf_fixed"])
    scale = max(1.0, float(np.abs(g["out_list_feat"]).max()))
    # north-star bound: 1e-4 m absolute on the stable sets; relative 1e-5 on the signed set
    tol = 1e-4 if "signed" not in name and "clip" not in name else 1e-5 * scale
    err = np.abs(out["list_feat"] - g["out_list_feat"]).max()
    assert err <= tol, (name, err)
    assert np.abs(out["feat_result"] - g["out_feat_result"]).max() <= tol
    # fixed pixels hold the sparse depth exactly after every iteration (nlspnmodel.py:357)
    if m["preserve"]:
        fix = g["in_feat_fix"] > 0
        for t in range(m["T"]):
            assert np.array_equal(out["list_feat"][t][fix], g["in_feat_fix"][fix])


@pytest.mark.parametrize("name", [n for n in PATHS if "fullmodel" not in n and "clip" not in n])
def test_backward_matches_reference_autograd(oracle, name):
    g = load_golden(name)
    out, m = _fwd(oracle, g)
    conf = g["in_confidence"] if m["use_conf"] else None
    gi, gg, gc, ggam = oracle.nlspn_backward(g["in_feat_init"], g["in_guidance"], conf, g["in_feat_fix"],
                                             m["gamma"], m["K"], out, g["out_g_list"], m["affinity"],
                                             m["preserve"])

    def rel(a, b):
        return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)

    N = m["K"] ** 2 - 1
    assert rel(gi, g["out_g_feat_init"]) < 1e-4
    assert rel(gg[:, 2 * N:], g["out_g_guidance"][:, 2 * N:]) < 2e-4          # raw affinities
    if m["use_conf"]:
        assert rel(gc, g["out_g_confidence"]) < 1e-4
    if "out_g_gamma" in g:
        ref = float(np.asarray(g["out_g_gamma"]).reshape(-1)[0])
        assert abs(ggam - ref) <= 2e-4 * max(abs(ref), 1e-6)
    # offset gradients: piecewise-constant derivative, a floor() flip changes single entries
    # completely -- require the bulk to agree and bound the outlier fraction.
    d = np.abs(gg[:, :2 * N] - g["out_g_guidance"][:, :2 * N])
    s = np.abs(g["out_g_guidance"][:, :2 * N]).max()
    assert (d > 1e-4 * s).mean() < 1e-3, (name, (d > 1e-4 * s).mean())


def test_fp64_oracle_close_to_fp32_reference(oracle):
    g = load_golden("path_stable_k3_t18")
    out, m = _fwd(oracle, g, np.float64)
    assert np.abs(out["list_feat"] - g["out_list_feat"]).max() < 1e-3


@pytest.mark.parametrize("name", golden_names("dcn_"))
def test_dcn_step_matches_reference_function(oracle, name):
    g = load_golden(name)
    y = oracle.dcn_step_fwd(g["in_x"], g["in_off"], g["in_msk"], g["in_w"].reshape(-1), g["in_b"])
    np.testing.assert_allclose(y, g["out_y"], rtol=0, atol=2e-5)
    gx, goff, gmsk, gw, gb = oracle.dcn_step_bwd(g["in_x"], g["in_off"], g["in_msk"], g["in_gout"],
                                                 g["in_w"].reshape(-1))
    np.testing.assert_allclose(gx, g["out_gx"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(gmsk, g["out_gmsk"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(goff, g["out_goff"], rtol=0, atol=1e-4)
    np.testing.assert_allclose(gw.reshape(-1), g["out_gw"].reshape(-1), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gb, g["out_gb"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", [n for n in ALL_PATHS if "fullmodel" not in n])
def test_torchvision_port_matches_reference(name):
    """oracle/torchvision_port.py (the CPU-baseline restatement) against the golden vectors."""
    import torch
    from oracle import torchvision_port as TP
    g = load_golden(name)
    m = _meta(g)
    grad = "out_g_list" in g
    t = lambda k: torch.from_numpy(g[k])
    fi, gd = t("in_feat_init").requires_grad_(grad), t("in_guidance").requires_grad_(grad)
    cf = t("in_confidence").requires_grad_(grad) if m["use_conf"] else None
    gam = torch.tensor([m["gamma"]], requires_grad=grad and m["affinity"] == "TGASS")
    out = TP.propagate(fi, gd, cf, t("in_feat_fix") if True else None, gam, m["K"], m["T"],
                       m["affinity"], m["preserve"], m["always_clip"], use_offset=m["use_offset"])
    lf = torch.stack(out["list_feat"], 0).detach().numpy()
    # same ops in the same order as the reference => bit-exact on CPU
    assert np.array_equal(lf, g["out_list_feat"])
    assert np.array_equal(out["aff"].detach().numpy(), g["out_aff"])
    if m["use_offset"]:
        assert np.array_equal(out["offset"].detach().numpy(), g["out_offset"])
    else:
        assert out["offset"] is None
    if grad:
        gl = t("out_g_list")
        torch.autograd.backward(out["list_feat"], [gl[i] for i in range(m["T"])])
        np.testing.assert_allclose(fi.grad.numpy(), g["out_g_feat_init"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(gd.grad.numpy(), g["out_g_guidance"], rtol=1e-4, atol=1e-6)
        if m["use_conf"]:
            np.testing.assert_allclose(cf.grad.numpy(), g["out_g_confidence"], rtol=1e-5, atol=1e-7)
        if "out_g_gamma" in g and gam.grad is not None:
            np.testing.assert_allclose(gam.grad.numpy().reshape(-1), np.asarray(g["out_g_gamma"]).reshape(-1), rtol=1e-4)


def test_torchvision_differs_from_the_reference_only_at_coordinate_minus_one(oracle):
    """the reference returns a ZERO offset gradient when a sampling coordinate is <= -1
    (modulated_deform_im2col_cuda.cuh:88-92,308-311); torchvision's CPU kernel uses a one-sided derivative at
    exactly -1.  The C oracle follows the reference; the stand-in is corrected by the keep-mask of
    oracle/torchvision_port.py.  This test pins both facts: the raw stand-in differs at exactly those
    entries and nowhere else."""
    import torch
    import torchvision  # noqa: F401
    from oracle import torchvision_port as TP
    g = torch.Generator().manual_seed(4)
    B, K, H, W = 2, 3, 7, 9
    KK, pad = K * K, 1
    x = torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
    off = torch.round(3.0 * torch.randn(B, 2 * KK, H, W, generator=g, dtype=torch.float64))   # integer coordinates
    off[:, :, 3:] += 0.37                                                                        # ... on the first rows only
    msk = torch.randn(B, KK, H, W, generator=g, dtype=torch.float64)
    w = torch.ones(1, 1, K, K, dtype=torch.float64)
    b = torch.zeros(1, dtype=torch.float64)
    gout = torch.randn(B, 1, H, W, generator=g, dtype=torch.float64)
    _, _, go_raw, _, _ = torch.ops.torchvision._deform_conv2d_backward(gout, x, w, off, msk, b, 1, 1, pad, pad, 1, 1, 1, 1, True)
    _, go_ref, _, _, _ = oracle.dcn_step_bwd(x.numpy(), off.numpy(), msk.numpy(), gout.numpy(), w.numpy(), want_wb=False)
    keep = TP._keep_mask(off, K, pad).numpy()
    diff = np.abs(go_raw.numpy() - go_ref) > 1e-12
    assert diff.any(), "expected the stand-in to differ somewhere (coordinates of exactly -1 were planted)"
    assert not (diff & (keep == 1)).any()                       # differences only where the reference zeroes
    np.testing.assert_allclose(go_raw.numpy() * keep, go_ref, rtol=0, atol=1e-12)


def test_oracle_known_answers_of_the_reference_dcn_test(oracle):
    """deformconv/test.py:69-110 (zero offset, unit mask == plain convolution) and :142-181 (identity kernel,
    zero offset == input) on the C oracle."""
    import torch
    g = torch.Generator().manual_seed(6)
    for K in (3, 5):
        x = torch.randn(2, 1, 11, 13, generator=g)
        KK = K * K
        off = torch.zeros(2, 2 * KK, 11, 13)
        msk = torch.ones(2, KK, 11, 13)
        w = torch.randn(1, 1, K, K, generator=g)
        b = torch.randn(1, generator=g)
        y = oracle.dcn_step_fwd(x.numpy(), off.numpy(), msk.numpy(), w.numpy(), b.numpy())
        ref = torch.nn.functional.conv2d(x, w, b, padding=(K - 1) // 2).numpy()
        np.testing.assert_allclose(y, ref, rtol=0, atol=2e-6)
        ident = torch.zeros(1, 1, K, K)
        ident[0, 0, K // 2, K // 2] = 1.0
        y = oracle.dcn_step_fwd(x.numpy(), off.numpy(), msk.numpy(), ident.numpy(), np.zeros(1, np.float32))
        assert np.array_equal(y, x.numpy())
