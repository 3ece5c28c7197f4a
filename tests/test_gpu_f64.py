"""GPU (B200): the double-precision single-step operator (boundary B1) -- the reference dispatches
its op over float and double (modulated_deform_conv_cuda.cu:93,224) and checks gradients with
torch.autograd.gradcheck (src/model/deformconv/test.py:405-433, eps 1e-3, atol 1e-3, rtol 1e-2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from nlspn_eccv20_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


def _inputs(K, B, H, W, seed, dev, far=False):
    g = torch.Generator().manual_seed(seed)
    KK = K * K
    x = torch.rand(B, 1, H, W, generator=g, dtype=torch.float64)
    off = 2.0 * torch.randn(B, 2 * KK, H, W, generator=g, dtype=torch.float64)
    if far:
        off[:, :, 0] = torch.round(off[:, :, 0])       # exact-integer coordinates (incl. -1, H)
        off[:, :, -1] *= 10.0
    msk = torch.sigmoid(torch.rand(B, KK, H, W, generator=g, dtype=torch.float64))
    w = torch.randn(1, 1, K, K, generator=g, dtype=torch.float64)
    b = torch.rand(1, generator=g, dtype=torch.float64)
    return [t.to(dev) for t in (x, off, msk, w, b)]


@pytest.mark.parametrize("K", [3, 5])
def test_gradcheck_as_the_reference_does(dev, K):
    """deformconv/test.py:405-433 (check_gradient_mdconv), same tolerances, on the drop-in Function."""
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction
    x, off, msk, w, b = (t.requires_grad_(True) for t in _inputs(K, 2, 4, 4, 3, dev))
    # keep sampling coordinates away from integers: the bilinear sampler is not differentiable there
    with torch.no_grad():
        frac = off - torch.floor(off)
        off += torch.where(frac < 0.05, 0.1, 0.0) - torch.where(frac > 0.95, 0.1, 0.0)
    fn = lambda *a: ModulatedDeformConvFunction.apply(*a, 1, (K - 1) // 2, 1, 1, 1, 1)
    assert torch.autograd.gradcheck(fn, (x, off, msk, w, b), eps=1e-3, atol=1e-3, rtol=1e-2, raise_exception=True)
    # and much tighter than the reference asks for, with a step that stays inside one bilinear cell
    assert torch.autograd.gradcheck(fn, (x, off, msk, w, b), eps=1e-6, atol=1e-7, rtol=1e-5, raise_exception=True)


@pytest.mark.parametrize("K", [3, 5, 7])
def test_f64_matches_c_oracle(dev, K, oracle):
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction
    x, off, msk, w, b = _inputs(K, 2, 13, 17, 5 + K, dev, far=True)
    gout = torch.randn(2, 1, 13, 17, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(dev)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    y = ModulatedDeformConvFunction.apply(*leaves, 1, (K - 1) // 2, 1, 1, 1, 64)
    y.backward(gout)
    n = lambda t: t.detach().cpu().numpy()
    yo = oracle.dcn_step_fwd(n(x), n(off), n(msk), n(w), n(b))
    gi, go, gm, gw, gb = oracle.dcn_step_bwd(n(x), n(off), n(msk), n(gout), n(w))
    assert y.dtype == torch.float64
    np.testing.assert_allclose(n(y), yo, rtol=0, atol=1e-12)
    np.testing.assert_allclose(n(leaves[0].grad), gi, rtol=0, atol=1e-11)
    np.testing.assert_allclose(n(leaves[1].grad), go, rtol=0, atol=1e-11)
    np.testing.assert_allclose(n(leaves[2].grad), gm, rtol=0, atol=1e-12)
    np.testing.assert_allclose(n(leaves[3].grad), gw, rtol=0, atol=1e-10)
    np.testing.assert_allclose(n(leaves[4].grad).reshape(-1), np.asarray(gb).reshape(-1), rtol=0, atol=1e-10)


def test_f64_matches_reference_cuda_kernels_in_double(dev):
    from oracle import ref_cuda
    if not ref_cuda.available():
        pytest.skip("oracle/_ref/DCN_ref.so not built")
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction
    K = 3
    x, off, msk, w, b = _inputs(K, 2, 21, 30, 9, dev, far=True)
    gout = torch.randn(2, 1, 21, 30, generator=torch.Generator().manual_seed(2), dtype=torch.float64).to(dev)
    ours = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    theirs = [t.clone().requires_grad_(True) for t in (x, off, msk, w, b)]
    y = ModulatedDeformConvFunction.apply(*ours, 1, 1, 1, 1, 1, 64)
    yr = ref_cuda.RefDeformStep.apply(*theirs, K)
    assert (y - yr).abs().max() <= 1e-13
    y.backward(gout)
    yr.backward(gout)
    for a, r_, tol in zip(ours, theirs, (1e-12, 1e-11, 1e-12, 1e-10, 1e-10)):
        assert (a.grad - r_.grad).abs().max() <= tol


def test_other_dtypes_raise(dev):
    from nlspn_eccv20_b200.dcn import ModulatedDeformConvFunction
    x, off, msk, w, b = (t.half() for t in _inputs(3, 1, 4, 4, 0, dev))
    with pytest.raises(RuntimeError):
        ModulatedDeformConvFunction.apply(x, off, msk, w, b, 1, 1, 1, 1, 1, 64)
    x, off, msk, w, b = _inputs(3, 1, 4, 4, 0, dev)
    with pytest.raises(RuntimeError):
        ModulatedDeformConvFunction.apply(x, off.float(), msk, w, b, 1, 1, 1, 1, 1, 64)
