"""Re-run ONE case of tests/test_gpu_reference_cuda.py::test_randomized_sweep... and show where the raw-affinity gradient
differs (test infrastructure: it executes oracle/).  python tests/repro_sweep_case.py CASE [--big]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    target, big = int(sys.argv[1]), "--big" in sys.argv
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import make_inputs
    from oracle import ref_cuda as ref
    ref.load()
    dev = torch.device("cuda:0")
    rnd = random.Random(20201018)
    for case in range(target + 1):
        K = rnd.choice([3, 3, 3, 5, 5, 7])
        T = rnd.randint(1, 8) if K < 7 else rnd.randint(1, 3)
        B = rnd.randint(1, 3)
        H, W = rnd.randint(9, 70), rnd.randint(9, 150)
        if case % 3 == 0:
            W = (W // 4) * 4 + 4
        if big:
            K, T = 3, rnd.randint(1, 6)
            H, W = rnd.randint(8, 260), 4 * rnd.randint(8, 110)
        affinity = rnd.choice(["TGASS", "TGASS", "ASS", "AS", "TC"])
        use_conf, preserve, clip = rnd.random() < 0.7, rnd.random() < 0.7, rnd.random() < 0.3
        sigma = rnd.choice([0.5, 2.0, 2.0, 6.0])
    d = make_inputs(B, H, W, K, max_depth=10.0, seed=1000 + target, conf_mean=3.0, off_sigma=sigma,
                    num_sample=max(1, H * W // 40), device=dev)
    N = K * K - 1
    mod = NLSPN(prop_kernel=K, prop_time=T, affinity=affinity, conf_prop=use_conf, preserve_input=preserve,
                always_clip=clip).to(dev)
    g_out = torch.randn(T, B, 1, H, W, generator=torch.Generator().manual_seed(target)).to(dev)
    leaves = lambda: [d[k].clone().requires_grad_(True) for k in ("feat_init", "guidance", "confidence")]
    fi, gd, cf = leaves()
    out = mod(fi, gd, cf if use_conf else None, d["feat_fix"] if preserve else None)
    torch.autograd.backward(out[1], [g_out[t] for t in range(T)])
    fi2, gd2, cf2 = leaves()
    gam = mod.aff_scale_const.detach().clone().requires_grad_(affinity == "TGASS")
    r = ref.propagate(fi2, gd2, cf2 if use_conf else None, d["feat_fix"] if preserve else None, gam, K, T,
                      affinity=affinity, preserve_input=preserve, always_clip=clip)
    torch.autograd.backward(r["list_feat"], [g_out[t] for t in range(T)])
    a, b = gd.grad[:, 2 * N:], gd2.grad[:, 2 * N:]
    diff = (a - b).abs()
    print("case", target, K, T, B, H, W, affinity, "max|g|", float(b.abs().max()), "max diff", float(diff.max()))
    px = diff.amax(dim=1)                                   # [B,H,W]
    idx = torch.nonzero(px > 1e-4 * b.abs().max())
    print("pixels over tolerance:", idx.shape[0])
    raw = d["guidance"][:, 2 * N:]
    g = float(mod.aff_scale_const) + 1e-8
    for bb, hh, ww in idx[:5].tolist():
        an = torch.tanh(raw[bb, :, hh, ww]) / g
        print((bb, hh, ww), "sum|a|+1e-4 =", "%.9f" % float(an.abs().sum() + 1e-4), "ours", a[bb, :, hh, ww].tolist()[:3],
              "ref", b[bb, :, hh, ww].tolist()[:3])


if __name__ == "__main__":
    main()
