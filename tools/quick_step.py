"""Quick device-timed step on one GPU (tool, not product): python tools/quick_step.py [--batch B] [--kernel K] [--iters T]
[--workload kitti|nyu] [--mode fwdbwd|fwd] [--opt name=value ...] -> one JSON line with ms/step, Gpix*iter/s and the
per-kernel-class split."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=8)
    p.add_argument("--kernel", type=int, default=3)
    p.add_argument("--iters", type=int, default=18)
    p.add_argument("--workload", default="kitti")
    p.add_argument("--mode", default="fwdbwd")
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--opt", action="append", default=[])
    p.add_argument("--deterministic", action="store_true")
    a = p.parse_args()
    from nlspn_eccv20_b200 import NLSPN, _lib
    from nlspn_eccv20_b200.synth import SHAPES, workload
    lib = _lib.load()
    for o in a.opt:
        k, v = o.split("=")
        _lib.set_option(k, int(v))
    dev = torch.device("cuda:0")
    H, W, _ = SHAPES[a.workload]
    d = workload(a.workload, a.batch, a.kernel, seed=7240, device=dev)
    mod = NLSPN(prop_kernel=a.kernel, prop_time=a.iters, deterministic=a.deterministic).to(dev)
    train = a.mode == "fwdbwd"

    def step():
        fi, gd, cf = (d[k].detach().requires_grad_(train) for k in ("feat_init", "guidance", "confidence"))
        mod.aff_scale_const.grad = None
        with torch.set_grad_enabled(train):
            out = mod(fi, gd, cf, d["feat_fix"])[0]
            loss = (out.clamp(min=0) - d["gt"]).abs().sum()
        if train:
            loss.backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    lib.nlspn_profile_enable(1)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    lib.nlspn_profile_enable(0)
    print(json.dumps({"opts": a.opt + (["deterministic"] if a.deterministic else []), "ms_per_step": round(ms, 4),
                      "gpix_iter_s": round(a.batch * H * W * a.iters / ms / 1e6, 3),
                      "kernel_ms_per_step": {k: round(v[0] / 3, 4) for k, v in prof.items()}}))


if __name__ == "__main__":
    main()
