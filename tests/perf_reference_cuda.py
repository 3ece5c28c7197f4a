#!/usr/bin/env python
"""GPU-vs-GPU: the reference propagation (its own DCNv2 CUDA kernels, oracle/_ref/DCN_ref.so, under
the reference's Python op chain restated in oracle/ref_cuda.py) timed on the same B200 next to
ours, same inputs, same metric.  Test/measurement infrastructure (lives under tests/ because it executes oracle/), not the product and not bench.py's
contract arm (that one is the CPU implementation).

    python tests/perf_reference_cuda.py [--workload kitti|nyu] [--batch B] [--kernel K] [--iters T] [--mode fwdbwd|fwd]
prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--workload", default="kitti")
    p.add_argument("--batch", type=int, default=8)
    p.add_argument("--kernel", type=int, default=3)
    p.add_argument("--iters", type=int, default=18)
    p.add_argument("--mode", default="fwdbwd")
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=2)
    a = p.parse_args()
    from nlspn_eccv20_b200 import NLSPN
    from nlspn_eccv20_b200.synth import SHAPES, workload
    from oracle import ref_cuda
    ref_cuda.load()
    dev = torch.device("cuda:0")
    H, W, _ = SHAPES[a.workload]
    K, T, B = a.kernel, a.iters, a.batch
    d = workload(a.workload, B, K, seed=7240, device=dev)
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    train = a.mode == "fwdbwd"
    gam = mod.aff_scale_const.detach().clone().requires_grad_(train)

    def ours():
        fi, gd, cf = (d[k].detach().requires_grad_(train) for k in ("feat_init", "guidance", "confidence"))
        mod.aff_scale_const.grad = None
        with torch.set_grad_enabled(train):
            out = mod(fi, gd, cf, d["feat_fix"])[0]
            loss = (out.clamp(min=0) - d["gt"]).abs().sum()
        if train:
            loss.backward()

    def theirs():
        fi, gd, cf = (d[k].detach().requires_grad_(train) for k in ("feat_init", "guidance", "confidence"))
        gam.grad = None
        with torch.set_grad_enabled(train):
            out = ref_cuda.propagate(fi, gd, cf, d["feat_fix"], gam, K, T)["feat_result"]
            loss = (out.clamp(min=0) - d["gt"]).abs().sum()
        if train:
            loss.backward()

    ms_ref = timed(theirs, a.steps, a.warmup)
    peak_ref = torch.cuda.max_memory_allocated() / 1e9
    torch.cuda.reset_peak_memory_stats()
    ms_ours = timed(ours, a.steps, a.warmup)
    peak_ours = torch.cuda.max_memory_allocated() / 1e9
    pix = B * H * W * T / 1e9
    print(json.dumps({"workload": "%s_%dx%d_B%d_K%d_T%d_%s" % (a.workload, H, W, B, K, T, a.mode),
                      "unit": "Gpix*iter/s",
                      "reference_cuda_kernels_on_b200": pix / (ms_ref * 1e-3), "reference_ms": ms_ref,
                      "ours": pix / (ms_ours * 1e-3), "ours_ms": ms_ours, "speedup": ms_ref / ms_ours,
                      "peak_mem_gb": {"reference": peak_ref, "ours": peak_ours},
                      "note": "reference = its own modulated_deform_conv_cuda.cu compiled for sm_100a "
                              "(oracle/build_ref_cuda.py) under the reference's per-iteration torch op chain"}))


if __name__ == "__main__":
    main()
