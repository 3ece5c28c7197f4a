#!/usr/bin/env python
"""BASELINE config 4: full NLSPN model (ResNet34 encoder-decoder + propagation) training step,
random-init, NYUv2 228x304, batch 12 per GPU, stock DDP + SyncBatchNorm over NCCL (the reference
uses apex DDP + apex SyncBN at opt level O0 = fp32, main.py:133-153), Adam (utility.py:50-73).

    python tests/perf_config4_train_step.py                       # 1 GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/perf_config4_train_step.py

Prints one JSON line (rank 0): step time (CUDA events, max over ranks), images/s, and the
propagation's share of the step, for `--prop ours` (fused op) and -- when oracle/_ref/DCN_ref.so is
present -- the same model with the reference's own CUDA kernels under its per-iteration op chain.
Measurement tool; the bench.py contract line is the propagation path itself.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--batch", type=int, default=12)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--network", default="resnet34")
    p.add_argument("--no-reference", action="store_true")
    a = p.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nlspn_eccv20_b200.model import NLSPNModel, NLSPNLoss, train_step
    from nlspn_eccv20_b200.nlspn import nlspn_propagate
    from nlspn_eccv20_b200.synth import SHAPES, workload
    from oracle import ref_cuda
    H, W, md = SHAPES["nyu"]
    B = a.batch
    d = workload("nyu", B, 3, seed=7240 + rank, device=dev)
    sample = {"rgb": torch.randn(B, 3, H, W, generator=torch.Generator().manual_seed(rank)).to(dev),
              "dep": d["feat_fix"], "gt": d["gt"]}
    loss_fn = NLSPNLoss(md)

    class RefPropModel(NLSPNModel):
        """Same network; propagation by the reference's own CUDA kernels and per-iteration op chain."""
        def forward(self, s):
            pi, gd, cf = self.heads(s["rgb"], s["dep"])
            r = ref_cuda.propagate(pi, gd, cf, s["dep"], self.aff_scale_const, self.prop_kernel, self.prop_time)
            return {"pred": torch.clamp(r["feat_result"], min=0)}

    def run(cls):
        torch.manual_seed(0)
        net = cls(network=a.network, prop_kernel=3, prop_time=18, max_depth=md).to(dev).train()
        if world > 1:
            net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(net)
            net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
        core = net.module if world > 1 else net
        opt = torch.optim.Adam(core.param_groups, lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
        for _ in range(a.warmup):
            train_step(net, loss_fn, opt, sample)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            loss, _ = train_step(net, loss_fn, opt, sample)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # the propagation alone, on this model's own head outputs (fwd + bwd)
        with torch.no_grad():
            pi, gd, cf = core.heads(sample["rgb"], sample["dep"])

        def prop_only():
            x, g, c = (t.detach().requires_grad_(True) for t in (pi, gd, cf))
            core.aff_scale_const.grad = None
            if cls is NLSPNModel:
                out = nlspn_propagate(x, g, c, sample["dep"], core.aff_scale_const, 3, 18)[0]
            else:
                out = ref_cuda.propagate(x, g, c, sample["dep"], core.aff_scale_const, 3, 18)["feat_result"]
            loss_fn(torch.clamp(out, min=0), sample["gt"]).backward()

        for _ in range(2):
            prop_only()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(a.steps):
            prop_only()
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / a.steps
        del net, opt
        torch.cuda.empty_cache()
        return float(ms), pms, float(loss)

    res = {}
    ms, pms, loss = run(NLSPNModel)
    res["ours"] = {"ms_per_step": ms, "images_per_s": world * B / (ms * 1e-3), "propagation_ms": pms,
                   "propagation_share": pms / ms, "loss": loss}
    if ref_cuda.available() and not a.no_reference:
        ms, pms, loss = run(RefPropModel)
        res["reference_cuda_propagation"] = {"ms_per_step": ms, "images_per_s": world * B / (ms * 1e-3),
                                             "propagation_ms": pms, "propagation_share": pms / ms, "loss": loss}
        res["step_speedup"] = ms / res["ours"]["ms_per_step"]
    if rank == 0:
        print(json.dumps({"config": "full NLSPN model train step, %s, NYUv2 %dx%d, batch %d/GPU, K=3, T=18, fp32 "
                                    "(cuDNN TF32 convolutions: torch default), Adam, %s" %
                                    (a.network, H, W, B, "DDP+SyncBN over NCCL" if world > 1 else "single GPU"),
                          "n_gpus": world, "steps": a.steps, "warmup": a.warmup, **res}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
