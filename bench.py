#!/usr/bin/env python
"""bench.py -- NLSPN propagation throughput on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch shard: fused prologue + T propagation
iterations (forward) and the fused reverse replay (backward), K=3, T=18 on KITTI 352x1216 frames
(the shape BASELINE.json's metric is quoted on), B frames per GPU (weak scaling: batch shards,
no collective on the data path -- SURVEY 8e).

metric  Gpix*iter/s = (GPUs * B * H * W * T) / seconds, forward+backward.
value   device-timed (CUDA events, max over ranks), inputs resident in HBM.
e2e     same metric through the public module with PINNED HOST inputs: H2D of the step's
        inputs and D2H of the step's loss inside the timed region.
roofline  dominant kernel = the one with the largest share of the step (K=3: pass A of the backward,
        bwd_state_kernel, 12N+40 B/px per launch; K>=5: bwd_gather_kernel, 16N+20 B/px);
        achieved = its algorithmic bytes per launch / its average launch duration (CUDA-event time of
        its phase in the timed region x its share of that phase from an event-bracketed pass);
        peak = MEASURED_PEAKS.json hbm_gbs; traffic = measured DRAM bytes per launch (profiles/ncu_traffic.json).
        roofline_step = SURVEY 8d's algorithmic bytes of the whole step against the same peak.
cpu_baseline  the reference path restated over torchvision.ops.deform_conv2d (north_star's CPU
        stand-in; oracle/torchvision_port.py, pinned against the unmodified reference), one image
        per host thread, on a bounded sample of the same workload.  Rank 0, N=1 only.
--impl reference  times that CPU implementation alone (rank 0), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "nlspn_propagation_fwd_bwd_gpix_iter_per_s"
UNIT = "Gpix*iter/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="kitti", choices=["kitti", "nyu"])
    p.add_argument("--batch", type=int, default=None, help="frames per GPU (default: kitti 8, nyu 12)")
    p.add_argument("--kernel", type=int, default=3)
    p.add_argument("--iters", type=int, default=18)
    p.add_argument("--mode", default="fwdbwd", choices=["fwdbwd", "fwd"])
    p.add_argument("--smooth-offsets", action="store_true")
    p.add_argument("--flush-l2", default="auto", choices=["auto", "on", "off"],
                   help="write a 256 MB buffer between timed steps (auto: when the inputs fit L2)")
    p.add_argument("--cuda-graph", action="store_true",
                   help="forward-only runs: replay the module's forward from a CUDA graph (GraphedNLSPN)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-images", type=int, default=None, help="frames in the CPU sample (default: min(cores, 8))")
    p.add_argument("--cpu-rows", type=int, default=176,
                   help="rows of each frame in the CPU sample (full width; bounds the CPU leg's run time)")
    return p.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def alg_bytes(K, T, mode):
    """Algorithmic bytes per pixel*iteration (fork semantics, fp32)."""
    N = K * K - 1
    fwd = 12 * N + 20 + (24 * N + 28) / T
    bwd = 36 * N + 36 + (28 * N + 12) / T
    return fwd if mode == "fwd" else fwd + bwd


class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        n = self._nvml
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.004)

    def start(self):
        if self._nvml is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=2)
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_run(args, H, W, md, n_images, steps, warmup):
    """Times oracle/torchvision_port.py (the reference path over torchvision) on host cores.
    Sample: n_images frames of min(H, --cpu-rows) rows x full width, one frame per host thread."""
    H = min(H, args.cpu_rows)
    import torch
    from nlspn_eccv20_b200.synth import make_inputs
    from oracle import torchvision_port as TP
    cores = os.cpu_count() or 1
    workers = min(cores, n_images)
    K, T = args.kernel, args.iters
    gamma = 0.5 * (K * K - 1)
    kw = dict(num_sample=500) if args.workload == "nyu" else dict(density=0.05)
    images = []
    for i in range(n_images):
        d = make_inputs(1, H, W, K, max_depth=md, seed=7240 + i, **kw)
        images.append(d)
    backward = args.mode == "fwdbwd"
    times = []
    for s in range(warmup + steps):
        dt = TP.time_batch_parallel(images, gamma, K, T, backward=backward, workers=workers)
        if s >= warmup:
            times.append(dt)
    torch.set_num_threads(cores)
    best = min(times)
    mean = sum(times) / len(times)
    pix_iter = n_images * H * W * T
    return dict(value=pix_iter / mean / 1e9, best=pix_iter / best / 1e9, seconds=mean, cores=workers,
                host_cores=cores, n_images=n_images, rows=H)


def main():
    args = parse()
    import torch
    from nlspn_eccv20_b200.synth import SHAPES
    H, W, md = SHAPES[args.workload]
    K, T = args.kernel, args.iters
    B = args.batch if args.batch is not None else (8 if args.workload == "kitti" else 12)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": "%s_%dx%d_B%d_per_gpu_K%d_T%d_%s" % (args.workload, H, W, B, K, T, args.mode),
              "frames_per_gpu": B, "height": H, "width": W, "prop_kernel": K, "prop_time": T,
              "affinity": "TGASS", "conf_prop": True, "preserve_input": True,
              "offsets": "smooth" if args.smooth_offsets else "iid N(0,2^2) px",
              "sharding": "batch shard per GPU, no collective on the data path",
              "l2": None}
    guid_mb = B * 3 * (K * K - 1) * H * W * 4 / 1e6
    flush = args.flush_l2 == "on" or (args.flush_l2 == "auto" and guid_mb <= 126.0)
    config["l2"] = ("a 256 MB buffer is written between timed steps (L2 flush; inputs are %.0f MB per GPU)" % guid_mb
                    if flush else
                    "inputs exceed L2 (guidance alone is %.0f MB per GPU vs 126 MB L2): no flush" % guid_mb)

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_img = args.cpu_images or min(os.cpu_count() or 1, 32)
        r = cpu_reference_run(args, H, W, md, n_img, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                 "sample": "%d frames x %d rows x %d cols per step (full-width strips of the %dx%d "
                                           "frame), one frame per host thread, torchvision deform_conv2d "
                                           "stand-in (oracle/torchvision_port.py), fwd%s, %d host cores"
                                           % (r["n_images"], r["rows"], W, H, W,
                                              "+bwd" if args.mode == "fwdbwd" else "", r["host_cores"])},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nlspn_eccv20_b200 import NLSPN, _lib
    from nlspn_eccv20_b200.synth import workload
    lib = _lib.load()

    host = workload(args.workload, B, K, seed=7240 + rank, smooth_offsets=args.smooth_offsets, pin=True)
    names = ["feat_init", "guidance", "confidence", "feat_fix"]
    dev_in = {k: host[k].to(dev) for k in names}
    gt = host["gt"].to(dev)
    mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
    train = args.mode == "fwdbwd"
    graphed = None
    if args.cuda_graph and not train:
        graphed = mod.graphed(dev_in["feat_init"], dev_in["guidance"], dev_in["confidence"], dev_in["feat_fix"])
        config["launch"] = "CUDA graph replay (GraphedNLSPN)"

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(inp, rec=None):
        fi, gd, cf = inp["feat_init"], inp["guidance"], inp["confidence"]
        if train:
            fi, gd, cf = (t.detach().requires_grad_(True) for t in (fi, gd, cf))
            mod.aff_scale_const.grad = None
        e0, e1, e2 = (ev(), ev(), ev()) if rec is not None else (None, None, None)
        if rec is not None:
            e0.record()
        with torch.set_grad_enabled(train):
            if graphed is not None and inp is dev_in:      # resident inputs ARE the graph's static buffers
                feat_result, list_feat, offset, aff, _ = graphed(*graphed.inputs)
            else:
                feat_result, list_feat, offset, aff, _ = (graphed or mod)(fi, gd, cf, inp["feat_fix"])
            pred = torch.clamp(feat_result, min=0)
            loss = (pred - gt).abs().sum()      # L1 surrogate of the reference loss (l1loss.py:27-42)
        if rec is not None:
            e1.record()
        if train:
            loss.backward()
        if rec is not None:
            e2.record()
            rec.append((e0, e1, e2))
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(dev_in)
    sync_all()

    # ---- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    rec = []
    n0 = lib.nlspn_launch_count()
    t_start, t_end = ev(), ev()
    flush_buf = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32) if flush else None
    t_start.record()
    for _ in range(args.steps):
        if flush:
            flush_buf.fill_(1.0)           # evicts L2; excluded from the step time below
        step(dev_in, rec)
    t_end.record()
    sync_all()
    launches = lib.nlspn_launch_count() - n0
    if graphed is not None:
        # replayed kernels do not pass through the library's launch counter: count one eager call
        n1 = lib.nlspn_launch_count()
        with torch.no_grad():
            mod(dev_in["feat_init"], dev_in["guidance"], dev_in["confidence"], dev_in["feat_fix"])
        torch.cuda.synchronize()
        launches = (lib.nlspn_launch_count() - n1) * args.steps
    clocks = sampler.stop()
    if flush:   # steps are timed individually (forward + backward events), the flush is not counted
        total_ms = sum(a.elapsed_time(c) for a, _, c in rec)
    else:
        total_ms = t_start.elapsed_time(t_end)
    fwd_ms = sum(a.elapsed_time(b) for a, b, _ in rec) / len(rec)
    bwd_ms = sum(b.elapsed_time(c) for _, b, c in rec) / len(rec)

    # ---- end to end: pinned host inputs -> H2D -> module -> D2H of the loss, every step.
    # The batch shard is fed in chunks of frames: a copy stream uploads chunk i+1 while the
    # compute stream runs the module on chunk i (images are independent, so the chunked step is
    # the same computation; this is the double-buffered prefetch a pinned DataLoader does).
    h2d = sum(host[k].numel() * 4 for k in names)
    # chunk of frames per upload: about one KITTI frame's worth of pixels (small frames are grouped so
    # the per-chunk launches stay above the launch-bound regime)
    auto_chunk = max(1, -(-400000 // (H * W)))
    chunk = max(1, min(B, int(os.environ.get("NLSPN_E2E_CHUNK", str(auto_chunk)))))
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream(dev)
    gt_chunks = [gt[i:i + chunk] for i in range(0, B, chunk)]

    def e2e_step():
        pending = None
        losses = []

        def upload(i):
            with torch.cuda.stream(copy_stream):
                inp = {k: host[k][i:i + chunk].to(dev, non_blocking=True) for k in names}
                evt = torch.cuda.Event()
                evt.record(copy_stream)
            return inp, evt

        pending = upload(0)
        for ci, i in enumerate(range(0, B, chunk)):
            inp, evt = pending
            pending = upload(i + chunk) if i + chunk < B else None
            main_stream.wait_event(evt)
            for t_ in inp.values():
                t_.record_stream(main_stream)
            fi, gd, cf = inp["feat_init"], inp["guidance"], inp["confidence"]
            if train:
                fi, gd, cf = (t_.requires_grad_(True) for t_ in (fi, gd, cf))
            with torch.set_grad_enabled(train):
                feat_result = mod(fi, gd, cf, inp["feat_fix"])[0]
                loss = (torch.clamp(feat_result, min=0) - gt_chunks[ci]).abs().sum()
            if train:
                loss.backward()
            losses.append(loss)
        return float(torch.stack(losses).sum().detach())      # D2H of the result + sync

    for _ in range(min(2, args.warmup)):
        mod.aff_scale_const.grad = None
        e2e_step()
    sync_all()
    e_start, e_end = ev(), ev()
    e_start.record()
    for _ in range(args.steps):
        mod.aff_scale_const.grad = None
        loss_host = e2e_step()
    e_end.record()
    sync_all()
    e2e_ms = e_start.elapsed_time(e_end)

    tms = torch.tensor([total_ms, e2e_ms, fwd_ms, bwd_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_ms, bwd_ms = (float(x) for x in tms.tolist())

    pix_iter_step = world * B * H * W * T
    value = pix_iter_step * args.steps / (total_ms * 1e-3) / 1e9
    e2e_val = pix_iter_step * args.steps / (e2e_ms * 1e-3) / 1e9
    peaks, peak_kind = measured_peaks()
    N = K * K - 1

    # ---- per-kernel split of the phases: a separate pass with the library's event hooks on
    # (events around every launch perturb throughput, so this pass is NOT the one `value` is from;
    # the kernel's average launch duration = phase time measured in the timed region x its share).
    lib.nlspn_profile_enable(1)
    graphed_keep, graphed = graphed, None          # the event hooks live in the eager launch path
    for _ in range(max(1, min(args.steps, 3))):
        step(dev_in)
    graphed = graphed_keep
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    lib.nlspn_profile_enable(0)
    # algorithmic bytes per LAUNCH and pixel (fp32, DESIGN.md "kernels"): what the kernel must move
    alg = {"prologue_fwd_kernel": 24 * N + 28,
           "iter_fwd_kernel": 12 * N + 20,
           "bwd_state_kernel": 12 * N + 40,
           "bwd_param_kernel": 8 * T + 24 * N + 8,
           "final_bwd_kernel": 8 * N + 4 * (N + 1) + 16 + 24,
           "iter_bwd_kernel": 36 * N + 36,
           # pass A in gather form (K >= 5): table entries 16 B x N + counter + one 16-byte block store;
           # the gy kernel then moves ~53 B/px (bwd_state_kernel's class); table build once per step
           "bwd_gather_kernel": 16 * N + 20,
           "table_build_kernel": 12 * N + 16 * N + 4}
    if "bwd_gather_kernel" in prof:
        alg["bwd_state_kernel"] = 53
    fwd_names = ("prologue_fwd_kernel", "iter_fwd_kernel")
    phase_prof = {"forward": sum(prof[k][0] for k in prof if k in fwd_names),
                  "backward": sum(prof[k][0] for k in prof if k not in fwd_names)}
    kernels = {}
    nprof = max(1, min(args.steps, 3))
    per_iter = ("iter_fwd_kernel", "bwd_state_kernel", "iter_bwd_kernel", "bwd_gather_kernel")
    for name, (ms, cnt) in prof.items():
        ph = "forward" if name in fwd_names else "backward"
        share = ms / phase_prof[ph] if phase_prof[ph] > 0 else 0.0
        phase_ms = fwd_ms if ph == "forward" else bwd_ms
        per_step = cnt / nprof
        step_ms = phase_ms * share
        bytes_step = alg.get(name, 0) * B * H * W * (T if name in per_iter else 1)
        gbs = bytes_step / (step_ms * 1e-3) / 1e9 if step_ms > 0 else 0.0
        kernels[name] = {"launches_per_step": per_step, "share_of_phase": share, "step_ms": step_ms,
                         "launch_ms": step_ms / per_step, "alg_bytes_per_launch": bytes_step / per_step,
                         "achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"]}
    kern = max(kernels, key=lambda k: kernels[k]["step_ms"])
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get("%s_%dx%d_B%d_per_gpu_K%d_T%d" % (args.workload, H, W, B, K, T), {}).get(kern)
    except Exception:
        traffic = None
    kb, kms, achieved = kernels[kern]["alg_bytes_per_launch"], kernels[kern]["launch_ms"], kernels[kern]["achieved_gbs"]
    step_gbs = alg_bytes(K, T, args.mode) * (pix_iter_step / world) * args.steps / (total_ms * 1e-3) / 1e9
    line = {"metric": METRIC if train else "nlspn_propagation_fwd_gpix_iter_per_s", "value": value,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
                    "chunk_frames": chunk},
            "gpu_launches": int(launches),
            "phases_ms": {"forward": fwd_ms, "backward": bwd_ms,
                          "forward_gpix_iter_per_s": pix_iter_step / world / (fwd_ms * 1e-3) / 1e9},
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peaks["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                         "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs)" if peak_kind == "measured" else "fallback 6.65 TB/s",
                         "alg_bytes_per_launch": kb,
                         "launch_ms": kms,
                         "note": "dominant kernel by time per step; launch duration = CUDA-event time of its "
                                 "phase in the timed region x the kernel's share of that phase (event-bracketed "
                                 "profile pass), / launches per step"},
            "kernels": kernels,
            "roofline_step": {"alg_bytes_per_pix_iter": alg_bytes(K, T, args.mode), "achieved": step_gbs,
                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": step_gbs / peaks["hbm_gbs"]}}
    # second roofline level: when the step's inputs are L2-resident the same algorithmic
    # bytes are also quoted against the measured L2 copy bandwidth (tools/l2_bench.cu -> profiles/l2_peak.json)
    try:
        with open(os.path.join(ROOT, "profiles", "l2_peak.json")) as f:
            l2 = json.load(f)
        if guid_mb <= 126.0:
            line["roofline_step_l2"] = {"achieved": step_gbs, "peak": l2["l2_copy_gbs"], "unit": "GB/s",
                                        "frac": step_gbs / l2["l2_copy_gbs"],
                                        "peak_source": "tools/l2_bench.cu, L2-resident copy (read+write), profiles/l2_peak.json"}
    except Exception:
        pass

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_img = args.cpu_images or min(os.cpu_count() or 1, 32)
        r = cpu_reference_run(args, H, W, md, n_img, 2, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": "%d frames x %d rows x %d cols (full-width strips of the %dx%d frame), one "
                                          "frame per host thread, fwd%s, torchvision deform_conv2d stand-in "
                                          "(oracle/torchvision_port.py), %d host cores, %.1f s per pass"
                                          % (r["n_images"], r["rows"], W, H, W, "+bwd" if train else "",
                                             r["host_cores"], r["seconds"])}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
