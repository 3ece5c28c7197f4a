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
e2e     same metric through the public module with PINNED HOST inputs: every step uploads its inputs
        (H2D) and reads back the loss AND the step's result feat_result (D2H into pinned memory) inside the
        timed region.  e2e_variants also reports the round-1 definition (loss only) and the strictest one
        (all input gradients copied back as well).
roofline  dominant kernel = the one with the largest share of the step (K=3: pass A of the backward,
        bwd_state_kernel, 12N+40 B/px per launch; K>=5: bwd_gather_kernel, 16N+20 B/px);
        achieved = its algorithmic bytes per launch / its average launch duration (CUDA-event time of
        its phase in the timed region x its share of that phase from an event-bracketed pass);
        peak = MEASURED_PEAKS.json hbm_gbs; traffic = measured DRAM bytes per launch (profiles/ncu_traffic.json).
        roofline_step = SURVEY 8d's algorithmic bytes of the whole step against the same peak;
        roofline_step_dram = the DRAM bytes the kernels really move (ncu) against the same peak.
ref_cuda  (N=1, when oracle/_ref/DCN_ref.so travelled with the repo) the reference's OWN CUDA kernels compiled
        for sm_100a, under the reference's per-iteration op chain, timed on the same GPU and inputs.
cpu_baseline  the reference path restated over torchvision.ops.deform_conv2d (north_star's CPU
        stand-in; oracle/torchvision_port.py, pinned against the unmodified reference), one full frame
        per host thread.  Rank 0, N=1 only.
other_configs  (N=1) BASELINE.json's configs 1, 2, 3 and 5 on this GPU, a few steps each.
--impl reference  times that CPU implementation alone (rank 0), same metric/config.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "nlspn_propagation_fwd_bwd_gpix_iter_per_s"
METRIC_FWD = "nlspn_propagation_fwd_gpix_iter_per_s"
UNIT = "Gpix*iter/s"
L2_MB = 126.0


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
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
    p.add_argument("--no-ref-cuda", action="store_true")
    p.add_argument("--no-other-configs", action="store_true")
    p.add_argument("--cpu-images", type=int, default=None, help="frames in the CPU sample (default: host cores, <= 32)")
    p.add_argument("--cpu-rows", type=int, default=0,
                   help="rows of each frame in the CPU sample (0 = full height; full width always)")
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


def l2_peak():
    try:
        with open(os.path.join(ROOT, "profiles", "l2_peak.json")) as f:
            return json.load(f)
    except Exception:
        return None


def alg_bytes(K, T, mode):
    """Algorithmic bytes per pixel*iteration (fork semantics, fp32)."""
    N = K * K - 1
    fwd = 12 * N + 20 + (24 * N + 28) / T
    bwd = 36 * N + 36 + (28 * N + 12) / T
    return fwd if mode == "fwd" else fwd + bwd


def bind_host(local, world):
    """Multi-rank runs: spread the ranks' host threads -- and with them the first-touch placement of their
    pinned staging buffers -- over the box's NUMA nodes, so that N ranks' H2D streams do not all pull from one
    node's DRAM.  Returns what was done (reported in the JSON line)."""
    info = {"numa_nodes": 1, "bound_to_node": None}
    try:
        nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"), key=lambda p: int(p.rsplit("node", 1)[1]))
        info["numa_nodes"] = max(1, len(nodes))
        if world > 1 and len(nodes) > 1:
            node = nodes[(local * len(nodes)) // world]
            cpus = set()
            for part in open(os.path.join(node, "cpulist")).read().strip().split(","):
                if "-" in part:
                    a, b = part.split("-")
                    cpus.update(range(int(a), int(b) + 1))
                elif part:
                    cpus.add(int(part))
            allowed = os.sched_getaffinity(0)
            cpus = cpus & allowed if cpus & allowed else cpus
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["bound_to_node"] = int(node.rsplit("node", 1)[1])
                info["cpus"] = len(cpus)
    except Exception as e:       # never fatal: this is placement, not correctness
        info["error"] = str(e)[:80]
    return info


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
    Sample: n_images full frames (--cpu-rows R > 0: full-width strips of R rows), one frame per host thread."""
    if args.cpu_rows and args.cpu_rows > 0:
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


def cpu_sample_text(r, H, W, train):
    shape = ("full %dx%d frames" % (H, W)) if r["rows"] == H else \
        ("%d-row x %d-col full-width strips of the %dx%d frame" % (r["rows"], W, H, W))
    return ("%d %s per step, one frame per host thread, fwd%s, torchvision deform_conv2d stand-in "
            "(oracle/torchvision_port.py), %d host cores, %.1f s per pass (mean; best pass = %.5f %s)"
            % (r["n_images"], shape, "+bwd" if train else "", r["host_cores"], r["seconds"], r["best"], UNIT))


class GpuRun:
    """One workload on one GPU: resident-input timing, optional per-kernel split, optional end-to-end legs."""

    def __init__(self, torch, dev, lib, workload_name, B, K, T, mode, seed, smooth=False, pin=False):
        from nlspn_eccv20_b200 import NLSPN
        from nlspn_eccv20_b200.synth import SHAPES, workload
        self.torch, self.dev, self.lib = torch, dev, lib
        self.H, self.W, self.md = SHAPES[workload_name]
        self.B, self.K, self.T, self.mode = B, K, T, mode
        self.train = mode == "fwdbwd"
        self.host = workload(workload_name, B, K, seed=seed, smooth_offsets=smooth, pin=pin)
        self.names = ["feat_init", "guidance", "confidence", "feat_fix"]
        self.dev_in = {k: self.host[k].to(dev) for k in self.names}
        self.gt = self.host["gt"].to(dev)
        self.mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
        self.graphed = None
        self.guid_mb = B * 3 * (K * K - 1) * self.H * self.W * 4 / 1e6

    def ev(self):
        return self.torch.cuda.Event(enable_timing=True)

    def use_graph(self):
        d = self.dev_in
        self.graphed = self.mod.graphed(d["feat_init"], d["guidance"], d["confidence"], d["feat_fix"])

    def step(self, inp, rec=None):
        torch, mod, train = self.torch, self.mod, self.train
        fi, gd, cf = inp["feat_init"], inp["guidance"], inp["confidence"]
        if train:
            fi, gd, cf = (t.detach().requires_grad_(True) for t in (fi, gd, cf))
            mod.aff_scale_const.grad = None
        e0, e1, e2 = (self.ev(), self.ev(), self.ev()) if rec is not None else (None, None, None)
        if rec is not None:
            e0.record()
        with torch.set_grad_enabled(train):
            if self.graphed is not None and inp is self.dev_in:      # resident inputs ARE the graph's static buffers
                feat_result = self.graphed(*self.graphed.inputs)[0]
            else:
                feat_result = (self.graphed or mod)(fi, gd, cf, inp["feat_fix"])[0]
            pred = torch.clamp(feat_result, min=0)
            loss = (pred - self.gt).abs().sum()      # L1 surrogate of the reference loss (l1loss.py:27-42)
        if rec is not None:
            e1.record()
        if train:
            loss.backward()
        if rec is not None:
            e2.record()
            rec.append((e0, e1, e2))
        return loss

    def timed(self, steps, warmup, flush, sync_all):
        """-> total_ms, fwd_ms, bwd_ms (per step), launches."""
        torch, lib = self.torch, self.lib
        for _ in range(warmup):
            self.step(self.dev_in)
        sync_all()
        rec = []
        n0 = lib.nlspn_launch_count()
        t_start, t_end = self.ev(), self.ev()
        flush_buf = torch.empty(64 * 1024 * 1024, device=self.dev, dtype=torch.float32) if flush else None
        t_start.record()
        for _ in range(steps):
            if flush:
                flush_buf.fill_(1.0)           # evicts L2; excluded from the step time below
            self.step(self.dev_in, rec)
        t_end.record()
        sync_all()
        launches = lib.nlspn_launch_count() - n0
        if self.graphed is not None:
            # replayed kernels do not pass through the library's launch counter: count one eager call
            n1 = lib.nlspn_launch_count()
            d = self.dev_in
            with torch.no_grad():
                self.mod(d["feat_init"], d["guidance"], d["confidence"], d["feat_fix"])
            torch.cuda.synchronize()
            launches = (lib.nlspn_launch_count() - n1) * steps
        if flush:   # steps are timed individually (forward + backward events), the flush is not counted
            total_ms = sum(a.elapsed_time(c) for a, _, c in rec)
        else:
            total_ms = t_start.elapsed_time(t_end)
        fwd_ms = sum(a.elapsed_time(b) for a, b, _ in rec) / len(rec)
        bwd_ms = sum(b.elapsed_time(c) for _, b, c in rec) / len(rec)
        return total_ms, fwd_ms, bwd_ms, int(launches)

    def kernel_split(self, nprof, fwd_ms, bwd_ms, peak_gbs):
        """Per-kernel split of the phases: a separate pass with the library's event hooks on (events around every
        launch perturb throughput, so this pass is NOT the one `value` is from; a kernel's average launch duration =
        phase time measured in the timed region x its share of that phase)."""
        from nlspn_eccv20_b200 import _lib
        lib, K, T, B, H, W = self.lib, self.K, self.T, self.B, self.H, self.W
        N = K * K - 1
        lib.nlspn_profile_enable(1)
        keep, self.graphed = self.graphed, None          # the event hooks live in the eager launch path
        for _ in range(nprof):
            self.step(self.dev_in)
        self.graphed = keep
        self.torch.cuda.synchronize()
        prof = _lib.profile_read()
        lib.nlspn_profile_enable(0)
        # algorithmic bytes per LAUNCH and pixel (fp32, DESIGN.md "kernels"): what the kernel must move
        alg = {"prologue_fwd_kernel": 24 * N + 28,
               "iter_fwd_kernel": 12 * N + 20,
               "bwd_state_kernel": 12 * N + 40,
               "bwd_param_kernel": 8 * T + 24 * N + 8,
               "final_bwd_kernel": 8 * N + 4 * (N + 1) + 16 + 24,
               "iter_bwd_kernel": 36 * N + 36,
               # pass A in gather form: table entries 16 B x N + counter + one 16-byte block store;
               # the gy kernel then moves ~53 B/px (bwd_state_kernel's class); table build once per step
               "bwd_gather_kernel": 16 * N + 20,
               "table_build_kernel": 12 * N + 16 * N + 4}
        if "bwd_gather_kernel" in prof:
            alg["bwd_state_kernel"] = 53
        elif K == 3:
            # pass A as the tile-local transpose (kernels_local.cuh): the once-per-step schedule build reads offsets,
            # affinities, confidence, fixed depth and writes the packed record (128 B) + block list (9.3 B) per pixel
            alg["table_build_kernel"] = 12 * N + 12 + 128 + 10
        fwd_names = ("prologue_fwd_kernel", "iter_fwd_kernel")
        phase_prof = {"forward": sum(prof[k][0] for k in prof if k in fwd_names),
                      "backward": sum(prof[k][0] for k in prof if k not in fwd_names)}
        kernels = {}
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
                             "launch_ms": step_ms / per_step if per_step else 0.0,
                             "alg_bytes_per_launch": bytes_step / per_step if per_step else 0.0,
                             "achieved_gbs": gbs, "frac": gbs / peak_gbs}
        return kernels

    # ---- end to end: pinned host inputs -> H2D -> module -> D2H, every step.
    # The batch shard is fed in chunks of frames: a copy stream uploads chunk i+1 while the compute stream runs
    # the module on chunk i, and a third stream drains chunk i's results to pinned host memory (images are
    # independent, so the chunked step is the same computation; this is the double-buffered prefetch a pinned
    # DataLoader does).  d2h: "loss" = the scalar only; "result" = + feat_result; "grads" = + every input gradient.
    def e2e(self, steps, warmup, sync_all, d2h="result"):
        torch, dev, mod, train = self.torch, self.dev, self.mod, self.train
        host, names, B, H, W = self.host, self.names, self.B, self.H, self.W
        auto_chunk = max(1, -(-800000 // (H * W)))   # about two KITTI frames' worth of pixels per upload (B200: chunk 1 / 2 / 4 / 8 -> 8.45 / 8.77 / 8.81 / 8.5 Gpix*iter/s; 2 is also the steadiest)
        chunk = max(1, min(B, int(os.environ.get("NLSPN_E2E_CHUNK", str(auto_chunk)))))
        copy_stream = torch.cuda.Stream(device=dev)
        down_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        gt_chunks = [self.gt[i:i + chunk] for i in range(0, B, chunk)]
        out_host = {}
        if d2h in ("result", "grads"):
            out_host["feat_result"] = torch.empty(B, 1, H, W).pin_memory()
        if d2h == "grads" and train:
            for k in ("feat_init", "guidance", "confidence"):
                out_host["g_" + k] = torch.empty_like(host[k]).pin_memory()
        d2h_bytes = 4 + sum(t.numel() * 4 for t in out_host.values())

        # device-side input ring: three preallocated chunk buffers per tensor, refilled by the copy stream once the
        # compute stream is done with them (no caching-allocator traffic, no record_stream bookkeeping on the hot
        # path: with freshly allocated chunk tensors one of the variants would run 20-25 % slow every few runs)
        ring = 3
        slots = [{k: torch.empty((chunk,) + tuple(host[k].shape[1:]), device=dev) for k in names} for _ in range(ring)]
        slot_free = [None] * ring
        n_up = [0]

        def upload(i):
            slot = n_up[0] % ring
            n_up[0] += 1
            n = min(chunk, B - i)
            with torch.cuda.stream(copy_stream):
                if slot_free[slot] is not None:
                    copy_stream.wait_event(slot_free[slot])
                inp = {}
                for k in names:
                    dst = slots[slot][k][:n]
                    dst.copy_(host[k][i:i + n], non_blocking=True)
                    inp[k] = dst
                evt = torch.cuda.Event()
                evt.record(copy_stream)
            return inp, evt, slot

        prefetch = os.environ.get("NLSPN_E2E_PREFETCH", "1") != "0"
        nxt = [None]     # chunk 0 of the NEXT step, uploaded while this step's last chunk computes (a pinned
                         # DataLoader's prefetch: the copy engine never idles between steps)

        def e2e_step(more):
            losses = []
            pending = nxt[0] if nxt[0] is not None else upload(0)
            nxt[0] = None
            for ci, i in enumerate(range(0, B, chunk)):
                inp, evt, slot = pending
                if i + chunk < B:
                    pending = upload(i + chunk)
                else:
                    pending = None
                    if more and prefetch:
                        nxt[0] = upload(0)
                main_stream.wait_event(evt)
                fi, gd, cf = inp["feat_init"], inp["guidance"], inp["confidence"]
                if train:
                    fi, gd, cf = (t_.detach().requires_grad_(True) for t_ in (fi, gd, cf))
                with torch.set_grad_enabled(train):
                    feat_result = mod(fi, gd, cf, inp["feat_fix"])[0]
                    loss = (torch.clamp(feat_result, min=0) - gt_chunks[ci]).abs().sum()
                if train:
                    loss.backward()
                losses.append(loss.detach())
                freed = torch.cuda.Event()
                freed.record(main_stream)          # the chunk's inputs are consumed: its slot may be refilled
                slot_free[slot] = freed
                if out_host:
                    done = torch.cuda.Event()
                    done.record(main_stream)
                    down_stream.wait_event(done)
                    outs = {"feat_result": feat_result.detach()}
                    if train:
                        outs.update({"g_feat_init": fi.grad, "g_guidance": gd.grad, "g_confidence": cf.grad})
                    with torch.cuda.stream(down_stream):
                        for k, buf in out_host.items():
                            outs[k].record_stream(down_stream)
                            buf[i:i + chunk].copy_(outs[k], non_blocking=True)
            total = float(torch.stack(losses).sum())      # D2H of the loss + sync of the compute stream
            down_stream.synchronize()                      # results are in host memory when the step ends
            return total

        for _ in range(min(3, warmup)):
            mod.aff_scale_const.grad = None
            e2e_step(False)
        sync_all()
        e_start, e_end = self.ev(), self.ev()
        e_start.record()
        for s_ in range(steps):
            mod.aff_scale_const.grad = None
            e2e_step(s_ + 1 < steps)     # every step's upload lies inside the timed region
        e_end.record()
        sync_all()
        h2d = sum(host[k].numel() * 4 for k in names)
        return e_start.elapsed_time(e_end), h2d, d2h_bytes, chunk


def ref_cuda_leg(torch, run, steps=3, warmup=1):
    """The reference's own CUDA kernels (oracle/_ref/DCN_ref.so) on this GPU, same inputs, same loss."""
    from oracle import ref_cuda
    if not ref_cuda.available():
        return None
    ref_cuda.load()
    d, K, T, train = run.dev_in, run.K, run.T, run.train
    gam = run.mod.aff_scale_const.detach().clone().requires_grad_(train)

    def theirs():
        fi, gd, cf = (d[k].detach().requires_grad_(train) for k in ("feat_init", "guidance", "confidence"))
        gam.grad = None
        with torch.set_grad_enabled(train):
            out = ref_cuda.propagate(fi, gd, cf, d["feat_fix"], gam, K, T)["feat_result"]
            loss = (out.clamp(min=0) - run.gt).abs().sum()
        if train:
            loss.backward()

    for _ in range(warmup):
        theirs()
    torch.cuda.synchronize()
    e0, e1 = run.ev(), run.ev()
    e0.record()
    for _ in range(steps):
        theirs()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "value": run.B * run.H * run.W * T / (ms * 1e-3) / 1e9, "unit": UNIT,
            "steps": steps, "warmup": warmup,
            "what": "the reference's modulated_deform_conv_cuda.cu/.cuh compiled unmodified for sm_100a "
                    "(oracle/build_ref_cuda.py -> oracle/_ref/DCN_ref.so) under the reference's per-iteration torch op "
                    "chain (oracle/ref_cuda.py), same GPU, same resident inputs, same loss"}


def dcn_step_leg(torch, run, steps=10, warmup=3):
    """Boundary B1 (the reference's own operator signature): ONE ModulatedDeformConvFunction forward + backward at
    the headline shape through nlspn_eccv20_b200.dcn, and -- when oracle/_ref/DCN_ref.so travelled -- the reference's
    kernels under the same autograd Function shape (oracle/ref_cuda.RefDeformStep)."""
    from nlspn_eccv20_b200 import dcn
    B, H, W, K = run.B, run.H, run.W, run.K
    g = torch.Generator(device=run.dev).manual_seed(7240)
    x = torch.rand(B, 1, H, W, device=run.dev, generator=g)
    off = 2.0 * torch.randn(B, 2 * K * K, H, W, device=run.dev, generator=g)
    msk = torch.rand(B, K * K, H, W, device=run.dev, generator=g)
    w = torch.ones(1, 1, K, K, device=run.dev)
    bias = torch.zeros(1, device=run.dev)
    pad = (K - 1) // 2

    def ours():
        a, o, m = (t.detach().requires_grad_(True) for t in (x, off, msk))
        y = dcn.ModulatedDeformConvFunction.apply(a, o, m, w, bias, 1, pad, 1, 1, 1, 64)
        y.sum().backward()

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = run.ev(), run.ev()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms = timed(ours)
    out = {"workload": "one DCN step fwd+bwd, %dx%d B=%d K=%d (offset, mask and input gradients)" % (H, W, B, K),
           "ms_per_step": ms, "value": B * H * W / (ms * 1e-3) / 1e9, "unit": UNIT, "steps": steps, "warmup": warmup}
    try:
        from oracle import ref_cuda
        if ref_cuda.available():
            ref_cuda.load()

            def theirs():
                a, o, m = (t.detach().requires_grad_(True) for t in (x, off, msk))
                y = ref_cuda.RefDeformStep.apply(a, o, m, w, bias, K)
                y.sum().backward()

            ms_ref = timed(theirs)
            out["ref_cuda_ms_per_step"] = ms_ref
            out["ours_over_reference_cuda"] = ms_ref / ms
    except Exception as e:     # the referee is optional
        out["ref_cuda_error"] = str(e)[:200]
    torch.cuda.empty_cache()
    return out


def heads_leg(torch, dev, B, H, W, K=3, steps=10, warmup=3):
    """the three final head convolutions (nlspnmodel.py:297,301,313) at the headline shape -- stock
    torch (three torch.cat + three cuDNN TF32 convolutions, NCHW) vs ours (one tcgen05 implicit GEMM, heads.py)."""
    from nlspn_eccv20_b200 import heads
    N3 = 3 * (K * K - 1)
    g = torch.Generator(device=dev).manual_seed(7240)
    x = [torch.randn(B, 64, H, W, device=dev, generator=g) for _ in range(4)]
    s = (128 * 9) ** -0.5
    w = [s * torch.randn(n, 128, 3, 3, device=dev, generator=g) for n in (1, N3, 1)]
    b = [0.1 * torch.randn(n, device=dev, generator=g) for n in (1, N3, 1)]
    args = (x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])

    def timed(fn):
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

    from nlspn_eccv20_b200 import _lib as L_, functional as F_
    from nlspn_eccv20_b200.synth import make_inputs
    fix = make_inputs(B, H, W, K, seed=11, density=0.05, device="cpu")["feat_fix"].to(dev)
    gamma = torch.tensor([0.5 * (K * K - 1)], device=dev)
    with torch.no_grad():
        ours = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
        with L_.options(heads_rows=0):
            ninetap = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
        stock = timed(lambda: heads.reference_heads(*args))
        o, r = heads.fused_heads(*args, prop_kernel=K), heads.reference_heads(*args)
        diff = max(float((a - c).abs().max()) for a, c in zip(o, r))
        prologue = timed(lambda: F_.prologue_fwd(o[1], o[2], o[0], fix, gamma, K))
        fused = timed(lambda: heads.fused_heads_prologue(*args, fix, gamma, K)) if heads.prologue_supported(W, K) else None
    flops = 2.0 * B * H * W * (128 * 9) * (N3 + 2)
    KK = K * K
    # algorithmic bytes: the four 64-channel inputs once; heads write 3N + 2 planes; fused: 3 KK + 4 planes + feat_fix read
    bytes_heads = 4.0 * B * H * W * (256 + N3 + 2)
    bytes_fused = 4.0 * B * H * W * (256 + 1 + 3 * KK + 4)
    peak = float(measured_peaks()[0].get("hbm_gbs", 0.0))
    out = {"workload": "head convolutions 128 -> 1 / %d / 1, 3x3, %dx%d B=%d, forward" % (N3, H, W, B),
           "ours_ms": ours, "ours_ninetap_form_ms": ninetap, "stock_torch_ms": stock, "ours_over_stock": stock / ours,
           "dtype": "tf32 (fp32 accumulate)", "useful_tflops": flops / (ours * 1e-3) / 1e12,
           "hbm_gbs": bytes_heads / (ours * 1e-3) / 1e9, "alg_bytes": bytes_heads,
           "max_abs_diff_vs_cudnn_tf32": diff, "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
           "prologue_ms": prologue, "heads_plus_prologue_ms": ours + prologue,
           "stock_heads_plus_prologue_ms": stock + prologue,
           "fused_heads_prologue_ms": fused,
           "fused_hbm_gbs": (bytes_fused / (fused * 1e-3) / 1e9) if fused else None}
    if peak > 0:
        out["hbm_frac"] = out["hbm_gbs"] / peak
        if fused:
            out["fused_hbm_frac"] = out["fused_hbm_gbs"] / peak
    # training: forward + backward of the same layers (every input, weight and bias gradient), stock autograd / cuDNN vs
    # heads.fused_heads (GEMM forward; weight + bias gradients on tcgen05, one-channel data gradients as a stencil,
    # the two wide data gradients in cuDNN); and the weight-gradient kernels alone against the bytes they must read
    leaves = [t.detach().clone().requires_grad_(True) for t in args]
    gout = [torch.randn(B, n, H, W, device=dev, generator=g) for n in (1, N3, 1)]

    def train(fn):
        for t in leaves:
            t.grad = None
        torch.autograd.backward(fn(), gout)
    train_ours = timed(lambda: train(lambda: heads.fused_heads(*leaves, prop_kernel=K)))
    train_stock = timed(lambda: train(lambda: heads.reference_heads(*leaves)))
    out["train"] = {"workload": "forward + backward (all gradients) of the same layers", "ours_ms": train_ours,
                    "stock_torch_ms": train_stock, "ours_over_stock": train_stock / train_ours}
    if heads.wgrad_supported(W, K):
        with torch.no_grad():
            gs, _ = heads.grad_prep(o[0], o[2], gout[0], gout[1], gout[2], K)
            prep = timed(lambda: heads.grad_prep(o[0], o[2], gout[0], gout[1], gout[2], K))
            wgrad = timed(lambda: heads.weight_grads(x[0], x[1], x[2], x[3], gs, K))
        # what nlspn_heads_wgrad must read: the four inputs once, the three gradient copies once for the wide launches and
        # eight channels of them for the one-channel heads (K = 3: one block of gradient channels)
        blocks = (N3 + 2 + 31) // 32
        bytes_wgrad = 4.0 * B * H * W * (blocks * 128 + 128 + 3 * (N3 + 2) + 3 * 8)
        out["train"].update({"grad_prep_ms": prep, "wgrad_ms": wgrad, "wgrad_alg_bytes": bytes_wgrad,
                             "wgrad_hbm_gbs": bytes_wgrad / (wgrad * 1e-3) / 1e9})
        if peak > 0:
            out["train"]["wgrad_hbm_frac"] = out["train"]["wgrad_hbm_gbs"] / peak
        del gs
    del x, w, b, args, o, r, leaves, gout
    torch.cuda.empty_cache()
    return out


def other_configs(torch, dev, lib, peaks):
    """BASELINE.json configs 1, 2, 3, 5 on this GPU (config 4 is the full model: tests/perf_config4_train_step.py)."""
    out = {}
    l2 = l2_peak()
    specs = [("config1_nyu_B1_fwd_eager", "nyu", 1, 3, 18, "fwd", False),
             ("config1_nyu_B1_fwd_cuda_graph", "nyu", 1, 3, 18, "fwd", True),
             ("config2_nyu_B12_fwdbwd", "nyu", 12, 3, 18, "fwdbwd", False),
             ("config3_kitti_B16_fwd", "kitti", 16, 3, 18, "fwd", False),
             ("config5_kitti_B8_K5_T36_fwdbwd", "kitti", 8, 5, 36, "fwdbwd", False)]
    for name, wl, B, K, T, mode, graph in specs:
        try:
            run = GpuRun(torch, dev, lib, wl, B, K, T, mode, seed=7240)
            if graph:
                run.use_graph()
            flush = run.guid_mb <= L2_MB
            steps = 20 if wl == "nyu" else 5
            total_ms, fwd_ms, bwd_ms, launches = run.timed(steps, 3, flush, torch.cuda.synchronize)
            pix = B * run.H * run.W * T
            val = pix * steps / (total_ms * 1e-3) / 1e9
            gbs = alg_bytes(K, T, mode) * val
            r = {"value": val, "unit": UNIT, "ms_per_step": total_ms / steps, "steps": steps, "warmup": 3,
                 "mode": mode, "gpu_launches_per_step": launches / steps, "l2_flush_between_steps": flush,
                 "roofline_step": {"achieved": gbs, "peak": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "unit": "GB/s"}}
            if l2 and flush:
                r["roofline_step_l2"] = {"achieved": gbs, "peak": l2["l2_copy_gbs"], "frac": gbs / l2["l2_copy_gbs"],
                                         "unit": "GB/s"}
            if graph:
                r["launch"] = "CUDA graph replay (GraphedNLSPN)"
            out[name] = r
            del run
            torch.cuda.empty_cache()
        except Exception as e:     # an extra config must never take the headline down
            out[name] = {"error": str(e)[:200]}
    return out


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
    train = args.mode == "fwdbwd"
    config = {"workload": "%s_%dx%d_B%d_per_gpu_K%d_T%d_%s" % (args.workload, H, W, B, K, T, args.mode),
              "frames_per_gpu": B, "height": H, "width": W, "prop_kernel": K, "prop_time": T,
              "affinity": "TGASS", "conf_prop": True, "preserve_input": True,
              "offsets": "smooth" if args.smooth_offsets else "iid N(0,2^2) px",
              "sharding": "batch shard per GPU, no collective on the data path",
              "l2": None}
    guid_mb = B * 3 * (K * K - 1) * H * W * 4 / 1e6
    flush = args.flush_l2 == "on" or (args.flush_l2 == "auto" and guid_mb <= L2_MB)
    config["l2"] = ("a 256 MB buffer is written between timed steps (L2 flush; inputs are %.0f MB per GPU)" % guid_mb
                    if flush else
                    "inputs exceed L2 (guidance alone is %.0f MB per GPU vs 126 MB L2): no flush" % guid_mb)

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        n_img = args.cpu_images or min(os.cpu_count() or 1, 32)
        r = cpu_reference_run(args, H, W, md, n_img, max(1, args.steps), max(0, args.warmup))
        config["cpu_sample"] = cpu_sample_text(r, H, W, train)
        line = {"impl": "reference", "metric": METRIC if train else METRIC_FWD, "value": r["value"], "unit": UNIT,
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": r["value"], "best": r["best"], "unit": UNIT, "cores": r["cores"],
                                 "kind": "port", "sample": config["cpu_sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- our arm (GPU)
    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU"
    host_info = bind_host(local, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from nlspn_eccv20_b200 import _lib
    lib = _lib.load()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    run = GpuRun(torch, dev, lib, args.workload, B, K, T, args.mode, seed=7240 + rank,
                 smooth=args.smooth_offsets, pin=True)
    if args.cuda_graph and not train:
        run.use_graph()
        config["launch"] = "CUDA graph replay (GraphedNLSPN)"

    # ---- device-resident timing
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, fwd_ms, bwd_ms, launches = run.timed(args.steps, args.warmup, flush, sync_all)
    clocks = sampler.stop()

    # ---- end to end (three D2H definitions; the middle one is the headline `e2e`)
    e2e_ms, h2d, d2h_bytes, chunk = run.e2e(args.steps, args.warmup, sync_all, "result")
    e2e_loss_ms, _, d2h_loss, _ = run.e2e(args.steps, args.warmup, sync_all, "loss")
    if train:
        e2e_grads_ms, _, d2h_grads, _ = run.e2e(args.steps, args.warmup, sync_all, "grads")
    else:
        e2e_grads_ms, d2h_grads = e2e_ms, d2h_bytes

    tms = torch.tensor([total_ms, e2e_ms, fwd_ms, bwd_ms, e2e_loss_ms, e2e_grads_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_ms, bwd_ms, e2e_loss_ms, e2e_grads_ms = (float(x) for x in tms.tolist())

    pix_iter_step = world * B * H * W * T
    rate = lambda ms: pix_iter_step * args.steps / (ms * 1e-3) / 1e9
    value = rate(total_ms)
    peaks, peak_kind = measured_peaks()

    kernels = run.kernel_split(max(1, min(args.steps, 3)), fwd_ms, bwd_ms, peaks["hbm_gbs"])
    kern = max(kernels, key=lambda k: kernels[k]["step_ms"])
    traffic = None
    traffic_all = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic_all = json.load(f).get("%s_%dx%d_B%d_per_gpu_K%d_T%d" % (args.workload, H, W, B, K, T), {})
        traffic = traffic_all.get(kern)
    except Exception:
        traffic = None
    kb, kms, achieved = kernels[kern]["alg_bytes_per_launch"], kernels[kern]["launch_ms"], kernels[kern]["achieved_gbs"]
    step_ms = total_ms / args.steps
    step_gbs = alg_bytes(K, T, args.mode) * (pix_iter_step / world) / (step_ms * 1e-3) / 1e9
    line = {"metric": METRIC if train else METRIC_FWD, "value": value,
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "clocks": clocks,
            "e2e": {"value": rate(e2e_ms), "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h_bytes * world, "chunk_frames": chunk,
                    "d2h": "loss + feat_result into pinned host memory"},
            "e2e_variants": {
                "loss_only_d2h": {"value": rate(e2e_loss_ms), "d2h_bytes_per_step": d2h_loss * world},
                "result_d2h": {"value": rate(e2e_ms), "d2h_bytes_per_step": d2h_bytes * world},
                "result_and_all_input_gradients_d2h": {"value": rate(e2e_grads_ms),
                                                       "d2h_bytes_per_step": d2h_grads * world}},
            "host": host_info,
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
    # measured-traffic utilisation: what the kernels really move through DRAM (ncu, per launch) / step time
    if traffic_all and all(k in traffic_all for k in kernels):
        dram = sum(traffic_all[k] * kernels[k]["launches_per_step"] for k in kernels)
        dram_gbs = dram / (step_ms * 1e-3) / 1e9
        line["roofline_step_dram"] = {"dram_bytes_per_step": dram, "achieved": dram_gbs, "peak": peaks["hbm_gbs"],
                                      "unit": "GB/s", "frac": dram_gbs / peaks["hbm_gbs"],
                                      "source": "profiles/ncu_traffic.json (dram__bytes_read+write per launch, ncu --set full) "
                                                "x launches per step / measured step time"}
    # second roofline level: when the step's inputs are L2-resident the same algorithmic
    # bytes are also quoted against the measured L2 copy bandwidth (tools/l2_bench.cu -> profiles/l2_peak.json)
    l2 = l2_peak()
    if l2 and guid_mb <= L2_MB:
        line["roofline_step_l2"] = {"achieved": step_gbs, "peak": l2["l2_copy_gbs"], "unit": "GB/s",
                                    "frac": step_gbs / l2["l2_copy_gbs"],
                                    "peak_source": "tools/l2_bench.cu, L2-resident copy (read+write), profiles/l2_peak.json"}

    if rank == 0 and world == 1:
        if not args.no_ref_cuda:
            try:
                rc = ref_cuda_leg(torch, run)
                if rc is not None:
                    rc["ours_over_reference_cuda"] = rc["ms_per_step"] / step_ms
                    line["ref_cuda"] = rc
            except Exception as e:
                line["ref_cuda"] = {"error": str(e)[:200]}
            try:
                line["dcn_step"] = dcn_step_leg(torch, run)
            except Exception as e:
                line["dcn_step"] = {"error": str(e)[:200]}
            try:
                line["heads"] = heads_leg(torch, dev, run.B, run.H, run.W, run.K)
            except Exception as e:
                line["heads"] = {"error": str(e)[:200]}
        del run
        torch.cuda.empty_cache()
        if not args.no_other_configs:
            line["other_configs"] = other_configs(torch, dev, lib, peaks)
        if not args.no_cpu_baseline:
            n_img = args.cpu_images or min(os.cpu_count() or 1, 32)
            r = cpu_reference_run(args, H, W, md, n_img, 2, 1)
            config["cpu_sample"] = cpu_sample_text(r, H, W, train)
            line["cpu_baseline"] = {"value": r["value"], "best": r["best"], "unit": UNIT, "cores": r["cores"],
                                    "kind": "port", "sample": config["cpu_sample"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
