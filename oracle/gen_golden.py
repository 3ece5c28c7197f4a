"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python oracle/gen_golden.py
The fixtures are committed; the GPU box never runs this script.

Each fixture holds the five inputs of the propagation path and everything the
reference produced from them on CPU (fp32) with torchvision's deform_conv2d as the
DCN stand-in (oracle/ref_harness.py), plus gradients from the reference's own
autograd for a fixed loss.  Cases:

  fullmodel_*   a real ``NLSPNModel.forward`` (ResNet34 encoder-decoder, random init);
                the five path inputs are captured with forward hooks on the heads
                (nlspnmodel.py:297,301,313) and the outputs are the forward's own dict
                (:379-381).  This pins the ORDER of the statements at :323-377.
  path_*        the same statements driven through the reference's helper methods
                (ref_harness.reference_propagate) on designed inputs (stable set,
                signed set, edge set, K=5, no-confidence, other affinity modes).
  dcn_*         single ModulatedDeformConvFunction.apply calls (boundary B1).
  gru_*         a real ``NLSPNModel(use_GRU=True, use_S2D=True)`` forward + backward (the fork's default
                configuration; affinities re-estimated between iterations), with and without offsets.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as RH  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 7240  # reference default seed (src/config.py:58-61)


def smooth_field(g, B, H, W, lo, hi, cell=8):
    gh, gw = max(2, H // cell + 1), max(2, W // cell + 1)
    grid = lo + (hi - lo) * torch.rand(B, 1, gh, gw, generator=g)
    f = torch.nn.functional.interpolate(grid, size=(H, W), mode="bicubic", align_corners=True)
    return f.clamp(lo, hi)


def make_inputs(g, B, H, W, K, max_depth=10.0, signed=False, conf_mean=3.0, density=0.05,
                off_sigma=2.0):
    N = K * K - 1
    gt = smooth_field(g, B, H, W, 0.5, max_depth)
    if signed:
        feat_init = max_depth * torch.rand(B, 1, H, W, generator=g)
        aff_raw = torch.randn(B, N, H, W, generator=g)
    else:
        feat_init = (gt + 0.005 * max_depth * torch.randn(B, 1, H, W, generator=g)).clamp(min=0)
        aff_raw = torch.randn(B, N, H, W, generator=g).abs()
    off_raw = off_sigma * torch.randn(B, 2 * N, H, W, generator=g)
    guidance = torch.cat([off_raw, aff_raw], 1)
    confidence = torch.sigmoid(conf_mean + torch.randn(B, 1, H, W, generator=g))
    mask = (torch.rand(B, 1, H, W, generator=g) < density).float()
    feat_fix = gt * mask
    return dict(feat_init=feat_init, guidance=guidance, confidence=confidence,
                feat_fix=feat_fix, gt=gt)


def loss_fn(out, gt):
    """Fixed differentiable loss that touches the final AND intermediate states."""
    pred = out["pred"]
    l = ((pred - gt).abs().mean() + ((pred - gt) ** 2).mean())
    for i, p in enumerate(out["pred_inter"][:-1]):
        l = l + 0.05 * ((i % 3) + 1) * (p * torch.cos(gt * (i + 1))).mean()
    return l


def run_path(model, inp, with_grad=True, use_conf=True):
    N = model.num_neighbors
    fi = inp["feat_init"].clone().requires_grad_(with_grad)
    if not model.args.offset:
        inp["guidance"] = inp["guidance"][:, 2 * N:].contiguous()   # raw affinities only (:306-308)
    gd = inp["guidance"].clone().requires_grad_(with_grad)
    cf = inp["confidence"].clone().requires_grad_(with_grad) if use_conf else None
    if with_grad and model.aff_scale_const.requires_grad:
        model.aff_scale_const.grad = None
    if model.args.offset:
        out = RH.reference_propagate(model, fi, gd[:, :2 * N], gd[:, 2 * N:], cf, inp["feat_fix"])
    else:
        out = RH.reference_propagate(model, fi, None, gd, cf, inp["feat_fix"])
    rec = dict(
        feat_result=out["feat_result"].detach(), pred=out["pred"].detach(),
        list_feat=torch.stack([p.detach() for p in out["pred_inter"]], 0),
        aff=out["aff"].detach(), gamma=out["gamma"].clone(),
    )
    if out["offset"] is not None:
        rec["offset"] = out["offset"].detach()
    if use_conf:
        rec["conf_fixed"] = out["confidence"].detach()
    if with_grad:
        # DIRECT upstream gradient of every list entry: evaluate the loss on detached leaf
        # copies, then push those gradients through the reference's graph.  (autograd.grad
        # wrt the live intermediates would return TOTAL derivatives.)  'pred' is
        # clamp(list[-1], 0) (nlspnmodel.py:375-377), so its gradient is part of g_list[-1].
        leaves = [p.detach().clone().requires_grad_(True) for p in out["pred_inter"]]
        loss = loss_fn(dict(pred=torch.clamp(leaves[-1], min=0), pred_inter=leaves), inp["gt"])
        gl = torch.autograd.grad(loss, leaves, allow_unused=True)
        gl = [torch.zeros_like(fi) if x is None else x for x in gl]
        rec["g_list"] = torch.stack(gl, 0)
        torch.autograd.backward(out["pred_inter"], gl)
        rec["loss"] = loss.detach()
        rec["g_feat_init"] = fi.grad.clone()
        rec["g_guidance"] = gd.grad.clone()
        if use_conf:
            rec["g_confidence"] = cf.grad.clone()
        if model.aff_scale_const.grad is not None:
            rec["g_gamma"] = model.aff_scale_const.grad.clone()
    return rec


def save(name, inp, rec, meta):
    d = {}
    for k, v in inp.items():
        if v is not None:
            d["in_" + k] = v.detach().numpy()
    for k, v in rec.items():
        d["out_" + k] = v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)
    for k, v in meta.items():
        d["meta_" + k] = np.asarray(v)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %-40s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def shell_model(**kw):
    """A reference NLSPNModel whose encoder is never run (resnet18 keeps it cheap)."""
    torch.manual_seed(SEED)
    return RH.build_reference_model(network="resnet18", **kw)


def case_fullmodel():
    """Real NLSPNModel.forward; inputs captured by hooks; checked against the harness."""
    torch.manual_seed(SEED)
    B, H, W, K, T = 1, 24, 32, 3, 6
    model = RH.build_reference_model(network="resnet34", prop_kernel=K, prop_time=T)
    model.eval()
    # the heads are freshly initialised, so guidance is tiny; scale it up so offsets matter
    with torch.no_grad():
        model.off_aff_dec0[0].weight.mul_(40.0)
    cap = {}
    hooks = [
        model.id_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("feat_init", o.detach().clone())),
        model.off_aff_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("guidance", o.detach().clone())),
        model.cf_dec0.register_forward_hook(lambda m, i, o: cap.__setitem__("confidence", o.detach().clone())),
    ]
    g = torch.Generator().manual_seed(SEED)
    gt = smooth_field(g, B, H, W, 0.5, 10.0)
    dep = gt * (torch.rand(B, 1, H, W, generator=g) < 0.1).float()
    rgb = torch.randn(B, 3, H, W, generator=g)
    with torch.no_grad():
        out = model({"rgb": rgb, "dep": dep})
    for h in hooks:
        h.remove()
    inp = dict(feat_init=cap["feat_init"], guidance=cap["guidance"],
               confidence=cap["confidence"], feat_fix=dep, gt=gt)
    rec = dict(pred=out["pred"], list_feat=torch.stack(out["pred_inter"], 0),
               feat_result=out["pred_inter"][-1], offset=out["offset"], aff=out["aff"],
               gamma=out["gamma"].clone(), conf_fixed=out["confidence"])
    # the harness (helper methods in the forward's order) must reproduce forward bit-exactly
    N = K * K - 1
    with torch.no_grad():
        h = RH.reference_propagate(model, inp["feat_init"], inp["guidance"][:, :2 * N],
                                   inp["guidance"][:, 2 * N:], inp["confidence"], dep)
    assert torch.equal(h["pred"], out["pred"]), "harness != reference forward"
    assert torch.equal(h["aff"], out["aff"]) and torch.equal(h["offset"], out["offset"])
    assert all(torch.equal(a, b) for a, b in zip(h["pred_inter"], out["pred_inter"]))
    print("harness == NLSPNModel.forward bit-exactly (pred, pred_inter, offset, aff)")
    save("fullmodel_k3_t6", inp, rec,
         dict(K=K, T=T, affinity="TGASS", preserve=1, use_conf=1, gamma=float(out["gamma"])))


def case_paths():
    specs = [
        # name, B, H, W, K, T, kwargs for inputs, model kwargs, use_conf
        ("path_stable_k3_t18", 2, 20, 28, 3, 18, dict(), dict(), True),
        ("path_stable_k5_t8", 1, 18, 22, 5, 8, dict(), dict(), True),
        ("path_signed_k3_t3", 2, 16, 20, 3, 3, dict(signed=True), dict(), True),
        ("path_noconf_k3_t6", 1, 16, 24, 3, 6, dict(), dict(conf_prop=False), False),
        ("path_nopreserve_k3_t5", 1, 14, 18, 3, 5, dict(), dict(preserve_input=False), True),
        ("path_faroff_k3_t4", 1, 12, 16, 3, 4, dict(off_sigma=12.0), dict(), True),
        ("path_AS_k3_t4", 1, 12, 16, 3, 4, dict(), dict(affinity="AS"), True),
        ("path_ASS_k3_t4", 1, 12, 16, 3, 4, dict(), dict(affinity="ASS"), True),
        ("path_TC_k3_t4", 1, 12, 16, 3, 4, dict(), dict(affinity="TC"), True),
        ("path_clip_k3_t4", 1, 12, 16, 3, 4, dict(signed=True), dict(always_clip=True), True),
        # always_clip with EXACT zeros in the state: torch.clamp(min=0) passes the gradient at x == 0
        ("path_clipzero_k3_t4", 1, 16, 24, 3, 4, dict(signed=True, off_sigma=1.0), dict(always_clip=True), True),
        # the fork's default: no offsets -> fixed-local 3x3 propagation (nlspnmodel.py:209-224)
        ("path_fixedlocal_k3_t6", 2, 14, 20, 3, 6, dict(), dict(offset=False), True),
        ("path_fixedlocal_noconf_k3_t3", 1, 9, 13, 3, 3, dict(signed=True), dict(offset=False, conf_prop=False), False),
    ]
    for name, B, H, W, K, T, ikw, mkw, use_conf in specs:
        g = torch.Generator().manual_seed(SEED + len(name))
        model = shell_model(prop_kernel=K, prop_time=T, **mkw)
        inp = make_inputs(g, B, H, W, K, **ikw)
        if name.startswith("path_stable_k3"):
            # plant edge conditions: exact-integer offsets putting taps on -1, 0, H-1, H
            o = inp["guidance"]
            o[0, 0, 0, :] = 0.0      # tap 0 dh: row 0 -> h_im = -1 exactly
            o[0, 1, :, 0] = 0.0      # tap 0 dw: col 0 -> w_im = -1 exactly
            o[0, 0, 5, :] = -4.0     # h_im = 0 exactly at row 5
            o[0, 12, H - 1, :] = 0.0  # tap 6 (after insert tap 7) dh: row H-1 -> h_im = H
            o[1, 2, 3, :] = 1.0
            o[1, 4:6, 7, :] = 40.0   # far out of range
        if name.startswith("path_clipzero"):
            # a block of exact zeros wider than T iterations can reach across (ReLU head output, no sparse depth)
            inp["feat_init"][:, :, :, :14] = 0.0
            inp["feat_fix"][:, :, :, :14] = 0.0
            inp["guidance"][:, :2 * (K * K - 1)].clamp_(-1.5, 1.5)
        with_grad = True
        rec = run_path(model, inp, with_grad=with_grad, use_conf=use_conf)
        a = model.args
        save(name, {k: v for k, v in inp.items() if use_conf or k != "confidence"}, rec,
             dict(K=K, T=T, affinity=a.affinity, preserve=int(a.preserve_input),
                  use_conf=int(use_conf), always_clip=int(a.always_clip), use_offset=int(a.offset),
                  gamma=float(model.aff_scale_const)))

    # zero guidance (zero_init_aff, config.py:233-236): every top/left border tap on -1
    g = torch.Generator().manual_seed(SEED + 99)
    model = shell_model(prop_kernel=3, prop_time=4)
    inp = make_inputs(g, 1, 10, 14, 3)
    inp["guidance"].zero_()
    rec = run_path(model, inp)
    save("path_zero_guidance_k3_t4", inp, rec,
         dict(K=3, T=4, affinity="TGASS", preserve=1, use_conf=1, always_clip=0,
              gamma=float(model.aff_scale_const)))


def case_dcn():
    """Single ModulatedDeformConvFunction.apply calls (boundary B1), incl. non-trivial w, b."""
    mod = RH.import_reference()
    Fn = mod.ModulatedDeformConvFunction
    for name, B, H, W, K, scale in [("dcn_k3", 2, 9, 11, 3, 2.0), ("dcn_k5", 1, 8, 10, 5, 3.0),
                                    ("dcn_k3_far", 1, 7, 9, 3, 10.0)]:
        g = torch.Generator().manual_seed(SEED + len(name) + K)
        x = torch.randn(B, 1, H, W, generator=g).requires_grad_(True)
        off = (scale * torch.randn(B, 2 * K * K, H, W, generator=g))
        off[0, 0, 0, :] = 0.0
        off[0, 1, :, 0] = 0.0
        off[0, 2, 2, :] = 1.0
        off = off.requires_grad_(True)
        msk = torch.randn(B, K * K, H, W, generator=g).requires_grad_(True)
        w = torch.randn(1, 1, K, K, generator=g).requires_grad_(True)
        b = torch.randn(1, generator=g).requires_grad_(True)
        out = Fn.apply(x, off, msk, w, b, 1, (K - 1) // 2, 1, 1, 1, 64)
        go = torch.randn(out.shape, generator=g)
        out.backward(go)
        d = dict(in_x=x, in_off=off, in_msk=msk, in_w=w, in_b=b, in_gout=go, out_y=out,
                 out_gx=x.grad, out_goff=off.grad, out_gmsk=msk.grad, out_gw=w.grad,
                 out_gb=b.grad)
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, meta_K=np.asarray(K),
                            **{k: v.detach().numpy() for k, v in d.items()})
        print("wrote %-40s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


GRU_MODULES = ("GRU", "encode_aff", "encode_dep", "decode_aff")


def case_gru():
    """The fork's GRU mode (nlspnmodel.py:365-373; its DEFAULT configuration, src/config.py:225-232): a real
    ``NLSPNModel(use_GRU=True, use_S2D=True).forward`` + backward.  Stored: the three head outputs (captured by
    hooks), the small GRU-side sub-modules' parameters, every intermediate prediction, the last re-estimated
    affinities, and the reference's autograd gradients wrt the head outputs, gamma and the GRU-side parameters."""
    for name, offset in (("gru_offset_k3_t4", True), ("gru_fixedlocal_k3_t4", False)):
        torch.manual_seed(SEED + len(name))
        B, H, W, K, T = 1, 24, 32, 3, 4
        model = RH.build_reference_model(network="resnet18", prop_kernel=K, prop_time=T, offset=offset,
                                         use_GRU=True, use_S2D=True, patch_height=H, patch_width=W)
        model.train(False)
        with torch.no_grad():
            model.off_aff_dec0[0].weight.mul_(40.0)       # fresh heads give tiny guidance: make offsets matter
        cap = {}

        def grab(key):
            def hook(m, i, o):
                o.retain_grad()
                cap[key] = o
            return hook
        hooks = [model.id_dec0.register_forward_hook(grab("feat_init")),
                 model.off_aff_dec0.register_forward_hook(grab("guidance")),
                 model.cf_dec0.register_forward_hook(grab("confidence"))]
        g = torch.Generator().manual_seed(SEED + 5)
        gt = smooth_field(g, B, H, W, 0.5, 10.0)
        dep = gt * (torch.rand(B, 1, H, W, generator=g) < 0.1).float()
        rgb = torch.randn(B, 3, H, W, generator=g)
        out = model({"rgb": rgb, "dep": dep})
        for h in hooks:
            h.remove()
        loss = loss_fn(out, gt)
        loss.backward()
        d = {"in_feat_init": cap["feat_init"], "in_guidance": cap["guidance"], "in_confidence": cap["confidence"],
             "in_feat_fix": dep, "in_gt": gt,
             "out_pred": out["pred"], "out_list_feat": torch.stack(out["pred_inter"], 0), "out_aff": out["aff"],
             "out_conf_fixed": out["confidence"], "out_loss": loss,
             "out_g_feat_init": cap["feat_init"].grad, "out_g_guidance": cap["guidance"].grad,
             "out_g_confidence": cap["confidence"].grad, "out_g_gamma": model.aff_scale_const.grad}
        if offset:
            d["out_offset"] = out["offset"]
        for k, v in model.state_dict().items():
            if k.split(".")[0] in GRU_MODULES:
                d["param_" + k] = v
        for k, v in model.named_parameters():
            if k.split(".")[0] in GRU_MODULES:
                d["pgrad_" + k] = v.grad
        meta = dict(K=K, T=T, affinity="TGASS", preserve=1, use_conf=1, always_clip=0, use_offset=int(offset),
                    gamma=float(model.aff_scale_const.detach()), max_depth=float(model.args.max_depth))
        path = os.path.join(OUT, name + ".npz")
        np.savez_compressed(path, **{k: v.detach().numpy() for k, v in d.items()},
                            **{"meta_" + k: np.asarray(v) for k, v in meta.items()})
        print("wrote %-40s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    only = sys.argv[1:]          # optional: names of the case groups to (re)generate
    for fn in (case_fullmodel, case_paths, case_dcn, case_gru):
        if not only or fn.__name__ in only:
            fn()
