"""Full-model inference (NLSPNModel.forward under no_grad) with and without the head GEMM / fused prologue (tool).
    python tools/model_inference.py [kitti|nyu] [B] [network]
Prints one JSON line: ms per forward for fused_heads=False (stock torch heads), fused heads + separate prologue, and heads
+ prologue in one kernel; plus the head / propagation kernel times from the library's own profile."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nlspn_eccv20_b200 import _lib  # noqa: E402
from nlspn_eccv20_b200.model import NLSPNModel  # noqa: E402
from nlspn_eccv20_b200.synth import SHAPES, workload  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "kitti"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
network = sys.argv[3] if len(sys.argv) > 3 else "resnet34"
dev = torch.device("cuda:0")
H, W, md = SHAPES[wl]
d = workload(wl, B, 3, seed=7240, device=dev)
sample = {"rgb": torch.randn(B, 3, H, W, device=dev), "dep": d["feat_fix"]}
torch.manual_seed(0)
net = NLSPNModel(network=network, prop_kernel=3, prop_time=18, max_depth=md).to(dev).eval()


def timed(n=10):
    with torch.no_grad():
        for _ in range(3):
            net(sample)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            net(sample)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"workload": "%s %dx%d B=%d %s K=3 T=18 inference" % (wl, H, W, B, network)}
for name, fh, fp in (("stock_heads_ms", False, False), ("fused_heads_ms", True, False), ("fused_heads_prologue_ms", True, "auto")):
    net.fused_heads, net.fused_prologue = fh, fp
    out[name] = timed()
lib = _lib.load()
lib.nlspn_profile_enable(1)
with torch.no_grad():
    for _ in range(3):
        net(sample)
torch.cuda.synchronize()
out["kernel_ms"] = {k: round(v[0] / 3, 4) for k, v in _lib.profile_read().items()}
lib.nlspn_profile_enable(0)
with torch.no_grad():
    a = net(sample)["pred"]
    net.fused_heads = False
    b = net(sample)["pred"]
out["max_abs_pred_diff_vs_stock_heads_m"] = float((a - b).abs().max())
out["max_abs_pred_m"] = float(b.abs().max())
out["median_abs_pred_diff_m"] = float((a - b).abs().median())
with torch.no_grad():
    net.fused_heads = True
    ha = net.heads(sample["rgb"], sample["dep"])
    net.fused_heads = False
    hb = net.heads(sample["rgb"], sample["dep"])
out["head_outputs_max_abs_diff"] = [float((x - y).abs().max()) for x, y in zip(ha, hb)]
out["head_outputs_max_abs"] = [float(y.abs().max()) for y in hb]
print(json.dumps(out))
