"""Forward + backward of the three head layers (tool): stock torch (cat + conv) vs heads.fused_heads.
    python tools/head_train_baseline.py [B] [H] [W]"""
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import heads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 352
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1216
K, N3 = 3, 24
dev = torch.device("cuda:0")
x = [torch.randn(B, 64, H, W, device=dev, requires_grad=True) for _ in range(4)]
s = (128 * 9) ** -0.5
w = [(s * torch.randn(n, 128, 3, 3, device=dev)).requires_grad_(True) for n in (1, N3, 1)]
b = [(0.1 * torch.randn(n, device=dev)).requires_grad_(True) for n in (1, N3, 1)]
args = (x[0], x[1], x[2], x[3], w[0], b[0], w[1], b[1], w[2], b[2])
g = None


def step(fn):
    global g
    for t in x + w + b:
        t.grad = None
    o = fn()
    if g is None:
        g = [torch.randn_like(t) for t in o]
        with torch.no_grad():      # no upstream gradient at the ReLU kink (TF32 rounding decides the branch there)
            z0 = torch.nn.functional.conv2d(torch.cat((x[0], x[3]), 1), w[0], b[0], 1, 1)
            g[0] = g[0] * (z0.abs() > 2e-2).to(g[0].dtype)
    torch.autograd.backward(o, g)


def timed(fn, n=5):
    for _ in range(2):
        step(fn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(fn)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"shape": [B, H, W, K]}
out["stock_fwd_bwd_ms"] = timed(lambda: heads.reference_heads(*args))
ref = [t.grad.clone() for t in x + w + b]
out["ours_fwd_bwd_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
ours = [t.grad.clone() for t in x + w + b]
names = ["id_fd1", "oa_fd1", "cf_fd1", "fe1", "w_id", "w_oa", "w_cf", "b_id", "b_oa", "b_cf"]
out["rel_grad_diff"] = {n: float((a - c).abs().max() / c.abs().max().clamp_min(1e-6)) for n, a, c in zip(names, ours, ref)}
with torch.no_grad():
    torch.cuda.synchronize()
print(json.dumps(out))
