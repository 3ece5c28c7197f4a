"""Weight gradients of the three head layers (tool): cuDNN's four calls (as heads.FusedHeadsFunction ran them before) vs
nlspn_heads_grad_prep + nlspn_heads_wgrad.     python tools/head_wgrad_bench.py [B] [H] [W] [K]"""
import json
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import heads  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 352
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1216
K = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N3 = 3 * (K * K - 1)
dev = torch.device("cuda:0")
x = [torch.randn(B, 64, H, W, device=dev) for _ in range(4)]
pred_init = torch.relu(torch.randn(B, 1, H, W, device=dev))
confidence = torch.sigmoid(torch.randn(B, 1, H, W, device=dev))
gi, gg, gc = (torch.randn(B, n, H, W, device=dev) for n in (1, N3, 1))


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stock():
    a = gi * (pred_init > 0)
    c = gc * confidence * (1.0 - confidence)
    g_all = torch.cat((a, c, gg), 1)
    outs = [torch.nn.grad.conv2d_weight(t, (n, 64, 3, 3), g, stride=1, padding=1) for t, n, g in
            ((x[0], 1, a), (x[1], N3, gg), (x[2], 1, c), (x[3], N3 + 2, g_all))]
    return outs, [g.sum(dim=(0, 2, 3)) for g in (a, gg, c)]


state = {}


def prep():
    state["g"] = heads.grad_prep(pred_init, confidence, gi, gg, gc, K)


def wgrad():
    state["dw"] = heads.weight_grads(x[0], x[1], x[2], x[3], state["g"][0], K)


out = {"shape": [B, H, W, K]}
out["cudnn_wgrad_bias_ms"] = timed(stock)
out["grad_prep_ms"] = timed(prep)
out["wgrad_ms"] = timed(wgrad)
for name, args in (("fe1", (None, None, None, x[3])), ("fe1+oa", (None, x[1], None, x[3])), ("fe1+id", (x[0], None, None, x[3]))):
    out["wgrad_%s_ms" % name] = timed(lambda: heads.weight_grads(*args, state["g"][0], K))
w1 = [torch.randn(1, 128, 3, 3, device=dev) for _ in range(2)]
out["dgrad_one_ms"] = timed(lambda: heads.dgrad_one(state["g"][0][1], w1[0], w1[1], K))
out["cudnn_dgrad_one_ms"] = timed(lambda: [torch.nn.grad.conv2d_input((B, 64, H, W), w[:, :64].contiguous(), gg[:, :1].contiguous(), stride=1, padding=1) for w in w1])
if heads.dgrad_supported(W, K):
    w3 = [torch.randn(n, 128, 3, 3, device=dev) for n in (1, N3, 1)]
    out["dgrad_wide_ms"] = timed(lambda: heads.dgrad_wide(state["g"][0], w3[0], w3[1], w3[2], K))
    wfe = torch.cat((w3[0][:, 64:], w3[2][:, 64:], w3[1][:, 64:]), 0).contiguous()
    woa = w3[1][:, :64].contiguous()
    out["cudnn_dgrad_wide_ms"] = timed(lambda: (torch.nn.grad.conv2d_input((B, 64, H, W), woa, gg, stride=1, padding=1),
                                                torch.nn.grad.conv2d_input((B, 64, H, W), wfe, state["g"][0][1], stride=1, padding=1)))
ref, _ = stock()
dw = state["dw"]
out["rel_diff_fe1"] = float((dw[:, 64:] - ref[3]).abs().max() / ref[3].abs().max())
out["rel_diff_oa"] = float((dw[2:, :64] - ref[1]).abs().max() / ref[1].abs().max())
print(json.dumps(out))
