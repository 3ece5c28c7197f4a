"""How long do the three final head convolutions of the reference take in stock PyTorch on this GPU?  (tool)
nlspnmodel.py:69-86,297,301,313: pred_init = relu(conv3x3(cat(id_fd1, fe1))), off_aff = conv3x3(cat(oa_fd1, fe1)),
confidence = sigmoid(conv3x3(cat(cf_fd1, fe1))); 128 -> 1 / 3N / 1 channels."""
import json
import sys
import torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 352
W = int(sys.argv[3]) if len(sys.argv) > 3 else 1216
K = int(sys.argv[4]) if len(sys.argv) > 4 else 3
N = K * K - 1
dev = torch.device("cuda:0")
fe1, id1, oa1, cf1 = (torch.randn(B, 64, H, W, device=dev) for _ in range(4))
c_id = torch.nn.Conv2d(128, 1, 3, 1, 1).to(dev)
c_oa = torch.nn.Conv2d(128, 3 * N, 3, 1, 1).to(dev)
c_cf = torch.nn.Conv2d(128, 1, 3, 1, 1).to(dev)


def heads(channels_last=False):
    a = torch.relu(c_id(torch.cat((id1, fe1), 1)))
    b = c_oa(torch.cat((oa1, fe1), 1))
    c = torch.sigmoid(c_cf(torch.cat((cf1, fe1), 1)))
    return a, b, c


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {"shape": [B, H, W, K]}
with torch.no_grad():
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        out["heads_ms_tf32=%s" % tf32] = timed(heads)
    torch.backends.cudnn.allow_tf32 = True
    out["cat_only_ms"] = timed(lambda: (torch.cat((id1, fe1), 1), torch.cat((oa1, fe1), 1), torch.cat((cf1, fe1), 1)))
    x = torch.cat((oa1, fe1), 1)
    out["conv_oa_only_ms"] = timed(lambda: c_oa(x))
    xcl = x.contiguous(memory_format=torch.channels_last)
    c_oa_cl = c_oa.to(memory_format=torch.channels_last)
    out["conv_oa_channels_last_ms"] = timed(lambda: c_oa_cl(xcl))
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import heads, _lib  # noqa: E402
args = (id1, oa1, cf1, fe1, c_id.weight, c_id.bias, c_oa.weight, c_oa.bias, c_cf.weight, c_cf.bias)
with torch.no_grad():
    out["fused_tcgen05_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
    with _lib.options(heads_ks=1):
        out["fused_tcgen05_8ch_stages_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
    with _lib.options(heads_ks=1, heads_reuse=0):
        out["fused_tcgen05_no_collector_reuse_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
    with _lib.options(heads_persist=0):
        out["fused_tcgen05_cta_per_tile_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
    with _lib.options(heads_rows=0):
        out["fused_tcgen05_ninetap_form_ms"] = timed(lambda: heads.fused_heads(*args, prop_kernel=K))
    fix = (torch.rand(B, 1, H, W, device=dev) < 0.05).float() * 5.0
    gam = torch.tensor([0.5 * N], device=dev)
    out["fused_heads_prologue_ms"] = timed(lambda: heads.fused_heads_prologue(*args, fix, gam, K))
    a = heads.fused_heads(*args, prop_kernel=K)
    r = heads.reference_heads(*args)
    out["max_abs_diff_vs_cudnn_tf32"] = [float((x - y).abs().max()) for x, y in zip(a, r)]
    lib = _lib.load()
    lib.nlspn_profile_enable(1)
    for _ in range(5):
        heads.fused_heads(*args, prop_kernel=K)
    torch.cuda.synchronize()
    out["kernel_only_ms"] = {k: v[0] / 5 for k, v in _lib.profile_read().items()}
    lib.nlspn_profile_enable(0)
print(json.dumps(out))
