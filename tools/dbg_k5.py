import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import NLSPN
from nlspn_eccv20_b200.synth import make_inputs
K = int(sys.argv[1]); H = int(sys.argv[2]); W = int(sys.argv[3]); T = int(sys.argv[4]); grad = int(sys.argv[5])
dev = torch.device("cuda:0")
inp = make_inputs(2, H, W, K, seed=1, device=dev)
mod = NLSPN(prop_kernel=K, prop_time=T).to(dev)
fi = inp["feat_init"].requires_grad_(bool(grad))
out = mod(fi, inp["guidance"], inp["confidence"], inp["feat_fix"])
torch.cuda.synchronize()
print("fwd ok", float(out[0].sum()))
if grad:
    out[0].sum().backward()
    torch.cuda.synchronize()
    print("bwd ok", float(fi.grad.sum()))
