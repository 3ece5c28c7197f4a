import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import functional as F_
from nlspn_eccv20_b200.synth import workload
dev = torch.device("cuda:0")
inp = workload("nyu", 1, 3, device=dev)
gam = torch.full((1,), 4.0, device=dev)
for T in (1, 2, 6, 18, 36):
    def run():
        return F_.forward(inp["guidance"], inp["confidence"], inp["feat_init"], inp["feat_fix"], gam, 3, T, keep_src=False)
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    print("PERSIST=%s T=%d  %.1f us per forward" % (os.environ.get("NLSPN_PERSIST", "auto"), T, e0.elapsed_time(e1) * 1000 / 50))
