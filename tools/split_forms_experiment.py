#!/usr/bin/env python
"""Experiment: backward of a batch split in two halves that run CONCURRENTLY on two streams, one half with
pass A as RED scatter (write-port bound), the other as tabulated gather (HBM-read bound)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlspn_eccv20_b200 import functional as F_
from nlspn_eccv20_b200.synth import workload

dev = torch.device("cuda:0")
K, T = 3, 18
gamma = torch.full((1,), 4.0, device=dev)


def prep(B, seed):
    inp = workload("kitti", B, K, seed=seed, device=dev)
    offset, aff, cfx, src, lf = F_.forward(inp["guidance"], inp["confidence"], inp["feat_init"], inp["feat_fix"], gamma, K, T)
    g_list = [None] * (T - 1) + [torch.ones_like(lf[0])]
    return (inp["guidance"], inp["feat_init"], inp["feat_fix"], offset, aff, cfx, src, lf, g_list, gamma, K, T)


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


full = prep(8, 1)
ha, hb = prep(4, 2), prep(4, 3)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def whole(form):
    os.environ["NLSPN_STATE_GATHER"] = form
    F_.backward(*full)


def halves(fa, fb, concurrent):
    cur = torch.cuda.current_stream()
    if concurrent:
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            os.environ["NLSPN_STATE_GATHER"] = fa
            F_.backward(*ha)
        with torch.cuda.stream(s2):
            os.environ["NLSPN_STATE_GATHER"] = fb
            F_.backward(*hb)
        cur.wait_stream(s1); cur.wait_stream(s2)
    else:
        os.environ["NLSPN_STATE_GATHER"] = fa
        F_.backward(*ha)
        os.environ["NLSPN_STATE_GATHER"] = fb
        F_.backward(*hb)


print("B=8 one call, RED    : %.3f ms" % timed(lambda: whole("0")))
print("B=8 one call, gather : %.3f ms" % timed(lambda: whole("1")))
print("4+4 sequential RED+RED       : %.3f ms" % timed(lambda: halves("0", "0", False)))
print("4+4 concurrent RED || RED    : %.3f ms" % timed(lambda: halves("0", "0", True)))
print("4+4 concurrent RED || gather : %.3f ms" % timed(lambda: halves("0", "1", True)))
print("4+4 concurrent gather||gather: %.3f ms" % timed(lambda: halves("1", "1", True)))
