#!/usr/bin/env python
"""Runs bench.py (no CPU baseline) with the given extra args and prints value, ms/step and the per-kernel ms."""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline"] + sys.argv[1:],
                     capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print("%.3f Gpix*iter/s  %.3f ms/step  e2e %.2f  %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"],
        {k: round(v["step_ms"], 3) for k, v in d["kernels"].items()}))
except Exception:
    print(out.stdout[-2000:], out.stderr[-3000:])
