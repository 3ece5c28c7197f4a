timeout 300 python -m pytest tests/test_gpu_heads.py -x -q 2>&1 | tail -3
timeout 120 python tools/head_baseline.py 8 > /tmp/a.json; python - <<'P'
import json
d=json.load(open('/tmp/a.json')); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith("fused") or k.startswith("kernel")})
P
timeout 100 python tools/head_baseline.py 12 228 304 3 > /tmp/b.json; python - <<'P'
import json
d=json.load(open('/tmp/b.json')); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith("fused") or k.startswith("heads_ms")})
P
# training path: gradient kernels alone, then forward + backward against the stock layers
timeout 100 python tools/head_wgrad_bench.py 8
timeout 100 python tools/head_train_baseline.py 8
timeout 100 python tools/head_train_baseline.py 12 228 304
