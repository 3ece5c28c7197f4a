#!/bin/bash
# BASELINE.json configs on one B200 (configs 3/5 per-GPU shards); writes one JSON line per run.
out=${1:-gpurun_out/configs_r01.jsonl}
: > $out
run() { echo "# $1" >> $out; shift; python bench.py --no-cpu-baseline "$@" 2>/dev/null | tail -1 >> $out; }
run "config1 NYU B=1 fwd (L2 flushed)"            --workload nyu --batch 1 --mode fwd --steps 20
run "config1 NYU B=1 fwd (unflushed)"             --workload nyu --batch 1 --mode fwd --steps 20 --flush-l2 off
run "config1 NYU B=1 fwd, CUDA graph replay (L2 flushed)" --workload nyu --batch 1 --mode fwd --steps 20 --cuda-graph
run "config2 NYU B=12 fwd+bwd (L2 flushed)"       --workload nyu --batch 12
run "config2 NYU B=12 fwd+bwd (unflushed)"        --workload nyu --batch 12 --flush-l2 off
run "config3 KITTI B=16 fwd (1 GPU)"              --workload kitti --batch 16 --mode fwd
run "config3 KITTI B=8 fwd (shard of 2 GPUs)"     --workload kitti --batch 8 --mode fwd
run "config3 KITTI B=4 fwd (shard of 4 GPUs)"     --workload kitti --batch 4 --mode fwd
run "config3 KITTI B=2 fwd (shard of 8 GPUs)"     --workload kitti --batch 2 --mode fwd
run "config5 KITTI B=8/GPU K=5 T=36 fwd+bwd"      --workload kitti --batch 8 --kernel 5 --iters 36 --steps 3
run "headline KITTI B=8 K=3 T=18 fwd+bwd"         --workload kitti --batch 8
run "headline, smooth offsets"                    --workload kitti --batch 8 --smooth-offsets
python bench.py --workload nyu --batch 1 --mode fwd --impl reference --steps 2 --warmup 1 --cpu-rows 228 2>/dev/null | tail -1 >> $out
