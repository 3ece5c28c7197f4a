#!/bin/bash
# ncu --set full capture of the gather-form pass-A kernels on config 5's per-GPU shard (K=5, T=36, B=8).
set -u
tag=${1:-r01k5}
args="--kernel 5 --iters 36 --steps 1 --warmup 1 --no-cpu-baseline"
python bench.py $args > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain bench failed"; exit 1; }
for k in bwd_gather_kernel bwd_gy_kernel table_build_kernel table_compact_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 0 --launch-count 1 \
      -f -o /tmp/${tag}_$k python bench.py $args > gpurun_out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_$k.raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page source --csv > gpurun_out/${tag}_$k.source.csv 2>/dev/null
done
ls -la gpurun_out | grep ${tag} | head -20
