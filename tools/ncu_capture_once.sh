#!/bin/bash
# `ncu --set full` of the once-per-step kernels at the headline shape (B = 8): tools/quick_step.py runs only
# whole-batch steps (bench.py's end-to-end leg feeds single-frame chunks, which is not the shape to profile)
set -u
tag=${1:-r02}
mkdir -p gpurun_out
for k in sched_build_kernel bwd_param_tiled_kernel prologue_fwd_kernel final_bwd_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
      -f -o /tmp/${tag}_$k python tools/quick_step.py --steps 1 > gpurun_out/${tag}_ncu_$k.log 2>&1
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_$k.raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page source --csv > gpurun_out/${tag}_$k.source.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_$k.details.txt 2>/dev/null
done
