#!/bin/bash
# ncu --set full capture of the pass-A kernels (tool): run under gpurun from the repo root
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"bwd_state_local_kernel|sched_build_kernel|bwd_param_tiled" \
    --launch-skip 12 --launch-count 4 -o gpurun_out/r02_local_full -f python tools/quick_step.py --steps 1 > gpurun_out/r02_local_ncu.log 2>&1
ncu -i gpurun_out/r02_local_full.ncu-rep --page raw --csv > gpurun_out/r02_local_full_raw.csv 2>/dev/null
