#!/bin/bash
# ncu --set full of the fused head kernel at KITTI B=8 (tool)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:head_persist --launch-skip 3 --launch-count 1 \
    -o /tmp/r02_head_persist -f python tools/head_baseline.py 8 > gpurun_out/r02_head_persist_ncu.log 2>&1
ncu -i /tmp/r02_head_persist.ncu-rep --page raw --csv > gpurun_out/r02_head_persist.raw.csv 2>/dev/null
ncu -i /tmp/r02_head_persist.ncu-rep --page source --csv > gpurun_out/r02_head_persist.source.csv 2>/dev/null
