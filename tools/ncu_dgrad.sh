#!/bin/bash
# ncu --set full of the heads' wide data-gradient GEMM at KITTI B=8 (tool)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:head_dgrad_wide --launch-skip 3 --launch-count 1 \
    -o /tmp/r02_head_dgrad -f python tools/head_wgrad_bench.py 8 > gpurun_out/r02_head_dgrad_ncu.log 2>&1
ncu -i /tmp/r02_head_dgrad.ncu-rep --page raw --csv > gpurun_out/r02_head_dgrad.raw.csv 2>/dev/null
ncu -i /tmp/r02_head_dgrad.ncu-rep --page source --csv > gpurun_out/r02_head_dgrad.source.csv 2>/dev/null
ncu -i /tmp/r02_head_dgrad.ncu-rep --page details > gpurun_out/r02_head_dgrad.details.txt 2>/dev/null
