#!/bin/bash
# ncu --set full of the head weight-gradient kernels at KITTI B=8 (tool): the wide launch (fe1 + guidance branch) and
# the one-channel launch (init + confidence branches)
mkdir -p gpurun_out
for sk in 0 1; do
  ncu --set full --clock-control none --import-source on -k regex:head_wgrad_roll --launch-skip $((4 + sk)) --launch-count 1 \
      -o /tmp/r02_head_wgrad_$sk -f python tools/head_wgrad_bench.py 8 > gpurun_out/r02_head_wgrad_ncu_$sk.log 2>&1
  ncu -i /tmp/r02_head_wgrad_$sk.ncu-rep --page raw --csv > gpurun_out/r02_head_wgrad_$sk.raw.csv 2>/dev/null
  ncu -i /tmp/r02_head_wgrad_$sk.ncu-rep --page source --csv > gpurun_out/r02_head_wgrad_$sk.source.csv 2>/dev/null
done
