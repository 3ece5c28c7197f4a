set -u
for k in bwd_param_tiled_kernel final_bwd_kernel prologue_fwd_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 0 --launch-count 1 \
      -f -o /tmp/r01d_$k python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01d_ncu_$k.log 2>&1
  ncu -i /tmp/r01d_$k.ncu-rep --page raw --csv > gpurun_out/r01d_$k.raw.csv 2>/dev/null
done
ls -la gpurun_out | grep r01d | head
