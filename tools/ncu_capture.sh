#!/bin/bash
# One `ncu --set full` capture of each main kernel (one launch each) of the headline step; reports land in
# gpurun_out/ as .ncu-rep (read here with `ncu -i ... --page raw --csv`).  Run AFTER the plain bench exited 0.
set -u
tag=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-ref-cuda > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err || { echo "plain bench failed"; exit 1; }
for k in bwd_state_local_kernel sched_build_kernel bwd_param_tiled_kernel iter_fwd_tiled_kernel prologue_fwd_kernel final_bwd_kernel; do
  skip=3; case $k in bwd_state_local_kernel|iter_fwd_tiled_kernel) skip=20;; esac
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip $skip --launch-count 1 \
      -f -o /tmp/${tag}_$k python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-ref-cuda > gpurun_out/${tag}_ncu_$k.log 2>&1
  # the reports are ~30 MB each (gpurun_out/ carries 64 MiB): export the two pages we read, drop the report
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_$k.raw.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page source --csv > gpurun_out/${tag}_$k.source.csv 2>/dev/null
  ncu -i /tmp/${tag}_$k.ncu-rep --page details > gpurun_out/${tag}_$k.details.txt 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs --no-ref-cuda > gpurun_out/${tag}_ncu_launches.log 2>&1
ls -la gpurun_out/ | tail -30
