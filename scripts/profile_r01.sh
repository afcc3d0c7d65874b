#!/bin/bash
# round-1 profiling pass (B200_PROFILING.md recipe): plain run first, then the launch list and one full capture per top kernel
CMD="python bench.py --views 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
TAG=${1:-r01b}
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 128 -c 200 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 8 -c 1 -f -o gpurun_out/prof_sweep_$TAG $CMD > gpurun_out/ncu_sweep_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fuse_view -s 8 -c 1 -f -o gpurun_out/prof_fuse_$TAG $CMD > gpurun_out/ncu_fuse_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_filter -s 40 -c 2 -f -o gpurun_out/prof_filter_$TAG $CMD > gpurun_out/ncu_filter_$TAG.log 2>&1
ls -la gpurun_out/
