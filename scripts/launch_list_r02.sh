#!/bin/bash
# launch list (gpu__time_duration per launch) of the 8-view C2 bench command, per B200_PROFILING.md
CMD="python bench.py --views 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r02_plain_views8.log 2>&1 || { tail -5 gpurun_out/r02_plain_views8.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_bench_views8.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log
