#!/bin/bash
# ncu --set full (+ source) of one k_sweep launch of the 8-view C2 bench: second outer launch of view 3 (iteration 2: converging estimates)
CMD="python bench.py --views 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline ${SWEEP_ARGS}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 60 -c 1 -f -o gpurun_out/r02_prof_k_sweep${SWEEP_TAG} $CMD > gpurun_out/r02_ncu_sweep.log 2>&1
tail -2 gpurun_out/r02_ncu_sweep.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
