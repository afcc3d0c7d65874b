#!/bin/bash
# ncu --set full of the persistent fusion kernel on a 16-view slice of C2 (same per-view shape as the full scene)
CMD="python bench.py --views 16 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
HCMVS_FUSE_DEBUG=1 $CMD > gpurun_out/r02_fuse_plain.log 2>&1 || { tail -5 gpurun_out/r02_fuse_plain.log; exit 1; }
grep "\[fuse\]" gpurun_out/r02_fuse_plain.log | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_fuse_scene -s 1 -c 1 -f -o gpurun_out/r02_prof_fuse_scene $CMD > gpurun_out/r02_ncu_fuse.log 2>&1
tail -3 gpurun_out/r02_ncu_fuse.log
ls -la gpurun_out/*.ncu-rep | tail -3
