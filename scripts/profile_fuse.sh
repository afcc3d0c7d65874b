CMD="python bench.py --views 12 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/s3_plain_fuse.log 2>&1 || { tail -5 gpurun_out/s3_plain_fuse.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_fuse_probe -s 14 -c 1 -f -o gpurun_out/s3_prof_fuse_probe $CMD > gpurun_out/s3_ncu_probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fuse_emit -s 14 -c 1 -f -o gpurun_out/s3_prof_fuse_emit $CMD > gpurun_out/s3_ncu_emit.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
