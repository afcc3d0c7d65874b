#!/bin/bash
# one N = 1 bench line per BASELINE config (C2 is the default run); C5 with --views to bound the set-up time
for c in 1 3 4; do
  extra=""; [ $c = 4 ] && extra="--views ${C4_VIEWS:-60}"
  timeout 1500 python bench.py --config $c --steps 2 --warmup 1 $extra > gpurun_out/r02_bench_c$c.json 2> gpurun_out/r02_bench_c$c.err || tail -3 gpurun_out/r02_bench_c$c.err
  tail -c 400 gpurun_out/r02_bench_c$c.json; echo
done
timeout 1500 python bench.py --config 5 --steps 2 --warmup 1 --views ${C5_VIEWS:-100} > gpurun_out/r02_bench_c5.json 2> gpurun_out/r02_bench_c5.err || tail -3 gpurun_out/r02_bench_c5.err
tail -c 600 gpurun_out/r02_bench_c5.json; echo
