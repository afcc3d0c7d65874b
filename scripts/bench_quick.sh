#!/bin/bash
# usage: scripts/bench_quick.sh <extra bench args...>  -> one summary line
python bench.py --views 8 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c '
import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]; s=d["stage_ms_per_step_rank0"]
print("value %.1f sweep %.1f Mpix*it/s frac %.3f hyp/px-it %.2f | ms: prep %.1f score %.1f sweeps %.1f end %.2f filter %.1f fuse %.1f | pts %d" % (d["value"], r["sweep_mpix_iter_s"], r["frac"], r["hyp_per_pixel_iter"], s["ms_prep"], s["ms_score"], s["ms_sweeps"], s["ms_end"], s["ms_filter"], s["ms_fuse"], d["fused_points"]))'
