#!/usr/bin/env python
"""Print the interesting fields of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"value {d['value']:.1f} {d['unit']}  scene {d.get('scene_seconds', 0):.3f} s  n_gpus {d['n_gpus']}  points {d.get('fused_points')}  launches {d.get('gpu_launches')}")
print("stages ms:", {k: round(v, 1) for k, v in (d.get("stage_ms_per_step_rank0") or {}).items()})
if r:
    print(f"sweep {r['sweep_mpix_iter_s']:.1f} Mpix*it/s  fp32 frac {r['frac']:.3f}  hyp/px-it {r['hyp_per_pixel_iter']:.2f}  avg launch {r['avg_launch_ms']:.2f} ms")
print("cpu:", d.get("cpu_baseline"))
print("e2e:", d.get("e2e"))
print("clocks:", d.get("clocks"))
