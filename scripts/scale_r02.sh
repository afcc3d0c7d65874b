#!/bin/bash
# usage: scripts/scale_r02.sh N  -> C2 (default config) and C4 (300 views) bench lines at N GPUs
N=$1
run() { # config steps warmup tag
  if [ "$N" = 1 ]; then python bench.py --config $1 --steps $2 --warmup $3 --no-cpu-baseline
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+$1)) bench.py --gpus $N --config $1 --steps $2 --warmup $3; fi
}
[ -z "$SKIP_C2" ] && { run 2 5 3 2> gpurun_out/r02_scale_c2_n$N.err | tail -1 > gpurun_out/r02_scale_c2_n$N.json; python scripts/show_bench.py gpurun_out/r02_scale_c2_n$N.json | head -2; }
[ -z "$SKIP_C4" ] && { run 4 2 1 2> gpurun_out/r02_scale_c4_n$N.err | tail -1 > gpurun_out/r02_scale_c4_n$N.json; python scripts/show_bench.py gpurun_out/r02_scale_c4_n$N.json | head -2; }
python - <<'P'
import json, glob, sys
for f in sorted(glob.glob("gpurun_out/r02_scale_c*_n%s.json" % sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_scale_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); e = d.get("e2e") or {}
        print(f, "value %.1f scene %.3f s | e2e %.1f (%.3f s) h2d %.0f MB" % (d["value"], d["scene_seconds"], e.get("value") or 0, e.get("seconds_per_scene") or 0, (e.get("h2d_bytes_per_step") or 0) / 1e6))
    except Exception as ex:
        print(f, "unreadable", ex)
P
