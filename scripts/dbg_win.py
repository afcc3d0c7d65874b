import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, common
syn, osc, gt, imgs, ok = common.make_scene(2, 1.0, 8)
ctx = common.make_context(syn, osc, imgs, ok, sampler=2)
ref = 3
osc.init_depth_sparse(ref)
d0, _, _, lo, hi = osc.get_depthmap(ref)
for iters in (1, 2, 3):
    ctx.set_params(nEstimationIters=iters)
    ctx.init_depthmap(ref, d0, None, lo, hi)
    ctx.reset_timers()
    ctx.estimate_depthmap(ref, 0, 23)
    t = ctx.timers()
    print("iters", iters, "share %.3f" % (t["n_window_walks"] * 32.0 / max(t["n_view_scores"], 1)), "sweeps ms %.2f" % t["ms_sweeps"], flush=True)
