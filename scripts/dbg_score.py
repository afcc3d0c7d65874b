import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
syn, osc, gt, imgs, ok = common.make_scene(1, 0.5)
ctx = common.make_context(syn, osc, imgs, ok)
ref = 0
for sampler in (0, 1):
    ctx.set_params(sampler=sampler)
    for k, (ds, ang) in enumerate(((0.0, 0.0), (0.004, 4.0), (0.03, 20.0))):
        d, n = common.perturbed_hypotheses(gt[ref][0], gt[ref][1], syn.K[ref], seed=10 * ref + k, depth_sigma=ds, angle_deg=ang)
        for smooth in (0, 1):
            want = osc.score_hypotheses(ref, d, n, smooth)
            got = ctx.score_hypotheses(ref, d, n, smooth)
            err = np.abs(want - got)
            idx = np.argsort(err.ravel())[::-1][:5]
            print(f"sampler {sampler} case {k} smooth {smooth}: max {err.max():.3e} p99.9 {np.percentile(err, 99.9):.3e} median {np.median(err):.3e} n>1e-4 {(err > 1e-4).sum()} n>1e-5 {(err > 1e-5).sum()}")
            for i in idx[:3]:
                y, x = divmod(int(i), err.shape[1])
                print(f"    ({y},{x}) want {want[y, x]:.6f} got {got[y, x]:.6f} depth {d[y, x]:.4f}")
