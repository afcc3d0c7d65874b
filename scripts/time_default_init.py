"""C2 end to end with the reference's DEFAULT initialisation (nMinViewsTrustPoint = 2: triangulated sparse points) vs the splat start."""
import sys, time; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from hcmvs_b200 import api, host
from hcmvs_b200.synth import SynthScene
syn = SynthScene(2, 1.0, 0)
rendered = [syn.render(i) for i in range(syn.n_views)]
imgs = [r[0] for r in rendered]
for trust in (2, 1):
    ctx = api.Context(0, nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=trust, adapthalfwin=5)
    host.HostScene.from_synth(syn, imgs).dense_reconstruction(ctx, seed=1, run_filter=True)  # warm-up
    hs = host.HostScene.from_synth(syn, imgs)
    torch.cuda.synchronize(); t0 = time.time()
    st = hs.dense_reconstruction(ctx, seed=1, run_filter=True)
    torch.cuda.synchronize(); dt = time.time()-t0
    ref = 24
    d, n, c, _, _ = ctx.get_depthmap(ref)
    g, gn = rendered[ref][1], rendered[ref][2]
    valid = (d > 0) & (g > 0)
    within = np.mean(np.abs(d[valid]/g[valid]-1) < 0.01)
    ang = np.degrees(np.arccos(np.clip((n*gn).sum(axis=2)[valid], -1, 1)))
    print(f"nMinViewsTrustPoint={trust}: e2e {dt:.3f} s, {st['n_points']} points, view {ref}: kept {(d>0).mean():.3f}, within 1% {within:.4f}, median normal error {np.median(ang):.2f} deg, h2d {st['h2d_bytes']/1e6:.0f} MB", flush=True)
    hs.close(); ctx.close()
