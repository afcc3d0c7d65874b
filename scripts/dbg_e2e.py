import sys, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hcmvs_b200 import api, host
from hcmvs_b200.synth import SynthScene
syn = SynthScene(2, 1.0, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
imgs = [syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views)]
params = dict(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
ctx = api.Context(0, **params)
for rep in range(4):
    hs = host.HostScene.from_synth(syn, imgs)
    t0 = time.time()
    st = hs.dense_reconstruction(ctx, seed=1, run_filter=True)
    t1 = time.time()
    print(rep, "total %.3f" % (t1 - t0), {k: round(v, 3) if isinstance(v, float) else v for k, v in st.items()})
