"""Debug: fuse with nMinViewsFuse = 3 against the oracle, report where the clouds differ."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import common
import test_gpu_filter_fuse as T
syn, osc, gt, imgs, ok = common.make_scene(2, 0.2, 12)
ctx = common.make_context(syn, osc, imgs, ok)
maps = [T.noisy_maps(gt[i], 100 + i) for i in range(syn.n_views)]
for nmin in (3, 2, 1):
    T.load_maps(syn, osc, ctx, maps)
    osc.set_params(nMinViewsFuse=nmin); ctx.set_params(nMinViewsFuse=nmin)
    want = osc.fuse(True, True)
    got = ctx.fuse_depthmaps(True, True)
    T.load_maps(syn, osc, ctx, maps)
    got2 = ctx.fuse_depthmaps(True, True)
    print("nMin", nmin, "points", len(want["xyz"]), len(got["xyz"]), "deterministic:", np.array_equal(got["views"], got2["views"]) and np.array_equal(got["n_views"], got2["n_views"]))
    if len(want["xyz"]) == len(got["xyz"]):
        bad = np.nonzero(want["n_views"] != got["n_views"])[0]
        print("  n_views differ at", len(bad), "points; first", bad[:10])
        if len(bad):
            ow = np.concatenate([[0], np.cumsum(want["n_views"])]); og = np.concatenate([[0], np.cumsum(got["n_views"])])
            for b in bad[:5]:
                print("   point", b, "want views", want["views"][ow[b]:ow[b+1]], "got", got["views"][og[b]:og[b+1]], "xyz", want["xyz"][b], got["xyz"][b])
