"""BASELINE.json configs C3 / C4 / C5 as GPU test cases (C1 and C2 are covered by test_gpu_parity / test_gpu_fullsize):

  C3  ETH3D-shaped 6048x4032 views — the maximum image size: one full-resolution reference view, estimation + fusion bookkeeping
  C4  video-frame scene 1920x1080 — a 14-frame window through the public end-to-end call (select views, estimate, filter, fuse)
  C5  fusion stress — FuseDepthMaps alone on precomputed noisy maps with <= 12 neighbours by frame distance: identical to the
      CPU oracle at a reduced size, bookkeeping invariants at full size

The oracle cannot run these sizes in seconds, so full-size cases use size-independent properties (determinism, analytic ground
truth, claim bookkeeping); the reduced-size C5 case is compared with the oracle bit for bit.
"""
import numpy as np
import pytest

import common
import oracle_lib as O
from hcmvs_b200.synth import SynthScene

pytestmark = pytest.mark.gpu


from hcmvs_b200.synth import c5_maps, frame_neighbors  # the C5 recipe is shared with bench.py --config 5


# ------------------------------------------------------------------------------------------------ C3
def test_c3_full_resolution_view():
    syn, osc, gt, imgs, ok = common.make_scene(3, 1.0, 6)
    assert (syn.width, syn.height) == (6048, 4032)
    ref = 2
    assert ok[ref] and len(osc.match_views(ref)) >= 3
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        osc.init_depth_sparse(ref)
        d0, _, _, lo, hi = osc.get_depthmap(ref)
        runs = []
        for _ in range(2):
            ctx.init_depthmap(ref, d0, None, lo, hi)
            ctx.reset_timers()
            ctx.estimate_depthmap(ref, 0, seed=31)
            runs.append(ctx.get_depthmap(ref))
        t = ctx.timers()
        d, n, c = runs[0][:3]
        assert np.array_equal(d, runs[1][0]) and np.array_equal(c, runs[1][2])          # deterministic at 24 Mpx
        assert np.all(d[:7] == 0) and np.all(d[:, :7] == 0) and np.all(d[-7:] == 0) and np.all(d[:, -7:] == 0)
        valid = d > 0
        g = gt[ref][0]
        seen = (g > lo) & (g < hi)
        assert valid[seen].mean() > 0.85
        assert common.agreement(g, d, mask=valid & seen) >= 0.98                          # within 1 % of the analytic depth
        assert t["n_pixel_iters"] == 3 * (syn.width - 14) * (syn.height - 14)             # every inner pixel, every iteration
        print(f"\nC3 view: {valid.mean():.3f} valid, {common.agreement(g, d, mask=valid & seen):.4f} within 1 % of GT, "
              f"sweeps {t['n_pixel_iters'] / t['ms_sweeps'] / 1e3:.1f} Mpix*iter/s")
        # fusion bookkeeping at this size (probe cache of 24 M pixels x neighbours): GT-derived maps for every view
        rng = np.random.default_rng(300)
        for i in range(syn.n_views):
            g, gn = gt[i]
            inside = (g > lo * 0.5) & (g < hi * 2)
            dd = np.where(inside, g * (1 + 0.002 * rng.standard_normal(g.shape, dtype=np.float32)), 0).astype(np.float32)
            cc = np.where(inside, rng.uniform(0.5, 1.0, g.shape).astype(np.float32), 0).astype(np.float32)
            ctx.set_depthmap(i, dd, gn, cc, lo * 0.5, hi * 2)
        cloud = ctx.fuse_depthmaps(True, True)
        assert len(cloud["xyz"]) > 5_000_000 and cloud["n_views"].min() >= 2 and np.isfinite(cloud["xyz"]).all()
        assert int(cloud["n_views"].sum()) == len(cloud["views"])
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------ C4
def test_c4_video_window_end_to_end():
    from hcmvs_b200 import api, host
    syn = SynthScene(4, 1.0, 14)
    assert (syn.width, syn.height) == (1920, 1080)
    rendered = [syn.render(i) for i in range(syn.n_views)]
    imgs = [r[0] for r in rendered]
    params = dict(common.BENCH_PARAMS)
    ctx = api.Context(0, **params)
    try:
        hs = host.HostScene.from_synth(syn, imgs)
        st = hs.dense_reconstruction(ctx, seed=4, run_filter=True)
        cloud = hs.cloud()
        n = len(cloud["xyz"])
        assert n == st["n_points"] and n > 1_000_000
        # every frame in the middle of the window found neighbours in the 3..65 degree band and got a depth map
        mid = syn.n_views // 2
        d = ctx.get_depthmap(mid)[0]
        g = rendered[mid][1]
        kept = d > 0
        assert kept.mean() > 0.5
        assert np.mean(np.abs(d[kept] / g[kept] - 1) < 0.01) >= 0.98                      # filtered depths stay on the surface
        rng = np.random.default_rng(1)
        idx = rng.choice(n, 4000, replace=False)
        z = np.array([syn.height_at(x, y) for x, y, _ in cloud["xyz"][idx]])
        assert np.percentile(np.abs(cloud["xyz"][idx, 2] - z), 95) < 0.01 * syn.cfg.cam_distance
        assert cloud["n_views"].min() >= 2
        print(f"\nC4 window: {n} points from {syn.n_views} frames, scene {sum(st[k] for k in ('sec_upload', 'sec_estimate', 'sec_filter', 'sec_fuse')):.2f} s")
        hs.close()
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------ C5
def _c5_load(syn, scale_seed, with_oracle):
    from hcmvs_b200 import api
    ctx = api.Context(0, **common.BENCH_PARAMS)
    osc = O.OracleScene(**common.BENCH_PARAMS) if with_oracle else None
    maps = []
    for i in range(syn.n_views):
        bgr, d, n = syn.render(i)
        m = c5_maps((d, n), scale_seed + i)
        maps.append(m)
        gray = np.zeros(d.shape, np.float32)  # fusion reads no gray image
        ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], gray, bgr)
        ctx.set_depthmap(i, *m)
        if osc is not None:
            osc.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr=bgr)
    for i in range(syn.n_views):
        ids = frame_neighbors(i, syn.n_views)
        ctx.set_neighbors(i, ids, min(5, len(ids)))
        ctx.set_fuse_priority(i, len(ids))
        if osc is not None:
            osc.set_neighbors(i, ids, min(5, len(ids)))
            osc.set_depthmap(i, *maps[i])
    return ctx, osc, maps


def test_c5_fusion_stress_matches_oracle_reduced():
    syn = SynthScene(5, 0.2, 40)
    ctx, osc, maps = _c5_load(syn, 500, True)
    try:
        want = osc.fuse(True, True)
        got = ctx.fuse_depthmaps(True, True)
        assert len(want["xyz"]) > 20000
        for k in ("n_views", "views", "xyz", "weights", "colors", "normals"):
            assert np.array_equal(want[k], got[k]), k
        for i in range(syn.n_views):
            assert np.array_equal(osc.get_depthmap(i)[0], ctx.get_depthmap(i)[0]), f"view {i}"
        print(f"\nC5 reduced: {len(got['xyz'])} points from {syn.n_views} maps, {ctx.timers()['n_fuse_rounds']} rounds")
    finally:
        ctx.close(); osc.close()


def test_c5_fusion_stress_full_resolution_invariants():
    syn = SynthScene(5, 1.0, 24)
    ctx, _, maps = _c5_load(syn, 900, False)
    try:
        ctx.reset_timers()
        cloud = ctx.fuse_depthmaps(True, True)
        t = ctx.timers()
        n = len(cloud["xyz"])
        assert n > 2_000_000
        assert cloud["n_views"].min() >= 2 and cloud["n_views"].max() <= 13
        off = np.concatenate([[0], np.cumsum(cloud["n_views"])])
        inc = np.diff(cloud["views"].astype(np.int64)) > 0
        inc[off[1:-1] - 1] = True
        assert inc.all()                                                                  # sorted, unique view lists
        per_view = np.bincount(cloud["views"], minlength=syn.n_views)
        for i in range(syn.n_views):
            assert per_view[i] <= (maps[i][0] > 0).sum()                                  # a pixel is claimed at most once
        rng = np.random.default_rng(2)
        idx = rng.choice(n, 4000, replace=False)
        z = np.array([syn.height_at(x, y) for x, y, _ in cloud["xyz"][idx]])
        assert np.percentile(np.abs(cloud["xyz"][idx, 2] - z), 95) < 0.01 * syn.cfg.cam_distance   # outliers do not survive the vote
        # idempotence: re-loading the same maps gives the same cloud
        for i in range(syn.n_views):
            ctx.set_depthmap(i, *maps[i])
        again = ctx.fuse_depthmaps(True, True)
        assert np.array_equal(cloud["xyz"], again["xyz"]) and np.array_equal(cloud["views"], again["views"])
        print(f"\nC5 full size: {n} points from {syn.n_views} maps of 1920x1080 in {t['ms_fuse']:.1f} ms ({t['n_fuse_rounds']} rounds)")
    finally:
        ctx.close()
