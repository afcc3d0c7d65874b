"""GPU parity of FilterDepthMap and FuseDepthMaps against the CPU oracle: integer / index work is bit-exact,
and because the GPU fusion reproduces the CPU's sequential claim order the whole cloud must be identical."""
import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def noisy_maps(gt, seed, outliers=0.02, sigma=0.002):
    """C5-style maps: GT depth * (1 + N(0, sigma)), a few outliers, rotated normals, conf ~ U(0.5, 1)."""
    rng = np.random.default_rng(seed)
    d, n = gt
    h, w = d.shape
    valid = d > 0
    depth = (d * (1 + sigma * rng.standard_normal((h, w)))).astype(np.float32)
    out = rng.uniform(size=(h, w)) < outliers
    depth[out] = rng.uniform(d[valid].min() * 0.8, d[valid].max() * 1.2, int(out.sum())).astype(np.float32)
    depth[rng.uniform(size=(h, w)) < 0.05] = 0          # holes
    depth[~valid] = 0
    nn = n.astype(np.float64) + 0.05 * rng.standard_normal((h, w, 3))
    nn /= np.maximum(np.linalg.norm(nn, axis=2, keepdims=True), 1e-9)
    conf = rng.uniform(0.5, 1.0, (h, w)).astype(np.float32)
    conf[depth == 0] = 0
    return depth, nn.astype(np.float32), conf


@pytest.fixture(scope="module")
def loaded():
    syn, osc, gt, imgs, ok = common.make_scene(2, 0.2, 12)
    ctx = common.make_context(syn, osc, imgs, ok)
    maps = [noisy_maps(gt[i], 100 + i) for i in range(syn.n_views)]
    yield syn, osc, ctx, maps, ok
    ctx.close()


def load_maps(syn, osc, ctx, maps):
    for i in range(syn.n_views):
        d, n, c = maps[i]
        lo, hi = float(d[d > 0].min() * 0.5), float(d.max() * 2.0)
        osc.set_depthmap(i, d, n, c, lo, hi)
        ctx.set_depthmap(i, d, n, c, lo, hi)


@pytest.mark.parametrize("adjust", [True, False])
def test_filter_depthmap_bit_exact(loaded, adjust):
    syn, osc, ctx, maps, ok = loaded
    load_maps(syn, osc, ctx, maps)
    for ref in (0, 5):
        nb = list(range(min(8, len(osc.neighbors(ref, 1)["ids"]))))
        want = osc.filter(ref, nb, adjust)
        got = ctx.filter_depthmap(ref, nb, adjust)
        assert want is not None
        assert np.array_equal(want[0] == 0, got[0] == 0), "kept/discarded pixel sets differ"
        assert np.array_equal(want[0], got[0]), np.abs(want[0] - got[0]).max()
        assert np.array_equal(want[1], got[1]), np.abs(want[1] - got[1]).max()
        assert (want[0] > 0).sum() > 1000  # the case is not vacuous


@pytest.mark.parametrize("min_views_fuse", [2, 3, 1])
def test_fuse_depthmaps_identical_cloud(loaded, min_views_fuse):
    """nMinViewsFuse 2 is the shipped value; 3 and 1 exercise the "will / will not be emitted" early decisions of k_fuse_view."""
    syn, osc, ctx, maps, ok = loaded
    load_maps(syn, osc, ctx, maps)
    osc.set_params(nMinViewsFuse=min_views_fuse); ctx.set_params(nMinViewsFuse=min_views_fuse)
    try:
        want = osc.fuse(True, True)
        got = ctx.fuse_depthmaps(True, True)
    finally:
        osc.set_params(nMinViewsFuse=2); ctx.set_params(nMinViewsFuse=2)
    assert len(want["xyz"]) > 10000
    assert len(want["xyz"]) == len(got["xyz"]), (len(want["xyz"]), len(got["xyz"]))
    assert np.array_equal(want["n_views"], got["n_views"])
    assert np.array_equal(want["views"], got["views"])          # claim bookkeeping: exact
    assert np.array_equal(want["xyz"], got["xyz"])               # f64 accumulation, un-fused: exact
    assert np.array_equal(want["weights"], got["weights"])
    assert np.array_equal(want["colors"], got["colors"])
    assert np.array_equal(want["normals"], got["normals"])
    # fusion zeroes occluded depths in place (SceneDensify.cpp:3447-3449): the maps must end up identical too
    for i in range(syn.n_views):
        assert np.array_equal(osc.get_depthmap(i)[0], ctx.get_depthmap(i)[0]), f"view {i}"
    # the fork's RemoveSmallSegments product (SceneDensify.cpp:2228-2260): per view, the estimate where the pixel joined a fused point.
    # Every claimed pixel is exactly one (point, view) entry of the cloud, so the per-view counts must equal the view-list histogram.
    per_view = np.bincount(want["views"], minlength=syn.n_views)
    for i in range(syn.n_views):
        df, nf = ctx.fused_support(i)
        after = ctx.get_depthmap(i)
        sup = df > 0
        assert int(sup.sum()) == int(per_view[i]), (i, int(sup.sum()), int(per_view[i]))
        assert np.array_equal(df[sup], after[0][sup]) and np.array_equal(nf[sup], after[1][sup]) and np.all(nf[~sup] == 0)
    print(f"\nfused {len(got['xyz'])} points in {ctx.timers()['n_fuse_rounds']} reserve/commit rounds")


def test_fuse_is_idempotent_on_reloaded_maps(loaded):
    """Size-independent property: re-loading the same maps and fusing again gives the same cloud (no state leaks)."""
    syn, osc, ctx, maps, ok = loaded
    load_maps(syn, osc, ctx, maps)
    a = ctx.fuse_depthmaps(True, True)
    load_maps(syn, osc, ctx, maps)
    b = ctx.fuse_depthmaps(True, True)
    for k in ("xyz", "views", "weights", "colors", "normals"):
        assert np.array_equal(a[k], b[k])


def test_estimate_point_colors_matches_oracle():
    """MVS::EstimatePointColors (--estimate-colors 1): nearest-view selection and the uint8-truncating bilinear sample, bit for bit."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import point_colors as PC
    from test_triangulate_init import _compose_p
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        rng = np.random.default_rng(11)
        n = 3000
        xy = rng.uniform(-2.2, 2.2, (n, 2))
        pts = np.stack([xy[:, 0], xy[:, 1], 0.05 * xy[:, 0] + 0.03 * xy[:, 1] + rng.normal(0, 0.01, n)], 1).astype(np.float32)   # around the C1 plane, some outside every image
        counts = rng.integers(0, 5, n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
        views = np.concatenate([rng.choice(syn.n_views, c, replace=False) for c in counts] + [np.zeros(0, np.int64)]).astype(np.uint32)
        got = ctx.estimate_point_colors(pts, off, views)
        P_list = [_compose_p(syn.K[v], syn.R[v], syn.Cc[v]) for v in range(syn.n_views)]
        want = PC.estimate_point_colors(P_list, imgs, pts, off, views)
        assert np.array_equal(got, want)
        white = (got == 255).all(axis=1)
        assert 0.02 < white.mean() < 0.9 and (counts[~white] > 0).all()
        # the device-resident fused cloud: recolouring must reproduce the same function of its own points / view lists
        for i in range(syn.n_views):
            if ok[i]:
                osc.init_depth_sparse(i)
                d0, _, _, lo, hi = osc.get_depthmap(i)
                ctx.init_depthmap(i, d0, None, lo, hi); ctx.estimate_depthmap(i, 0, seed=2)
        cloud = ctx.fuse_depthmaps(color=False, normal=True)
        recol = ctx.estimate_point_colors()
        off2 = np.concatenate([[0], np.cumsum(cloud["n_views"])]).astype(np.uint32)
        sel = rng.choice(len(cloud["xyz"]), 2000, replace=False)
        sub_off = np.concatenate([[0], np.cumsum(cloud["n_views"][sel])]).astype(np.uint32)
        sub_views = np.concatenate([cloud["views"][off2[k]:off2[k + 1]] for k in sel]).astype(np.uint32)
        want2 = PC.estimate_point_colors(P_list, imgs, cloud["xyz"][sel], sub_off, sub_views)
        got2 = recol[sel]
        assert np.array_equal(got2, want2)
    finally:
        ctx.close()


def test_pointcloud_filter_votes_match_oracle():
    """Scene::PointCloudFilter: the binned GPU vote == the brute-force numpy restatement, integer for integer — on a synthetic cloud with
    a second, occluded / occluding layer (so that votes exist), points listed for views they do not project into (fallback path),
    and on a real fused cloud."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import cloud_filter as CF
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    cams = {v: dict(C=np.asarray(syn.Cc[v], np.float64), K=np.asarray(syn.K[v], np.float64).ravel(), width=imgs[v].shape[1]) for v in range(syn.n_views)}
    try:
        rng = np.random.default_rng(21)
        n = 5000
        xy = rng.uniform(-0.8, 0.8, (n, 2))
        z = 0.05 * xy[:, 0] + 0.03 * xy[:, 1]
        layer = rng.uniform(size=n)
        z = np.where(layer < 0.15, z + rng.uniform(0.2, 1.5, n), np.where(layer < 0.3, z - rng.uniform(0.2, 1.5, n), z + rng.normal(0, 0.002, n)))
        pts = np.stack([xy[:, 0], xy[:, 1], z], 1).astype(np.float32)
        pts[:20] *= 40                                                 # far outside every image: listed views they do not project into
        counts = rng.integers(1, 5, n)
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
        views = np.concatenate([np.sort(rng.choice(syn.n_views, c, replace=False)) for c in counts]).astype(np.uint32)
        got, stats = ctx.pointcloud_filter(pts, off, views)
        want = CF.visibility(cams, pts, off, views)
        assert np.array_equal(got, want)
        assert (got > 0).sum() > 5 and (got < 0).sum() > 20 and stats[0] >= 20 and stats[2] > 0
        # the fused cloud resident on the device
        for i in range(syn.n_views):
            if ok[i]:
                osc.init_depth_sparse(i)
                d0, _, _, lo, hi = osc.get_depthmap(i)
                ctx.init_depthmap(i, d0, None, lo, hi); ctx.estimate_depthmap(i, 0, seed=4)
        cloud = ctx.fuse_depthmaps(color=False, normal=False)
        vis, stats = ctx.pointcloud_filter()
        m = len(cloud["xyz"])
        assert len(vis) == m
        sel = np.sort(rng.choice(m, 1500, replace=False))             # the oracle is O(n^2): check against a brute-force vote restricted to sources in sel
        off2 = np.concatenate([[0], np.cumsum(cloud["n_views"])]).astype(np.uint32)
        sub = cloud["xyz"][:4000]; sub_off = off2[:4001]; sub_views = cloud["views"][:off2[4000]]
        got_sub, _ = ctx.pointcloud_filter(sub, sub_off, sub_views)
        assert np.array_equal(got_sub, CF.visibility(cams, sub, sub_off, sub_views))
        print(f"\npoint-cloud filter: {m} points, {int(stats[2])} candidate tests ({stats[2] / max(off2[-1], 1):.0f} per cone), {(vis <= -1).sum()} points at or below -1")
    finally:
        ctx.close()


def test_estimate_point_normals_matches_oracle():
    """MVS::EstimatePointNormals: grid k-NN + Jacobi PCA on the device vs cKDTree + eigh. The normal is a function of the neighbour
    set; wherever that set is unambiguous and the least-variance direction is well separated the two agree to 1e-3 degree."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import point_normals as PN
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        rng = np.random.default_rng(33)
        clouds = []
        xy = rng.uniform(-1.5, 1.5, (20000, 2))                                                      # a noisy slanted plane (the C1 surface)
        clouds.append(np.stack([xy[:, 0], xy[:, 1], 0.05 * xy[:, 0] + 0.03 * xy[:, 1] + rng.normal(0, 0.003, len(xy))], 1))
        clouds.append(rng.normal(0, 1, (6000, 3)) * [3.0, 1.0, 0.2])                                 # a volumetric blob: uneven density, no surface
        clouds.append(np.concatenate([clouds[0][:300] * 50, rng.uniform(-1, 1, (12, 3)) * 1e-3]))   # far apart sparse points + a tiny cluster
        # isolated outliers tens to thousands of point spacings away from a dense surface: the fine grid's capped search cannot finish them,
        # the coarser levels of the grid hierarchy must (exact k-NN, like CGAL's unbounded search)
        far = np.concatenate([rng.uniform(-1, 1, (30, 3)) * [40, 40, 40], rng.uniform(-1, 1, (10, 3)) * [600, 600, 600]])
        clouds.append(np.concatenate([clouds[0], far]))
        cams = np.asarray(syn.Cc, np.float64).reshape(-1, 3)
        for ci, pts in enumerate(clouds):
            pts = pts.astype(np.float32)
            n = len(pts)
            counts = rng.integers(1, 4, n)
            off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
            views = rng.integers(0, syn.n_views, int(off[-1])).astype(np.uint32)
            for k in (16, 5):
                got = ctx.estimate_point_normals(pts, off, views, num_neighbors=k)
                want, gap = PN.estimate_point_normals(pts, off, views, cams, num_neighbors=k)
                assert np.allclose(np.linalg.norm(got, axis=1), 1, atol=1e-5)
                ang = np.degrees(np.arccos(np.clip((got * want).sum(axis=1), -1, 1)))
                well = gap > 1e-3
                assert well.mean() > 0.9 and np.mean(ang[well] < 1e-3 * 57.3 + 0.02) > 0.999, (ci, k, float(np.mean(ang[well] < 0.08)), float(ang[well].max()))
                first = views[off[:-1].astype(np.int64)]
                assert np.all((got * (cams[first].astype(np.float32) - pts)).sum(axis=1) >= -1e-6)    # faces its first view
        with pytest.raises(Exception):
            ctx.estimate_point_normals(clouds[0].astype(np.float32), off, views, num_neighbors=64)
    finally:
        ctx.close()
