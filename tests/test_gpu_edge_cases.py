"""Edge cases of the hot path through the C ABI (GPU): degenerate inputs, ragged sizes, the single-neighbour aggregator,
the generic patch-side path, and the error convention (status + message, never a crash or a silent fallback)."""
import numpy as np
import pytest

import common
from hcmvs_b200 import api

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    yield syn, osc, gt, imgs, ok, ctx
    ctx.close()


def test_single_matching_view_and_generic_patch_side(small):
    """N == 1 takes ScorePixel's single-score branch (DepthMap.cpp:1015); adapthalfwin 3 runs the runtime patch-side loop."""
    syn, osc, gt, imgs, ok, ctx = small
    ref = 2
    nb = osc.neighbors(ref, 1)
    for ahw, nmatch in ((5, 1), (3, 3), (6, 2)):
        osc.set_params(adapthalfwin=ahw); ctx.set_params(adapthalfwin=ahw)
        osc.set_neighbors(ref, nb["ids"], nmatch, nb["score"]); ctx.set_neighbors(ref, nb["ids"], nmatch, nb["score"])
        try:
            d, n = common.perturbed_hypotheses(gt[ref][0], gt[ref][1], syn.K[ref], seed=5 + ahw, depth_sigma=0.01, angle_deg=8.0)
            for smooth in (0, 1):
                want = osc.score_hypotheses(ref, d, n, smooth)
                got = ctx.score_hypotheses(ref, d, n, smooth)
                assert np.abs(want - got).max() <= 1e-4, (ahw, nmatch, smooth, np.abs(want - got).max())
        finally:
            osc.set_params(adapthalfwin=5); ctx.set_params(adapthalfwin=5)
            m = min(5, len(nb["ids"]))
            osc.set_neighbors(ref, nb["ids"], m, nb["score"]); ctx.set_neighbors(ref, nb["ids"], m, nb["score"])


def test_degenerate_hypotheses_score_like_the_oracle(small):
    """Zero / negative depth, a normal facing away, NaN: every lane must end in the reference's answer (thRobust paths), not a fault."""
    syn, osc, gt, imgs, ok, ctx = small
    ref = 1
    d, n = common.perturbed_hypotheses(gt[ref][0], gt[ref][1], syn.K[ref], seed=3, depth_sigma=0.0, angle_deg=0.0)
    d = d.copy(); n = n.copy()
    d[10:20] = 0.0                      # INVERT(0) -> huge homography
    d[20:30] = -1.0
    d[30:40] *= 100.0                   # patch far outside every neighbour
    n[40:50] *= -1.0                    # facing away
    n[50:60] = 0.0                      # zero normal
    want = osc.score_hypotheses(ref, d, n, 1)
    got = ctx.score_hypotheses(ref, d, n, 1)
    # a zero normal makes ComputeAngle 0/0: the reference's CLAMP keeps the NaN and so must the kernel (a NaN score loses every
    # `conf > nconf` test); everything else is a number and must agree
    assert np.array_equal(np.isnan(want), np.isnan(got))
    assert np.isnan(want[50:60]).mean() > 0.5            # the case is exercised
    fin = ~np.isnan(want)
    assert np.abs(want[fin] - got[fin]).max() <= 1e-4, np.abs(want[fin] - got[fin]).max()


def test_ragged_image_sizes_and_empty_maps_in_filter_and_fusion():
    """Views of different sizes, one neighbour with an all-zero depth map, one view without maps: filter and fusion == oracle."""
    import oracle_lib as O
    from hcmvs_b200.synth import SynthScene
    import test_gpu_configs as cfgs
    syn = SynthScene(5, 0.15, 8)
    osc = O.OracleScene(**common.BENCH_PARAMS)
    ctx = api.Context(0, **common.BENCH_PARAMS)
    try:
        maps = []
        for i in range(syn.n_views):
            bgr, d, n = syn.render(i)
            K = syn.K[i].copy()
            if i % 3 == 1:              # crop: a smaller image of the same camera (principal point unchanged)
                bgr, d, n = bgr[:-24, :-40].copy(), d[:-24, :-40].copy(), n[:-24, :-40].copy()
            m = list(cfgs.c5_maps((d, n), 700 + i))
            if i == 4:
                m[0] = np.zeros_like(m[0]); m[2] = np.zeros_like(m[2])       # an empty depth map
            maps.append(m)
            osc.add_image(K, syn.R[i], syn.Cc[i], bgr=bgr)
            ctx.set_view(i, K, syn.R[i], syn.Cc[i], np.zeros(d.shape, np.float32), bgr)
        for i in range(syn.n_views):
            ids = cfgs.frame_neighbors(i, syn.n_views)
            osc.set_neighbors(i, ids, min(5, len(ids))); ctx.set_neighbors(i, ids, min(5, len(ids)))
            ctx.set_fuse_priority(i, len(ids))
            if i != 6:                  # view 6 never gets maps
                osc.set_depthmap(i, *maps[i]); ctx.set_depthmap(i, *maps[i])
        for ref in (0, 2):
            ids = [int(v) for v in cfgs.frame_neighbors(ref, syn.n_views)]
            nbidx = [k for k, v in enumerate(ids) if v != 6][:6]
            want = osc.filter(ref, nbidx, True)
            got = ctx.filter_depthmap(ref, nbidx, True)
            assert np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])
        with pytest.raises(api.HcmvsError, match="no depth map"):
            ctx.filter_depthmap(0, [k for k, v in enumerate(cfgs.frame_neighbors(0, syn.n_views)) if v == 6] + [0, 1], True)
        want = osc.fuse(True, True)
        got = ctx.fuse_depthmaps(True, True)
        assert len(want["xyz"]) > 1000
        for k in ("n_views", "views", "xyz", "weights", "colors", "normals"):
            assert np.array_equal(want[k], got[k]), k
        assert 4 not in set(got["views"].tolist()) and 6 not in set(got["views"].tolist())
    finally:
        ctx.close(); osc.close()


def test_error_convention(small):
    syn, osc, gt, imgs, ok, ctx = small
    K, R, Cc = syn.K[0], syn.R[0], syn.Cc[0]
    with pytest.raises(api.HcmvsError, match="unsupported"):
        ctx.set_view(40, K, R, Cc, np.zeros((12, 12), np.float32), None)        # smaller than the 15x15 window
    Ks = K.copy(); Ks[1] = 0.5
    with pytest.raises(api.HcmvsError, match="skew"):
        ctx.set_view(40, Ks, R, Cc, np.zeros((64, 64), np.float32), None)
    with pytest.raises(api.HcmvsError, match="not set"):
        ctx.estimate_depthmap(39, 0, 1)
    other = api.Context(0, **common.BENCH_PARAMS)
    try:
        other.set_view(0, K, R, Cc, osc.gray(0), imgs[0])
        with pytest.raises(api.HcmvsError, match="no depth map"):
            other.estimate_depthmap(0, 0, 1)
        d = np.ones(osc.gray(0).shape, np.float32)
        with pytest.raises(api.HcmvsError, match="depth range"):
            other.init_depthmap(0, d, None, 2.0, 1.0)
        other.init_depthmap(0, d, None, 1.0, 2.0)
        with pytest.raises(api.HcmvsError, match="matching neighbours"):
            other.estimate_depthmap(0, 0, 1)
        with pytest.raises(api.HcmvsError, match="lists itself"):
            other.set_neighbors(0, np.array([0], np.uint32), 1)
        with pytest.raises(api.HcmvsError, match="no view with depth map and neighbours"):
            other.fuse_depthmaps(True, True)
        with pytest.raises(api.HcmvsError):
            other.set_params(adapthalfwin=9)
    finally:
        other.close()
    # the context is still usable after every refused call
    d, n = common.perturbed_hypotheses(gt[0][0], gt[0][1], syn.K[0], seed=1)
    assert np.isfinite(ctx.score_hypotheses(0, d, n, 0)).all()


def test_triangulated_init_edge_cases(small):
    """hcmvs_init_depthmap_triangles: a minimal mesh (3 points + nothing else), triangles partly or wholly outside the image, slivers
    thinner than the 1/16-pixel grid, and malformed input -> the numpy oracle's maps, or a clean error."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import triangulate_init as T
    syn, osc, gt, imgs, ok, ctx = small
    ref = 1
    h, w = imgs[ref].shape[:2]
    K = np.asarray(syn.K[ref], np.float64).ravel()
    cases = {
        "one triangle": (np.array([[10.2, 8.7, 5.0], [100.9, 20.1, 5.5], [40.3, 90.6, 6.0]]), np.array([[0, 1, 2]])),
        "clockwise (draws nothing)": (np.array([[10.2, 8.7, 5.0], [100.9, 20.1, 5.5], [40.3, 90.6, 6.0]]), np.array([[0, 2, 1]])),
        "partly outside": (np.array([[-50.0, -30.0, 5.0], [w + 40.0, 10.0, 6.0], [w / 2, h + 70.0, 5.5], [w / 2, h / 2, 5.2]]), np.array([[0, 1, 3], [1, 2, 3], [2, 0, 3]])),
        "sliver + overlap (last face wins)": (np.array([[20.0, 20.0, 5.0], [120.0, 20.01, 5.0], [70.0, 20.02, 5.0], [20.0, 100.0, 6.0], [120.0, 100.0, 6.0], [70.0, 10.0, 4.0]]),
                                              np.array([[0, 1, 2], [0, 1, 3], [1, 4, 3], [0, 5, 1], [0, 1, 4]])),
        "behind the camera (z <= 0 is skipped)": (np.array([[10.0, 10.0, -5.0], [150.0, 10.0, -5.0], [80.0, 110.0, -5.0]]), np.array([[0, 1, 2]])),
    }
    for name, (v, t) in cases.items():
        ctx.init_depthmap_triangles(ref, v, t, 1.0, 100.0)
        d, n = ctx.get_depthmap(ref)[:2]
        od, on = T.rasterize(v, t.astype(np.int64), K, w, h)
        assert np.array_equal(d, od) and np.array_equal(n, on), name
    assert (ctx.get_depthmap(ref)[0] == 0).all()                                 # the last case drew nothing
    v, t = cases["one triangle"]
    for bad_v, bad_t, lo, hi in ((v, np.array([[0, 1, 7]]), 1.0, 2.0), (v[:2], t, 1.0, 2.0), (v, t, 2.0, 1.0), (np.where(v > 90, np.inf, v), t, 1.0, 2.0)):
        with pytest.raises(api.HcmvsError):
            ctx.init_depthmap_triangles(ref, bad_v, bad_t, lo, hi)


def test_row_band_of_a_small_image_and_bad_ranges(small):
    """hcmvs_estimate_depthmap_rows where the halo is larger than the image (the band becomes the whole view) and invalid ranges."""
    syn, osc, gt, imgs, ok, ctx = small
    ref = 3
    osc.init_depth_sparse(ref)
    d0, _, _, lo, hi = osc.get_depthmap(ref)
    ctx.init_depthmap(ref, d0, None, lo, hi); ctx.estimate_depthmap(ref, 0, 13)
    want = ctx.get_depthmap(ref)
    h = d0.shape[0]
    got = [np.zeros_like(x) for x in want[:3]]
    for r0, r1 in ((0, 17), (17, h - 40), (h - 40, h)):
        ctx.init_depthmap(ref, d0, None, lo, hi); ctx.estimate_depthmap_rows(ref, r0, r1, 0, 13)
        band = ctx.get_depthmap(ref)
        for g, b in zip(got, band[:3]):
            g[r0:r1] = b[r0:r1]
    for g, x in zip(got, want[:3]):
        assert np.array_equal(g, x)
    for r0, r1 in ((5, 5), (-1, 10), (0, h + 1)):
        with pytest.raises(api.HcmvsError):
            ctx.estimate_depthmap_rows(ref, r0, r1, 0, 13)


def test_row_bands_in_a_later_outer_iteration():
    """it_external = 1 ('+'-shaped candidate set, reach 5, EndDepthMapTmp at the end): bands + halo still reproduce the full estimate."""
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok, nEstimationIters_external=2, propagatehalfwin=5, propagatestep=4)
    try:
        ref = 6
        osc.init_depth_sparse(ref)
        d0, _, _, lo, hi = osc.get_depthmap(ref)
        ctx.init_depthmap(ref, d0, None, lo, hi)
        ctx.estimate_depthmap(ref, 0, 21)
        first = ctx.get_depthmap(ref)                                 # state after outer iteration 0 (no PASS C yet)
        ctx.estimate_depthmap(ref, 1, 21)
        want = ctx.get_depthmap(ref)
        h = d0.shape[0]
        got = [np.zeros_like(x) for x in want[:3]]
        for r0, r1 in ((0, h // 3), (h // 3, 2 * h // 3), (2 * h // 3, h)):
            ctx.set_depthmap(ref, first[0], first[1], first[2], first[3], first[4])
            ctx.estimate_depthmap_rows(ref, r0, r1, 1, 21)
            band = ctx.get_depthmap(ref)
            for g, b in zip(got, band[:3]):
                g[r0:r1] = b[r0:r1]
        for g, x in zip(got, want[:3]):
            assert np.array_equal(g, x)
        assert (want[0] > 0).mean() > 0.5
    finally:
        ctx.close()
