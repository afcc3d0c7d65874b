"""Golden-vector tests. CPU: the oracle still reproduces the committed fixture (tests/golden/make_golden.py).
GPU: the CUDA path reproduces the same fixture through the C ABI, independently of the oracle build on the box."""
import os
import sys

import numpy as np
import pytest

import common

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "c1_quarter.npz"))


def test_oracle_reproduces_golden():
    cur = make_golden.build()
    assert set(cur) == set(G.files)
    exact = ["image_sha", "nb_ids", "nb_all_count", "init_depth", "init_range", "gramap", "fuse_views_sha", "fuse_count",
             "filter_strict_depth", "fuse_colors_head"] + [f"hyp{k}_score0" for k in range(3)]
    for k in exact:
        assert np.array_equal(cur[k], G[k]), k
    for k in range(3):  # smoothness uses libm expf / acosf: last-bit drift allowed
        assert np.abs(cur[f"hyp{k}_score1"] - G[f"hyp{k}_score1"]).max() < 2e-6
    assert np.allclose(cur["nb_score"], G["nb_score"], rtol=1e-6)
    for k in ("filter_adjust_depth", "filter_adjust_conf", "fuse_xyz_head"):
        assert np.array_equal(cur[k], G[k]), k
    # PatchMatch results depend on libm sin/cos in the random draws: statistical agreement with the stored run
    assert common.agreement(G["raster_depth"], cur["raster_depth"]) > 0.97
    assert common.agreement(G["redblack_depth"], cur["redblack_depth"]) > 0.97


@pytest.mark.gpu
def test_gpu_reproduces_golden():
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    ref = 0
    for k in range(3):
        d, n = G[f"hyp{k}_depth"], G[f"hyp{k}_normal"]
        assert np.abs(ctx.score_hypotheses(ref, d, n, 0) - G[f"hyp{k}_score0"]).max() <= 1e-4
        assert np.abs(ctx.score_hypotheses(ref, d, n, 1) - G[f"hyp{k}_score1"]).max() <= 1e-4
    lo, hi = float(G["init_range"][0]), float(G["init_range"][1])
    ctx.init_depthmap(ref, G["init_depth"], None, lo, hi)
    assert np.array_equal(ctx.gradient_map(ref), G["gramap"])
    ctx.estimate_depthmap(ref, 0, seed=3)
    gd = ctx.get_depthmap(ref)[0]
    assert common.agreement(G["redblack_depth"], gd) >= 0.99           # same algorithm, same counter RNG
    both = (G["raster_depth"] > 0) & (G["redblack_depth"] > 0) & (np.abs(G["raster_depth"] - G["redblack_depth"]) / np.maximum(G["raster_depth"], 1e-9) < 0.01)
    assert common.agreement(G["raster_depth"], gd, mask=both) >= 0.98   # vs the reference's raster sweep, where that is well defined
    for i in range(syn.n_views):
        d = G[f"map{i}_depth"]
        ctx.set_depthmap(i, d, gt[i][1], G[f"map{i}_conf"], float(gt[i][0][gt[i][0] > 0].min() * 0.5), float(gt[i][0].max() * 2))
    nbf = list(range(min(8, len(G["nb_ids"]))))
    fd, fc = ctx.filter_depthmap(ref, nbf, True)
    assert np.array_equal(fd, G["filter_adjust_depth"]) and np.array_equal(fc, G["filter_adjust_conf"])
    fd, _ = ctx.filter_depthmap(ref, nbf, False)
    assert np.array_equal(fd, G["filter_strict_depth"])
    cloud = ctx.fuse_depthmaps(True, True)
    assert [len(cloud["xyz"]), len(cloud["views"])] == list(G["fuse_count"])
    assert make_golden.sha(cloud["views"]) == G["fuse_views_sha"][0] and make_golden.sha(cloud["xyz"]) == G["fuse_xyz_sha"][0]
    assert np.array_equal(cloud["colors"][:2000], G["fuse_colors_head"])
    ctx.close()


# ------------------------------------------------------------------------------------------------ second fixture: later-added paths
import make_golden_ext  # noqa: E402

GX = np.load(os.path.join(HERE, "golden", "c1_quarter_ext.npz"))


def test_oracle_reproduces_extended_golden():
    cur = make_golden_ext.build()
    assert set(cur) == set(GX.files)
    exact = ["resize_dst1", "resize_dst3_sha", "coarse_resized_depth", "coarse_range", "coarse_init_depth", "scaled_dn_gray_sha", "scaled_up_gray_sha",
             "scaled_dn_K", "scaled_up_K", "scaled_dn_score0", "scaled_up_score0", "vs_views"]
    for k in exact:
        assert np.array_equal(cur[k], GX[k]), k
    # PatchMatch runs draw through libm sin/cos: statistical agreement with the stored run
    assert common.agreement(GX["coarse_redblack_depth"], cur["coarse_redblack_depth"]) > 0.97
    assert common.agreement(GX["vs_redblack_depth"], cur["vs_redblack_depth"]) > 0.97


@pytest.mark.gpu
def test_gpu_reproduces_extended_golden():
    from hcmvs_b200 import host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        # coarse-level hand-off: resize, widened range, last-iteration hypothesis
        ref = 3
        ctx.set_params(nEstimationIters=2, nEstimationIters_external=1)
        osc.init_depth_sparse(ref)
        d0, _, _, lo, hi = osc.get_depthmap(ref)
        assert np.array_equal(d0, GX["coarse_init_depth"])
        ctx.init_depthmap(ref, GX["coarse_init_depth"], None, lo, hi)
        ctx.set_coarse_estimate(ref, GX["coarse_depth"], GX["coarse_normal"])
        assert np.array_equal(ctx.get_coarse_estimate(ref)[0], GX["coarse_resized_depth"])
        assert np.array_equal(np.array(ctx.get_depthmap(ref)[3:], np.float32), GX["coarse_range"])
        ctx.estimate_depthmap(ref, 0, seed=33)
        assert common.agreement(GX["coarse_redblack_depth"], ctx.get_depthmap(ref)[0]) >= 0.99
        ctx.set_coarse_estimate(ref, None, None)
        ctx.set_params(nEstimationIters=3, nEstimationIters_external=1)
        # rescaled matching view
        ref, slot = 5, 1
        nb = int(osc.match_views(ref)[slot])
        for tag, scale in (("dn", 0.8), ("up", 1.25)):
            g, Ks = host.scale_image(osc.gray(nb), scale, syn.K[nb])
            assert make_golden_ext.sha(g) == GX[f"scaled_{tag}_gray_sha"][0] and np.array_equal(Ks, GX[f"scaled_{tag}_K"])
            ctx.set_neighbor_image(ref, slot, Ks, g)
            got = ctx.score_hypotheses(ref, GX["scaled_hyp_depth"], GX["scaled_hyp_normal"], 0)
            assert np.abs(got - GX[f"scaled_{tag}_score0"]).max() <= 1e-4
            ctx.set_neighbor_image(ref, slot, None, None)
        # viewspread from stored neighbour maps
        ref = 4
        over = dict(nEstimationIters=1, nEstimationIters_external=2, propagatehalfwin=5, propagatestep=4, viewspread=1)
        ctx.set_params(**over)
        for v in GX["vs_views"]:
            v = int(v)
            r = GX[f"vs_view{v}_range"]
            if v == ref:
                ctx.init_depthmap(v, GX[f"vs_view{v}_depth"], GX[f"vs_view{v}_normal"], float(r[0]), float(r[1]))   # builds the gradient map
            ctx.set_depthmap(v, GX[f"vs_view{v}_depth"], GX[f"vs_view{v}_normal"], GX[f"vs_view{v}_conf"], float(r[0]), float(r[1]))
        ctx.snapshot_maps()
        ctx.estimate_depthmap(ref, 1, seed=61)
        assert common.agreement(GX["vs_redblack_depth"], ctx.get_depthmap(ref)[0]) >= 0.99
    finally:
        ctx.close()


# ------------------------------------------------------------------------------------------------ default init / cloud post-processing
TRI = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_quarter_triinit.npz")


def test_host_triangulation_reproduces_golden(built):
    """TriangulatePointsDelaunay in the product host code == the committed oracle fixture (generated with scipy's Qhull): faces and
    projected vertices exactly, corner depths to f32 rounding. Needs neither scipy nor a GPU."""
    import common
    from hcmvs_b200 import api, host
    g = np.load(TRI)
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    hs = host.HostScene.from_synth(syn, imgs)
    for ref in (0, 7):
        assert hs.select_views(api.default_params(), ref) > 0
        v, t, lo, hi = hs.triangulate_init(ref)
        assert np.array_equal(t, g[f"v{ref}_faces"]) and np.array_equal(v[:-4], g[f"v{ref}_vertices"][:-4])
        assert np.allclose(v[-4:], g[f"v{ref}_vertices"][-4:], rtol=2e-6, atol=0) and np.array_equal(np.array([lo, hi], np.float32), g[f"v{ref}_range"])
    hs.close()


@pytest.mark.gpu
def test_gpu_default_init_and_cloud_postprocessing_reproduce_golden():
    """Device rasteriser, EstimatePointColors and PointCloudFilter votes against the committed oracle outputs, bit for bit."""
    import common
    g = np.load(TRI)
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        for ref in (0, 7):
            lo, hi = g[f"v{ref}_range"]
            ctx.init_depthmap_triangles(ref, g[f"v{ref}_vertices"], g[f"v{ref}_faces"], float(lo) * 0.9, float(hi) * 1.1)
            d, n = ctx.get_depthmap(ref)[:2]
            assert np.array_equal(d, g[f"v{ref}_depth"]) and np.array_equal(n, g[f"v{ref}_normal"])
        pts, off, views = g["cloud_points"], g["cloud_offsets"], g["cloud_views"]
        assert np.array_equal(ctx.estimate_point_colors(pts, off, views), g["cloud_colors"])
        assert np.array_equal(ctx.pointcloud_filter(pts, off, views)[0], g["cloud_visibility"])
    finally:
        ctx.close()
