"""The drop-in path with FILES on both sides (SURVEY §8b inputs/outputs): an MVSI project + image files on disk ->
Scene::LoadInterface -> DenseReconstruction on the GPU -> depthNNNN.dmap, scene_dense.ply, scene_dense.mvs — compared with the
in-memory path on the same scene."""
import os

import numpy as np
import pytest

import common

pytestmark = pytest.mark.gpu


def test_project_on_disk_end_to_end(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.5)
    hs0 = host.HostScene.from_synth(syn, imgs)                     # images named 00000.png ...
    for i, im in enumerate(imgs):
        assert cv2.imwrite(str(tmp_path / f"{i:05d}.png"), im)
    hs0.save_mvs(tmp_path / "scene.mvs")
    hs = host.HostScene.load_mvs(tmp_path / "scene.mvs")           # decodes the PNGs through the product's own reader
    assert hs.num_images() == syn.n_views
    for i in range(syn.n_views):
        assert np.array_equal(hs.image_bgr(i), imgs[i])
    clouds, depths = [], []
    for scene, dmap_dir in ((hs0, None), (hs, str(tmp_path))):
        ctx = api.Context(0, **common.BENCH_PARAMS)
        try:
            st = scene.dense_reconstruction(ctx, seed=7, run_filter=True, dmap_dir=dmap_dir)
            clouds.append(scene.cloud()); depths.append(ctx.get_depthmap(3)[0])
            assert st["n_points"] == len(clouds[-1]["xyz"]) > 50_000
            if dmap_dir:  # the fused cloud lives in the context's page-locked arena: write the outputs while the context is alive
                scene.save_ply(str(tmp_path / "scene_dense.ply"))
                scene.save_mvs(tmp_path / "scene_dense.mvs", dense=True)
        finally:
            ctx.close()
    # cameras of the loaded project differ from the in-memory ones by <= 1 ulp in K (normalise / scale-back of the project format):
    # the two runs agree to that level
    n0, n1 = len(clouds[0]["xyz"]), len(clouds[1]["xyz"])
    assert abs(n0 - n1) <= 0.002 * n0
    assert common.agreement(depths[0], depths[1], th=1e-4) >= 0.999
    # outputs: raw depth-data files written BEFORE the filter (ExportDepthDataRaw), the PLY and the dense project
    for i in range(syn.n_views):
        if not ok[i]:
            continue
        back = host.read_dmap(str(tmp_path / f"depth{i:04d}.dmap"))
        assert back["depth"].shape == imgs[i].shape[:2] and back["normal"] is not None and back["conf"] is not None
        assert back["ids"][0] == i and (back["depth"] > 0).mean() > 0.5
        valid = back["depth"] > 0
        assert np.mean(np.abs(back["depth"][valid] / gt[i][0][valid] - 1) < 0.01) > 0.9
    raw = open(tmp_path / "scene_dense.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"element vertex %d\n" % n1 in head and b"property float32 nx" in head and len(body) == n1 * 27
    dense = host.HostScene.load_mvs(tmp_path / "scene_dense.mvs", load_images=False)
    xyz, off, ids, wts = dense.sparse()
    assert len(xyz) == n1 and np.array_equal(xyz, clouds[1]["xyz"]) and np.array_equal(np.diff(off), clouds[1]["n_views"])
    assert np.array_equal(wts, clouds[1]["weights"])
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "mvsi_ref_tool")
    if os.path.exists(tool):  # the reference's own loader accepts what the GPU path wrote
        import subprocess
        assert subprocess.run([tool, "to-flat", str(tmp_path / "scene_dense.mvs"), str(tmp_path / "dense.flat")], capture_output=True, text=True, check=True).stdout.strip() == "5"
    hs0.close(); hs.close(); dense.close()


def test_async_download_slots_match_synchronous_readback():
    """hcmvs_download_depthmap_begin / _wait (page-locked slots, copy on its own stream) return exactly what hcmvs_get_depthmap does,
    also when later work has already overwritten the view's maps, and misuse is an error, not a crash."""
    from hcmvs_b200 import api
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = common.make_context(syn, osc, imgs, ok)
    try:
        want = {}
        for k, ref in enumerate((0, 2, 5, 7, 8)):
            osc.init_depth_sparse(ref)
            d0, _, _, lo, hi = osc.get_depthmap(ref)
            ctx.init_depthmap(ref, d0, None, lo, hi)
            ctx.estimate_depthmap(ref, 0, seed=9)
            if k < 4:
                ctx.download_begin(ref, k)                                   # queued behind the estimation, returns at once
            want[ref] = None
        ctx.set_depthmap(0, np.zeros_like(d0), None, None, lo, hi)           # overwrite view 0 AFTER its read-back was queued
        for k, ref in enumerate((0, 2, 5, 7)):
            got = ctx.download_wait(k)
            if ref != 0:
                ref_d, ref_n, ref_c, rlo, rhi = ctx.get_depthmap(ref)
                assert np.array_equal(got[0], ref_d) and np.array_equal(got[1], ref_n) and np.array_equal(got[2], ref_c) and (got[3], got[4]) == (rlo, rhi)
            else:
                assert (got[0] > 0).mean() > 0.5                              # the estimate, not the zeros written later
        ctx.download_begin(8, 0)                                              # a slot is reusable once waited for
        assert np.array_equal(ctx.download_wait(0)[0], ctx.get_depthmap(8)[0])
        with pytest.raises(api.HcmvsError):
            ctx.download_wait(1)                                              # nothing pending
        with pytest.raises(api.HcmvsError):
            ctx.download_begin(0, 99)
    finally:
        ctx.close()


def test_host_pointcloud_filter_removal_order():
    """Scene::PointCloudFilter through the host mirror: device votes, then RFOREACH + RemovePoint (the LAST point moves into the hole,
    PointCloud.cpp:54-69 / cList::RemoveAt) — the surviving cloud, in order, equals a Python simulation on the same votes."""
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    ctx = api.Context(0, **common.BENCH_PARAMS)
    hs = host.HostScene.from_synth(syn, imgs)
    try:
        hs.dense_reconstruction(ctx, seed=3, run_filter=True)
        before = hs.cloud()
        vis, _ = ctx.pointcloud_filter()                     # the same cloud, resident on the device
        th = -1                                               # DensifyPointCloud --filter-point-cloud -1
        order = list(range(len(vis)))
        size = len(order)
        for i in range(len(vis) - 1, -1, -1):
            if vis[i] <= th:
                if i + 1 != size:
                    order[i] = order[size - 1]
                size -= 1
        order = np.asarray(order[:size])
        removed = hs.pointcloud_filter(ctx, th)
        after = hs.cloud()
        assert removed == len(vis) - size and 0 < removed < len(vis) // 4
        assert np.array_equal(after["xyz"], before["xyz"][order]) and np.array_equal(after["n_views"], before["n_views"][order])
        assert np.array_equal(after["colors"], before["colors"][order]) and np.array_equal(after["normals"], before["normals"][order])
        off = np.concatenate([[0], np.cumsum(before["n_views"])])
        want_views = np.concatenate([before["views"][off[i]:off[i + 1]] for i in order])
        assert np.array_equal(after["views"], want_views)
    finally:
        hs.close(); ctx.close()


@pytest.mark.gpu
def test_reused_context_repeats_the_scene_bit_for_bit():
    """hcmvs_begin_scene: on a reused context the initial-map uploads of later views only wait for their own view's previous use, so they
    overlap the estimation of the views before them. Three scenes in a row on ONE context (the second and third take that path: every
    view already holds maps) must give the same cloud as the first, and as a fresh context, point for point."""
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.5)
    clouds = []
    ctx = api.Context(0, **common.BENCH_PARAMS)
    hs = host.HostScene.from_synth(syn, imgs)
    try:
        for rep in range(3):
            hs.dense_reconstruction(ctx, seed=11, run_filter=True)
            clouds.append(hs.cloud())
    finally:
        hs.close(); ctx.close()
    ctx = api.Context(0, **common.BENCH_PARAMS)
    hs = host.HostScene.from_synth(syn, imgs)
    try:
        hs.dense_reconstruction(ctx, seed=11, run_filter=True)
        clouds.append(hs.cloud())
    finally:
        hs.close(); ctx.close()
    assert len(clouds[0]["xyz"]) > 10000
    for c in clouds[1:]:
        for k in ("xyz", "n_views", "colors", "normals"):
            assert np.array_equal(c[k], clouds[0][k]), k
