"""CPU tests of the oracle (the checker): helpers against closed forms / OpenCV, the restated stages against
analytic ground truth of the synthetic scenes. The reference ships no golden vectors (parity unpinned), so
these are what pins the restatement."""
import ctypes as C

import numpy as np
import pytest

import common
import oracle_lib as O


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    L = O.lib()
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        c = np.array(ctr, np.uint32); k = np.array(key, np.uint32); out = np.zeros(4, np.uint32)
        L.orc_philox(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        assert tuple(int(x) for x in out) == want


def test_zigzag_order_matches_reference_code():
    """MapMatrix2ZigzagIdx, DepthMap.cpp:354-381, traced by hand for a 3x3 block: anti-diagonals, each walked from
    top-right to bottom-left -> 1 | 2 4 | 3 5 7 | 6 8 | 9. (The comment above the function, :350-353, shows the
    alternating zig-zag 1 2 4 7 5 3 6 8 9, which is NOT what the code does; the code is the truth.)"""
    L = O.lib()
    out = np.zeros((9, 2), np.uint16)
    n = L.orc_zigzag(3, 3, 16, out.ctypes.data_as(C.c_void_p))
    assert n == 9
    visited = [int(y) * 3 + int(x) + 1 for x, y in out]
    assert visited == [1, 2, 4, 3, 5, 7, 6, 8, 9]
    # any size: a permutation of all pixels, stripes of rawStride rows
    for (w, h, stride) in ((37, 23, 8), (64, 130, 64), (5, 1, 16)):
        out = np.zeros((w * h, 2), np.uint16)
        assert L.orc_zigzag(w, h, stride, out.ctypes.data_as(C.c_void_p)) == w * h
        lin = out[:, 1].astype(np.int64) * w + out[:, 0]
        assert len(np.unique(lin)) == w * h


def test_median_and_gradient_map_match_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    L = O.lib()
    img = rng.uniform(0, 10, (61, 83)).astype(np.float32)
    img[rng.uniform(size=img.shape) < 0.5] = 0
    mine = img.copy()
    L.orc_median3(mine.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0])
    assert np.array_equal(mine, cv2.medianBlur(img, 3))
    bgr = rng.integers(0, 256, (57, 71, 3), dtype=np.uint8)
    gra = np.zeros(bgr.shape[:2], np.uint8)
    L.orc_gramap(bgr.ctypes.data_as(C.c_void_p), bgr.shape[1], bgr.shape[0], gra.ctypes.data_as(C.c_void_p))
    src = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    gx = cv2.convertScaleAbs(cv2.Sobel(src, cv2.CV_16S, 1, 0, ksize=3))
    gy = cv2.convertScaleAbs(cv2.Sobel(src, cv2.CV_16S, 0, 1, ksize=3))
    assert np.array_equal(gra, cv2.addWeighted(gx, 0.5, gy, 0.5, 0))


def test_togray_formula():
    rng = np.random.default_rng(1)
    bgr = rng.integers(0, 256, (9, 11, 3), dtype=np.uint8)
    out = np.zeros(bgr.shape[:2], np.float32)
    O.lib().orc_togray(bgr.ctypes.data_as(C.c_void_p), 11, 9, out.ctypes.data_as(C.c_void_p))
    f = bgr.astype(np.float32) * np.float32(1.0 / 255.0)
    want = (np.float32(0.114) * f[..., 0] + np.float32(0.587) * f[..., 1]) + np.float32(0.299) * f[..., 2]
    assert np.array_equal(out, want)


def test_bilinear_sample_and_direction_encoding():
    L = O.lib()
    img = np.arange(20, dtype=np.float32).reshape(4, 5)  # I(x,y) = 5y + x is reproduced exactly by bilinear interpolation
    for x, y in ((1.25, 1.5), (2.0, 2.0), (3.999, 1.001)):
        got = L.orc_sample(img.ctypes.data_as(C.c_void_p), 5, 4, x, y)
        assert abs(got - (5 * y + x)) < 1e-5
    for a, b in ((0.3, 2.0), (2.9, 1.7), (1.0, 3.0)):
        n = np.zeros(3, np.float32); ab = np.zeros(2, np.float32)
        L.orc_dir2normal(a, b, n.ctypes.data_as(C.c_void_p))
        assert abs(np.linalg.norm(n) - 1) < 1e-6
        L.orc_normal2dir(n.ctypes.data_as(C.c_void_p), ab.ctypes.data_as(C.c_void_p))
        assert abs(ab[0] - a) < 1e-5 and abs(ab[1] - b) < 1e-5


@pytest.fixture(scope="module")
def small():
    return common.make_scene(1, 0.25)


def test_score_is_low_at_ground_truth_and_rises_off_it(small):
    syn, osc, gt, imgs, ok = small
    ref = 0
    d, n = gt[ref]
    s_gt = osc.score_hypotheses(ref, d, n, 0)
    h, w = d.shape
    assert np.all(s_gt[:7] == 2) and np.all(s_gt[:, :7] == 2) and np.all(s_gt[h - 7:] == 2) and np.all(s_gt[:, w - 7:] == 2)
    inner = s_gt[7:-7, 7:-7]
    visible = inner < 0.66 - 1e-6            # thRobust = 1.2 * 0.55: patches seen by fewer than one view score exactly that
    assert visible.mean() > 0.8
    assert np.median(inner[visible]) < 0.02
    s_off = osc.score_hypotheses(ref, d * 1.03, n, 0)[7:-7, 7:-7]
    assert np.median(s_off[visible]) > 10 * np.median(inner[visible])
    # smoothness bonus only ever lowers the score (factors in (0,1])
    s_sm = osc.score_hypotheses(ref, d, n, 1)[7:-7, 7:-7]
    assert np.all(s_sm <= inner + 1e-7)


def test_view_selection_invariants(small):
    syn, osc, gt, imgs, ok = small
    assert all(ok)
    for i in range(syn.n_views):
        nb = osc.neighbors(i, 1)
        assert i not in nb["ids"] and len(nb["ids"]) <= 12
        assert np.all(np.diff(nb["score"]) <= 0)                       # sorted by decreasing score
        assert np.all((nb["angle"] >= np.deg2rad(3)) & (nb["angle"] < np.deg2rad(65)))
        assert np.all((nb["scale"] >= 0.2) & (nb["scale"] < 3.2)) and np.all(nb["area"] >= 0.01)
        assert len(osc.match_views(i)) == 5


def test_estimate_recovers_ground_truth_and_redblack_agrees(small):
    syn, osc, gt, imgs, ok = small
    ref = 3
    res = {}
    for name, mode in (("raster", 0), ("redblack", 2)):
        osc.init_depth_sparse(ref)
        st = osc.estimate(ref, seed=11, threads=1 if mode == 0 else 4, mode=mode, far_reach=11)
        res[name] = osc.get_depthmap(ref)
        assert abs(st["n_hyp"] / st["n_pixel_iters"] - 8.0) < 0.1          # 2 propagation + 6 refinement hypotheses
    g = gt[ref][0]
    for name in res:
        d, n, c = res[name][:3]
        valid = d > 0
        assert valid[7:-7, 7:-7].mean() > 0.9
        assert np.all(d[~valid] == 0) and np.all(c[~valid] == 0)
        assert np.all((c >= 0) & (c <= 1))                                   # EndDepthMapTmp inverted the score
        assert common.agreement(g, d, mask=valid) > 0.93, name
    # identical seeds reproduce (q1), different seeds agree statistically
    osc.init_depth_sparse(ref)
    osc.estimate(ref, seed=11, threads=1, mode=0)
    assert np.array_equal(osc.get_depthmap(ref)[0], res["raster"][0])
    assert common.agreement(res["raster"][0], res["redblack"][0]) > 0.93


def test_end_depthmap_semantics(small):
    syn, osc, gt, imgs, ok = small
    d = np.array([[1.0, 2.0, 0.0, 3.0, 4.0]], np.float32)
    # needs a full-size map: embed in the image
    ref = 5
    h, w = osc.sizes[ref]
    D = np.zeros((h, w), np.float32); Cf = np.zeros((h, w), np.float32); N = np.zeros((h, w, 3), np.float32)
    D[0, :5] = d; Cf[0, :5] = [0.1, 0.55, 0.2, 1.5, 0.5499]; N[0, :5] = [0, 0, -1]
    osc.set_depthmap(ref, D, N, Cf, 0.5, 10)
    osc.end_depthmap(ref)
    d2, n2, c2, _, _ = osc.get_depthmap(ref)
    assert list(d2[0, :5]) == [1.0, 0.0, 0.0, 0.0, 4.0]                      # conf >= 0.55 or depth <= 0 -> removed
    assert np.allclose(c2[0, :5], [0.9, 0, 0, 0, 1 - 0.5499])
    assert np.all(n2[0, 1] == 0) and np.all(n2[0, 0] == [0, 0, -1])


def test_filter_and_fuse_on_ground_truth_maps():
    syn, osc, gt, imgs, ok = common.make_scene(2, 0.125, 9)
    rng = np.random.default_rng(5)
    for i in range(syn.n_views):
        d, n = gt[i]
        conf = np.full(d.shape, 0.8, np.float32)
        d = d.copy()
        if i == 4:
            d[20:30, 20:30] *= 1.5                                            # a blob of wrong depths in the reference view
        osc.set_depthmap(i, d, n, conf, float(d[d > 0].min() * 0.5), float(d.max() * 2))
    nb = list(range(min(8, len(osc.neighbors(4, 1)["ids"]))))
    fd, fc = osc.filter(4, nb, True)
    assert np.all(fd[22:28, 22:28] == 0)                                      # inconsistent depths are discarded
    good = np.ones(fd.shape, bool); good[18:32, 18:32] = False; good &= gt[4][0] > 0
    kept = fd[good] > 0
    assert kept.mean() > 0.9
    assert np.abs(fd[good][kept] / gt[4][0][good][kept] - 1).max() < 0.02     # averaged depth stays on the surface
    cloud = osc.fuse(True, True)
    assert len(cloud["xyz"]) > 1000
    assert cloud["n_views"].min() >= 2                                        # nMinViewsFuse
    z = np.array([syn.height_at(x, y) for x, y, _ in cloud["xyz"][::50]])
    err = np.abs(cloud["xyz"][::50, 2] - z)
    assert np.percentile(err, 90) < 0.01 * syn.cfg.cam_distance               # accuracy vs the analytic surface
    off = np.concatenate([[0], np.cumsum(cloud["n_views"])])
    for k in range(0, len(off) - 1, 97):
        v = cloud["views"][off[k]:off[k + 1]]
        assert np.all(np.diff(v.astype(np.int64)) > 0)                        # PointCloud::pointViews sorted, unique


def test_resize_area_up_matches_opencv():
    """The restore tree brings the previous level's maps to the current size with cv::resize(INTER_AREA)
    (restore/libs/MVS/SceneDensify.cpp:523-524); enlarging, OpenCV runs its linear kernel with "area mode" coordinates.
    The oracle's restatement is compared bit for bit with the cv2 in this image — a third-party pin for that piece."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    for (sh, sw, dh, dw) in ((30, 40, 60, 80), (37, 53, 120, 160), (100, 75, 101, 76), (48, 64, 480, 640), (5, 7, 33, 20), (60, 80, 120, 160)):
        for cn in (1, 3):
            src = rng.uniform(0.1, 10, (sh, sw) if cn == 1 else (sh, sw, cn)).astype(np.float32)
            want = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
            got = O.resize_area_up(src, dw, dh)
            assert np.array_equal(want, got), (sh, sw, dh, dw, cn, np.abs(want - got).max())


# ------------------------------------------------------------------------------------------------ numpy oracles of the cloud / init steps
def _oracle_path():
    import os, sys
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)


def test_cloud_filter_oracle_closed_form():
    """Scene::PointCloudFilter restatement on a hand-made configuration: two points on one viewing ray. The far point's cone (height
    1.02 x 2) holds the near point, which lies in front and is not depth-similar -> the near point loses #views(far); the near point's
    cone ends at 1.02 and never reaches the far point; a third point off the ray by 3 pixels is outside both one-pixel cones."""
    _oracle_path()
    import cloud_filter as CF
    f, w = 1000.0, 1000
    cams = {0: dict(C=np.zeros(3), K=np.array([f, 0, 500, 0, f, 500, 0, 0, 1.0]), width=w), 1: dict(C=np.array([5.0, 0, 0]), K=np.array([f, 0, 500, 0, f, 500, 0, 0, 1.0]), width=w)}
    pts = np.array([[0, 0, 1.0], [0, 0, 2.0], [3.0 / f * 2.0, 0, 2.0]], np.float32)
    off = np.array([0, 1, 3, 4], np.uint32)           # point 0: view 0; point 1: views 0 and 1; point 2: view 0
    views = np.array([0, 0, 1, 0], np.uint32)
    vis = CF.visibility(cams, pts, off, views)
    assert vis.tolist() == [-2, 0, 0]
    # the same ray seen the other way round: a point BEHIND the apex inside its height -> +#views(behind)
    pts2 = np.array([[0, 0, 2.0], [0, 0, 2.03]], np.float32)   # 1.5 % apart: not depth-similar (1 %), inside 1.02 x 2 = 2.04
    vis2 = CF.visibility(cams, pts2, np.array([0, 1, 3], np.uint32), np.array([0, 0, 1], np.uint32))
    assert vis2.tolist() == [-2, 2]                   # cone of point 0 sees point 1 behind it (+2); cone of point 1 (view 0) sees point 0 in front (-2)


def test_point_colors_oracle_closed_form():
    """EstimatePointColors restatement: a point that projects exactly on a pixel centre takes that pixel (weights 1, 0: no truncation
    loss), the nearer of two views wins, the 1-pixel border and unseen points are white; a half-way sample shows the uint8 truncation
    of every product (TPixel<uint8_t>::operator*)."""
    _oracle_path()
    import point_colors as PC
    K = np.array([[100.0, 0, 8], [0, 100.0, 6], [0, 0, 1]])
    P_near = np.hstack([K, np.zeros((3, 1))])                                  # camera at the origin looking down +z
    P_far = np.hstack([K, (K @ np.array([0, 0, 5.0]))[:, None]])               # the same camera moved back by 5
    rng = np.random.default_rng(0)
    img_a = rng.integers(0, 256, (12, 16, 3)).astype(np.uint8); img_b = rng.integers(0, 256, (12, 16, 3)).astype(np.uint8)
    X = lambda u, v, z: [(u - 8) * z / 100.0, (v - 6) * z / 100.0, z]
    pts = np.array([X(5, 4, 2.0), X(5, 4, 2.0), X(0.5, 4, 2.0), X(5.5, 4, 2.0), X(5, 4, 2.0)], np.float32)
    off = np.array([0, 2, 3, 4, 5, 5], np.uint32)
    views = np.array([1, 0, 0, 0, 0], np.uint32)                               # point 0 lists the far view first: the nearer one must still win
    got = PC.estimate_point_colors([P_near, P_far], [img_a, img_b], pts, off, views)
    assert got[0].tolist() == img_a[4, 5].tolist() and got[1].tolist() == img_a[4, 5].tolist()
    assert got[2].tolist() == [255, 255, 255] and got[4].tolist() == [255, 255, 255]   # inside the 1-pixel border / seen by no view
    half = (np.float32(0.5) * img_a[4, 5].astype(np.float32)).astype(np.uint8).astype(np.int32) + (np.float32(0.5) * img_a[4, 6].astype(np.float32)).astype(np.uint8)
    assert got[3].tolist() == half.astype(np.uint8).tolist()                   # floor(a/2) + floor(b/2), not round((a+b)/2)


def test_point_normals_and_triangulated_init_oracles_closed_form():
    """k-NN PCA on an exact plane returns the plane's normal facing the camera; one triangle on a fronto-parallel plane rasterises to a
    constant depth with the top-left fill rule deciding the shared edge of two triangles exactly once."""
    _oracle_path()
    import point_normals as PN
    import triangulate_init as T
    rng = np.random.default_rng(1)
    xy = rng.uniform(-1, 1, (500, 2))
    n_true = np.array([0.2, -0.3, 1.0]); n_true /= np.linalg.norm(n_true)
    pts = np.stack([xy[:, 0], xy[:, 1], -(n_true[0] * xy[:, 0] + n_true[1] * xy[:, 1]) / n_true[2]], 1).astype(np.float32)
    off = np.arange(501, dtype=np.uint32); views = np.zeros(500, np.uint32)
    nrm, gap = PN.estimate_point_normals(pts, off, views, np.array([[0, 0, 10.0], [0, 0, -10.0]]), 16)
    assert np.abs((nrm * n_true).sum(axis=1) - 1).max() < 1e-5 and gap.min() > 1e-2      # camera 0 is on the +n side
    nrm2, _ = PN.estimate_point_normals(pts, off, np.ones(500, np.uint32), np.array([[0, 0, 10.0], [0, 0, -10.0]]), 16)
    assert np.abs((nrm2 * n_true).sum(axis=1) + 1).max() < 1e-5                          # seen from below: flipped
    K = np.array([100.0, 0, 8, 0, 100.0, 6, 0, 0, 1])
    v = np.array([[2.0, 2.0, 4.0], [12.0, 2.0, 4.0], [2.0, 10.0, 4.0], [12.0, 10.0, 4.0]])
    faces = T.canonical_faces(np.array([[0, 1, 2], [1, 3, 2]]), v[:, :2])
    d, n = T.rasterize(v, faces, K, 16, 12)
    inside = np.zeros((12, 16), bool); inside[2:10, 2:12] = True                          # pixel centres in [2, 12) x [2, 10): the fill rule keeps left / top edges
    assert np.array_equal(d > 0, inside) and np.allclose(d[inside], 4.0, rtol=1e-6) and np.allclose(np.abs(n[inside]), [0, 0, 1], atol=1e-6)
    d1, _ = T.rasterize(v, faces[:1], K, 16, 12); d2, _ = T.rasterize(v, faces[1:], K, 16, 12)
    assert not np.any((d1 > 0) & (d2 > 0)) and np.array_equal((d1 > 0) | (d2 > 0), inside)  # the diagonal's pixels belong to exactly one face


def test_speckle_filter_and_gap_interpolation_oracles_hand_cases():
    """oracle/oracle_capi.cpp: orc_remove_small_segments / orc_gap_interpolation on cases small enough to check by hand."""
    import ctypes as C
    L = O.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # a 6x5 map: a 2x2 island of depth 2 inside a field of depth 1 (not similar at 1 %), one hole
    d = np.ones((5, 6), np.float32); d[1:3, 1:3] = 2.0; d[4, 5] = 0
    c = np.full((5, 6), 0.5, np.float32)
    d1, c1 = d.copy(), c.copy()
    assert L.orc_remove_small_segments(p(d1), None, p(c1), 6, 5, 5, C.c_float(0.01)) == 4      # the 4-pixel island dies at speckle size 5
    assert (d1[1:3, 1:3] == 0).all() and (c1[1:3, 1:3] == 0).all() and (d1 == 1).sum() == 25
    d2 = d.copy()
    assert L.orc_remove_small_segments(p(d2), None, None, 6, 5, 4, C.c_float(0.01)) == 0      # ... and survives at 4
    # asymmetric similarity: 1.0 -> 1.0101 passes (|diff|/1.0 = 0.0101 >= 0.0101? no) — use a pair whose two quotients straddle the threshold
    a, b, th = np.float32(1.0), np.float32(1.0102), 0.0101
    assert abs(a - b) / a >= th > abs(a - b) / b                                              # b -> a similar, a -> b not
    row = np.array([[a, b]], np.float32)
    r1 = row.copy()
    # seeds in column-major order: pixel a first; a cannot reach b, so {a} and {b} are two 1-pixel segments
    assert L.orc_remove_small_segments(p(r1), None, None, 2, 1, 2, C.c_float(th)) == 2
    r2 = row[:, ::-1].copy()
    # now b is the first seed and reaches a: one 2-pixel segment survives
    assert L.orc_remove_small_segments(p(r2), None, None, 2, 1, 2, C.c_float(th)) == 0
    # gap interpolation: a row 1 0 0 4 -> not similar (no fill); 1 0 0 1.03 at th 0.05 -> 1.01, 1.02
    g = np.array([[1, 0, 0, 1.03, 0, 0, 0, 0, 0, 0, 0, 0, 1.0]], np.float32)
    cf = np.array([[0.9, 0, 0, 0.4, 0, 0, 0, 0, 0, 0, 0, 0, 0.7]], np.float32)
    assert L.orc_gap_interpolation(p(g), None, p(cf), 13, 1, 7, C.c_float(0.05)) == 2
    assert np.allclose(g[0, :4], [1, 1.01, 1.02, 1.03], atol=1e-6) and np.allclose(cf[0, 1:3], 0.4) and (g[0, 4:12] == 0).all()   # the 8-pixel gap stays
