"""the stock speckle filter (DepthMapsData::RemoveSmallSegments, the body the fork keeps under `#if 0`,
SceneDensify.cpp:1956-2042) and the stock small-gap branch of DepthMapsData::GapInterpolation (:2294-2352, :2640-2683) on the device,
against the CPU restatements in oracle/oracle_capi.cpp. The speckle filter must be IDENTICAL although the CPU's flood fill is order
dependent (asymmetric similarity test, column-major seeds)."""
import ctypes as C

import numpy as np
import pytest

import common
import oracle_lib as O

pytestmark = pytest.mark.gpu


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _speckled_map(gt, seed, noise):
    """GT depth with multiplicative noise of the order of the similarity threshold (0.7 %: many one-directional edges), random holes,
    random islands of outliers of every size around the speckle threshold."""
    rng = np.random.default_rng(seed)
    d, n = gt
    depth = (d * (1 + noise * rng.standard_normal(d.shape))).astype(np.float32)
    h, w = d.shape
    depth[rng.uniform(size=d.shape) < 0.03] = 0
    for _ in range(300):
        r = int(rng.integers(1, 9)); cy = int(rng.integers(r, h - r)); cx = int(rng.integers(r, w - r))
        yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
        m = yy * yy + xx * xx <= r * r
        patch = depth[cy - r:cy + r + 1, cx - r:cx + r + 1]
        patch[m] = patch[m] * np.float32(rng.uniform(1.05, 1.4))
    conf = rng.uniform(0.1, 1.0, d.shape).astype(np.float32)
    conf[depth == 0] = 0
    nn = n.copy(); nn[depth == 0] = 0
    return depth, nn.astype(np.float32), conf


@pytest.mark.parametrize("speckle,noise", [(100, 0.0035), (20, 0.0035), (100, 0.0005)])
def test_remove_small_segments_identical_to_the_cpu_flood_fill(speckle, noise):
    syn, osc, gt, imgs, ok = common.make_scene(2, 0.5, 6)
    ctx = common.make_context(syn, osc, imgs, ok)
    L = O.lib()
    try:
        for view in (0, 3):
            depth, normal, conf = _speckled_map(gt[view], 40 + view, noise)
            lo, hi = float(depth[depth > 0].min() * 0.5), float(depth.max() * 2)
            ctx.set_depthmap(view, depth, normal, conf, lo, hi)
            removed = ctx.remove_small_segments(view, speckle)
            gd, gn, gc = ctx.get_depthmap(view)[:3]
            od, on, oc = depth.copy(), normal.copy(), conf.copy()
            th = np.float32(ctx.params.fDepthDiffThreshold) * np.float32(0.7)
            want = L.orc_remove_small_segments(_p(od), _p(on), _p(oc), od.shape[1], od.shape[0], speckle, C.c_float(float(th)))
            assert want > 500, "the case is vacuous"
            assert removed == want
            assert np.array_equal(gd, od) and np.array_equal(gc, oc) and np.array_equal(gn, on)
            assert (gd > 0).sum() > 0.5 * (depth > 0).sum()
    finally:
        ctx.close()


def test_gap_interpolation_small_gaps_match_the_cpu():
    syn, osc, gt, imgs, ok = common.make_scene(2, 0.5, 4)
    ctx = common.make_context(syn, osc, imgs, ok)
    L = O.lib()
    try:
        rng = np.random.default_rng(77)
        d, n = gt[1]
        depth = (d * (1 + 0.002 * rng.standard_normal(d.shape))).astype(np.float32)
        h, w = depth.shape
        for _ in range(4000):  # horizontal and vertical runs of 1..12 invalid pixels (gaps above 7 must stay), some touching the border
            ln = int(rng.integers(1, 13)); y = int(rng.integers(0, h)); x = int(rng.integers(0, w))
            if rng.uniform() < 0.5:
                depth[y, x:x + ln] = 0
            else:
                depth[y:y + ln, x] = 0
        depth[40:60, 100:130] *= np.float32(1.2)   # a depth step: gaps across it are not similar
        normal = n.astype(np.float32).copy(); normal[depth == 0] = 0
        conf = rng.uniform(0.1, 1.0, depth.shape).astype(np.float32); conf[depth == 0] = 0
        gd, gn, gc, filled = ctx.gap_interpolation(depth, normal, conf, 7)
        od, on, oc = depth.copy(), normal.copy(), conf.copy()
        th = np.float32(ctx.params.fDepthDiffThreshold) * np.float32(2.5)
        want = L.orc_gap_interpolation(_p(od), _p(on), _p(oc), w, h, 7, C.c_float(float(th)))
        assert want > 5000 and filled == want
        assert np.array_equal(gd, od) and np.array_equal(gc, oc)          # depth / confidence: the same float sequence
        assert np.abs(gn - on).max() <= 2e-6                               # normals go through atan2 / acos / sincos (libm vs CUDA: ulps)
        assert ((od == 0) & (depth == 0)).sum() > 0                        # long gaps and border gaps are left alone
        # depth only (no normal / confidence maps)
        gd2, _, _, filled2 = ctx.gap_interpolation(depth, None, None, 7)
        assert filled2 == want and np.array_equal(gd2, od)
    finally:
        ctx.close()
