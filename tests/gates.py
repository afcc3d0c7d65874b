"""The north-star acceptance gates as functions (BASELINE.json `north_star`), shared by the CPU and GPU tests.

Depth maps: the reference's own comparison, MVS::CompareDepthMaps (libs/MVS/DepthMap.cpp:2931-3046): over the WHOLE map (no mask), a
pixel that is valid in one map only is `missing` / `extra`; a pixel valid in both whose relative error |d - d_ref| / d_ref exceeds
the threshold (0.01) is an `error` pixel. Fused clouds: accuracy = distance from each point of the tested cloud to the nearest
point of the reference cloud, completeness = the same with the roles exchanged (the DTU / ETH3D definitions), reported as the
fraction below a threshold and the median.
"""
import numpy as np


def compare_depth_maps(depth, depth_ref, threshold=0.01):
    """CompareDepthMaps(depthMap, depthMapGT, ., threshold) -> dict of fractions of the map area + `within` = the north star's
    fraction of the reference's valid pixels that the tested map reproduces within `threshold` (missing pixels count as failures)."""
    d = np.asarray(depth, np.float32); g = np.asarray(depth_ref, np.float32)
    assert d.shape == g.shape
    area = float(d.size)
    extra = (d != 0) & (g == 0)
    missing = (d == 0) & (g != 0)
    both = (d != 0) & (g != 0)
    err = np.zeros(d.shape, np.float32)
    err[both] = np.abs(d[both] - g[both]) / g[both]
    n_err = int((both & (err > threshold)).sum())
    n_ref = int((g != 0).sum())
    return dict(error=n_err / area, missing=float(missing.sum()) / area, extra=float(extra.sum()) / area,
                median=float(np.median(err[both])) if both.any() else 0.0,
                within=float((both & (err <= threshold)).sum()) / max(n_ref, 1), n_ref=n_ref)


def cloud_distances(points, points_ref):
    """Nearest-neighbour distance from every point of `points` to `points_ref` (k-d tree)."""
    from scipy.spatial import cKDTree
    if len(points) == 0 or len(points_ref) == 0:
        return np.full(len(points), np.inf)
    return cKDTree(np.asarray(points_ref, np.float64)).query(np.asarray(points, np.float64), k=1, workers=-1)[0]


def cloud_accuracy_completeness(points, points_ref, threshold):
    """-> dict(accuracy, completeness: fractions within `threshold`; acc_median, comp_median)."""
    a = cloud_distances(points, points_ref)
    c = cloud_distances(points_ref, points)
    return dict(accuracy=float(np.mean(a <= threshold)) if len(a) else 0.0, completeness=float(np.mean(c <= threshold)) if len(c) else 0.0,
                acc_median=float(np.median(a)) if len(a) else np.inf, comp_median=float(np.median(c)) if len(c) else np.inf,
                n=len(points), n_ref=len(points_ref))


def gt_cloud(syn, gt, stride=1):
    """The synthetic ground-truth surface as a cloud: every `stride`-th pixel of every view lifted with its exact depth."""
    out = []
    for i, (d, _) in enumerate(gt):
        h, w = d.shape
        ys, xs = np.mgrid[0:h:stride, 0:w:stride]
        z = d[::stride, ::stride].astype(np.float64)
        K = np.asarray(syn.K[i], np.float64).reshape(3, 3); R = np.asarray(syn.R[i], np.float64).reshape(3, 3); C = np.asarray(syn.Cc[i], np.float64)
        m = z > 0
        Xc = np.stack([(xs[m] - K[0, 2]) / K[0, 0] * z[m], (ys[m] - K[1, 2]) / K[1, 1] * z[m], z[m]], 1)
        out.append(Xc @ R + C)  # R^T Xc + C
    return np.concatenate(out).astype(np.float32)


def visible_in_views(syn, gt, ref, view_ids, border=8):
    """Per pixel of `ref`: in how many of `view_ids` does its ground-truth point project at least `border` px inside the frame.
    Computed from the scene geometry alone (cameras + exact depth) — independent of any estimator."""
    d = gt[ref][0].astype(np.float64)
    h, w = d.shape
    ys, xs = np.mgrid[0:h, 0:w]
    K = np.asarray(syn.K[ref], np.float64).reshape(3, 3); R = np.asarray(syn.R[ref], np.float64).reshape(3, 3); C = np.asarray(syn.Cc[ref], np.float64)
    Xc = np.stack([(xs - K[0, 2]) / K[0, 0] * d, (ys - K[1, 2]) / K[1, 1] * d, d], 2).reshape(-1, 3)
    Xw = Xc @ R + C
    cnt = np.zeros(h * w, np.int32)
    for v in view_ids:
        Kv = np.asarray(syn.K[v], np.float64).reshape(3, 3); Rv = np.asarray(syn.R[v], np.float64).reshape(3, 3); Cv = np.asarray(syn.Cc[v], np.float64)
        Y = (Xw - Cv) @ Rv.T
        with np.errstate(divide="ignore", invalid="ignore"):
            u = Kv[0, 0] * Y[:, 0] / Y[:, 2] + Kv[0, 2]; t = Kv[1, 1] * Y[:, 1] / Y[:, 2] + Kv[1, 2]
        hv, wv = gt[v][0].shape
        cnt += ((Y[:, 2] > 0) & (u >= border) & (t >= border) & (u <= wv - 1 - border) & (t <= hv - 1 - border)).astype(np.int32)
    return (cnt.reshape(h, w) * (d > 0)).astype(np.int32)
