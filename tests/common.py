"""Shared scene builders for the tests: one synthetic scene fed to the CPU oracle and to the CUDA library."""
import numpy as np

import oracle_lib as O
from hcmvs_b200.synth import SynthScene

BENCH_PARAMS = dict(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)

_cache = {}


def make_scene(config=1, scale=0.25, n_views=0, **params):
    """Synthetic scene -> (syn, oracle_scene, gt, images). View selection is run in the oracle."""
    key = (config, scale, n_views, tuple(sorted(params.items())))
    if key in _cache:
        return _cache[key]
    p = dict(BENCH_PARAMS); p.update(params)
    syn = SynthScene(config, scale, n_views)
    osc = O.OracleScene(**p)
    gt, imgs = [], []
    for i in range(syn.n_views):
        bgr, d, n = syn.render(i)
        osc.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr=bgr)
        gt.append((d, n)); imgs.append(bgr)
    osc.set_sparse(syn.sparse_xyz, syn.sparse_off, syn.sparse_views)
    ok = []
    for i in range(syn.n_views):
        r = osc.select_views(i)
        m = osc.init_views(i, p["nNumViews"]) if r > 0 else -1
        ok.append(r > 0 and m > 0)
    out = (syn, osc, gt, imgs, ok)
    _cache[key] = out
    return out


def make_context(syn, osc, imgs, ok, device=0, **params):
    """CUDA context loaded with the same images, cameras and (oracle-selected) neighbours."""
    from hcmvs_b200 import api
    p = {k: v for k, v in BENCH_PARAMS.items()}
    p.update(params)
    ctx = api.Context(device, **p)
    for i in range(syn.n_views):
        ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], osc.gray(i), imgs[i])
    for i in range(syn.n_views):
        if not ok[i]:
            continue
        nb = osc.neighbors(i, which=1)
        ctx.set_neighbors(i, nb["ids"], len(osc.match_views(i)), nb["score"])
        ctx.set_fuse_priority(i, len(osc.neighbors(i, which=0)["ids"]))
    return ctx


def agreement(a, b, th=0.01, mask=None):
    """Fraction of pixels valid in `a` (and in mask) whose depth in b is within th relative."""
    valid = a > 0
    if mask is not None:
        valid &= mask
    rel = np.abs(a - b) / np.maximum(a, 1e-12)
    return float(((rel < th) & valid & (b > 0)).sum()) / max(int(valid.sum()), 1)


def perturbed_hypotheses(gt_depth, gt_normal, K, seed, depth_sigma=0.01, angle_deg=10.0):
    """Per-pixel hypotheses around the ground truth: depth*(1+N(0,s)), normal rotated by up to angle_deg, kept facing the camera."""
    rng = np.random.default_rng(seed)
    h, w = gt_depth.shape
    d = (gt_depth * (1.0 + depth_sigma * rng.standard_normal((h, w)))).astype(np.float32)
    d = np.maximum(d, 1e-3).astype(np.float32)
    n = gt_normal.astype(np.float64) + np.tan(np.deg2rad(angle_deg)) * rng.uniform(-1, 1, (h, w, 3))
    n /= np.linalg.norm(n, axis=2, keepdims=True)
    ys, xs = np.mgrid[0:h, 0:w]
    ray = np.stack([(xs - K[2]) / K[0], (ys - K[5]) / K[4], np.ones_like(xs, dtype=np.float64)], axis=2)
    flip = (n * ray).sum(axis=2) >= 0
    n[flip] *= -1
    # a zero GT normal (ray missed the surface) would be invalid: replace by facing-the-camera
    bad = ~np.isfinite(n).all(axis=2)
    n[bad] = np.array([0, 0, -1.0])
    return d, n.astype(np.float32)
