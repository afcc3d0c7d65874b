#!/usr/bin/env python
"""Generate tests/golden/c1_quarter_triinit.npz: the reference's default depth-map initialisation (TriangulatePoints2DepthMap,
libs/MVS/DepthMap.cpp:1797-1936) computed by the numpy / scipy ORACLE (oracle/triangulate_init.py: Qhull Delaunay, numpy rasteriser)
for views 0 and 7 of the seeded synthetic scene C1 at 1/4 scale. The fixture pins the product's host triangulation (CPU test) and
the device rasteriser (GPU test) against values that do not depend on the product's code; also the point colours and the point-cloud
filter votes of a small seeded cloud (oracle/point_colors.py, oracle/cloud_filter.py).

    python tests/golden/make_golden_triinit.py        # rewrites the fixture (commit the result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import common  # noqa: E402
import triangulate_init as T  # noqa: E402
import point_colors as PC  # noqa: E402
import cloud_filter as CF  # noqa: E402
from test_triangulate_init import _compose_p, _view_points, _avg_depth  # noqa: E402


def seeded_cloud(syn):
    rng = np.random.default_rng(77)
    n = 1500
    xy = rng.uniform(-0.8, 0.8, (n, 2))
    z = 0.05 * xy[:, 0] + 0.03 * xy[:, 1]
    layer = rng.uniform(size=n)
    z = np.where(layer < 0.2, z + rng.uniform(0.2, 1.5, n), np.where(layer < 0.4, z - rng.uniform(0.2, 1.5, n), z + rng.normal(0, 0.002, n)))
    pts = np.stack([xy[:, 0], xy[:, 1], z], 1).astype(np.float32)
    counts = rng.integers(1, 5, n)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    views = np.concatenate([np.sort(rng.choice(syn.n_views, c, replace=False)) for c in counts]).astype(np.uint32)
    return pts, off, views


def build():
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    out = {}
    for ref in (0, 7):
        idx, xyz = _view_points(syn, ref)
        P = _compose_p(syn.K[ref], syn.R[ref], syn.Cc[ref])
        h, w = imgs[ref].shape[:2]
        v, t, lo, hi = T.triangulate(P, syn.K[ref], w, h, xyz[idx], _avg_depth(P, xyz[idx]))
        d, n = T.rasterize(v, t, syn.K[ref], w, h)
        out[f"v{ref}_vertices"] = v; out[f"v{ref}_faces"] = t.astype(np.uint32); out[f"v{ref}_range"] = np.array([lo, hi], np.float32)
        out[f"v{ref}_depth"] = d; out[f"v{ref}_normal"] = n
    pts, off, views = seeded_cloud(syn)
    P_list = [_compose_p(syn.K[v], syn.R[v], syn.Cc[v]) for v in range(syn.n_views)]
    cams = {v: dict(C=np.asarray(syn.Cc[v], np.float64), K=np.asarray(syn.K[v], np.float64).ravel(), width=imgs[v].shape[1]) for v in range(syn.n_views)}
    out["cloud_points"] = pts; out["cloud_offsets"] = off; out["cloud_views"] = views
    out["cloud_colors"] = PC.estimate_point_colors(P_list, imgs, pts, off, views)
    out["cloud_visibility"] = CF.visibility(cams, pts, off, views)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "c1_quarter_triinit.npz"), **build())
    print("written", os.path.join(HERE, "c1_quarter_triinit.npz"))
