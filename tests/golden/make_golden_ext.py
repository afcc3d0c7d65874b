#!/usr/bin/env python
"""Generate tests/golden/c1_quarter_ext.npz: oracle outputs for the paths added after the first fixture — the restore tree's
coarse-level hand-off (INTER_AREA enlargement + last-iteration hypothesis), viewspread, and a rescaled matching view
(ViewData::ScaleImage). Same role as make_golden.py: pins the ORACLE and gives the GPU tests an oracle-independent target.

    python tests/golden/make_golden_ext.py    # rewrites the fixture (commit the result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402
import oracle_lib as O  # noqa: E402
from make_golden import sha  # noqa: E402


def pool2(a):
    h, w = a.shape[0] // 2 * 2, a.shape[1] // 2 * 2
    a = a[:h, :w].astype(np.float64)
    return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]) / 4).astype(np.float32)


def build():
    from hcmvs_b200 import host
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    out = {}
    # ---- cv::resize(INTER_AREA) enlarging, as restated by the oracle
    rng = np.random.default_rng(11)
    src1 = rng.uniform(0.5, 9, (37, 53)).astype(np.float32); src3 = rng.uniform(-1, 1, (37, 53, 3)).astype(np.float32)
    out["resize_src1"] = src1; out["resize_src3"] = src3
    out["resize_dst1"] = O.resize_area_up(src1, 120, 90); out["resize_dst3_sha"] = np.array([sha(O.resize_area_up(src3, 120, 90))])
    # ---- coarse-level hand-off
    ref = 3
    cd = pool2(gt[ref][0]); cn = pool2(gt[ref][1]); cn /= np.maximum(np.linalg.norm(cn, axis=2, keepdims=True), 1e-9)
    out["coarse_depth"] = cd; out["coarse_normal"] = cn
    osc.set_params(nEstimationIters=2, nEstimationIters_external=1)
    osc.init_depth_sparse(ref)
    osc.set_coarse(ref, cd, cn)
    out["coarse_resized_depth"] = osc.get_coarse(ref)[0]
    out["coarse_range"] = np.array(osc.get_depthmap(ref)[3:], np.float32)
    d0 = osc.get_depthmap(ref)
    out["coarse_init_depth"] = d0[0]
    osc.estimate(ref, seed=33, threads=4, mode=2, far_reach=11)
    out["coarse_redblack_depth"] = osc.get_depthmap(ref)[0]
    osc.set_coarse(ref, None, None)
    osc.set_params(nEstimationIters=3, nEstimationIters_external=1)
    # ---- a rescaled matching view
    ref = 5
    slot = 1
    nb = int(osc.match_views(ref)[slot])
    for tag, scale in (("dn", 0.8), ("up", 1.25)):
        g, Ks = host.scale_image(osc.gray(nb), scale, syn.K[nb])
        osc.set_neighbor_image(ref, slot, Ks, g)
        d, n = common.perturbed_hypotheses(gt[ref][0], gt[ref][1], syn.K[ref], seed=900, depth_sigma=0.01, angle_deg=8.0)
        out[f"scaled_{tag}_gray_sha"] = np.array([sha(g)]); out[f"scaled_{tag}_K"] = Ks
        out[f"scaled_{tag}_score0"] = osc.score_hypotheses(ref, d, n, 0)
        osc.set_neighbor_image(ref, slot, None, None)
    out["scaled_hyp_depth"] = d; out["scaled_hyp_normal"] = n
    # ---- viewspread at outer iteration 1
    ref = 4
    over = dict(nEstimationIters=1, nEstimationIters_external=2, propagatehalfwin=5, propagatestep=4, viewspread=1)
    osc.set_params(**over)
    views = [ref] + [int(v) for v in osc.match_views(ref)]
    for v in views:
        osc.init_depth_sparse(v)
        osc.estimate(v, it_external=0, seed=50 + v, threads=4, mode=2, far_reach=11)
        m = osc.get_depthmap(v)
        out[f"vs_view{v}_depth"] = m[0]; out[f"vs_view{v}_normal"] = m[1]; out[f"vs_view{v}_conf"] = m[2]; out[f"vs_view{v}_range"] = np.array(m[3:], np.float32)
    out["vs_views"] = np.array(views)
    osc.snapshot_maps()
    osc.estimate(ref, it_external=1, seed=61, threads=4, mode=2, far_reach=11)
    out["vs_redblack_depth"] = osc.get_depthmap(ref)[0]
    osc.set_params(nEstimationIters=3, nEstimationIters_external=1, propagatehalfwin=1, propagatestep=4, viewspread=0)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "c1_quarter_ext.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
