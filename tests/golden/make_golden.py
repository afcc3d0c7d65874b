#!/usr/bin/env python
"""Generate tests/golden/c1_quarter.npz: outputs of the CPU oracle on the seeded synthetic scene C1 at 1/4 scale
(10 views 160x120). The reference ships no golden vectors and cannot be built here, so these vectors pin the ORACLE
(regression + cross-platform libm drift) and give the GPU tests an oracle-independent target.

    python tests/golden/make_golden.py        # rewrites the fixture (commit the result)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build():
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    out = {}
    ref = 0
    out["image_sha"] = np.array([sha(imgs[i]) for i in range(syn.n_views)])
    nb = osc.neighbors(ref, 1)
    out["nb_ids"] = nb["ids"]; out["nb_score"] = nb["score"]; out["nb_all_count"] = np.array([len(osc.neighbors(i, 0)["ids"]) for i in range(syn.n_views)])
    for k, (ds, ang) in enumerate(((0.0, 0.0), (0.004, 4.0), (0.03, 20.0))):
        d, n = common.perturbed_hypotheses(gt[ref][0], gt[ref][1], syn.K[ref], seed=k, depth_sigma=ds, angle_deg=ang)
        out[f"hyp{k}_depth"] = d; out[f"hyp{k}_normal"] = n
        out[f"hyp{k}_score0"] = osc.score_hypotheses(ref, d, n, 0)
        out[f"hyp{k}_score1"] = osc.score_hypotheses(ref, d, n, 1)
    osc.init_depth_sparse(ref)
    d0, _, _, lo, hi = osc.get_depthmap(ref)
    out["init_depth"] = d0; out["init_range"] = np.array([lo, hi], np.float32)
    out["gramap"] = osc.gramap(ref)
    osc.estimate(ref, seed=3, threads=1, mode=0)
    out["raster_depth"] = osc.get_depthmap(ref)[0]
    osc.init_depth_sparse(ref)
    osc.estimate(ref, seed=3, threads=4, mode=2, far_reach=11)
    rb = osc.get_depthmap(ref)
    out["redblack_depth"] = rb[0]; out["redblack_conf"] = rb[2]
    # filter + fuse on ground-truth maps with deterministic noise
    rng = np.random.default_rng(42)
    for i in range(syn.n_views):
        d, n = gt[i]
        dn = (d * (1 + 0.002 * rng.standard_normal(d.shape))).astype(np.float32)
        dn[rng.uniform(size=d.shape) < 0.03] = 0
        conf = rng.uniform(0.5, 1, d.shape).astype(np.float32); conf[dn == 0] = 0
        out[f"map{i}_depth"] = dn.astype(np.float32); out[f"map{i}_conf"] = conf
        osc.set_depthmap(i, dn, n, conf, float(d[d > 0].min() * 0.5), float(d.max() * 2))
    nbf = list(range(min(8, len(nb["ids"]))))
    fd, fc = osc.filter(ref, nbf, True)
    out["filter_adjust_depth"] = fd; out["filter_adjust_conf"] = fc
    fd, fc = osc.filter(ref, nbf, False)
    out["filter_strict_depth"] = fd
    cloud = osc.fuse(True, True)
    out["fuse_count"] = np.array([len(cloud["xyz"]), len(cloud["views"])])
    out["fuse_views_sha"] = np.array([sha(cloud["views"]), sha(cloud["n_views"])])
    out["fuse_xyz_head"] = cloud["xyz"][:2000]; out["fuse_xyz_sha"] = np.array([sha(cloud["xyz"])])
    out["fuse_colors_head"] = cloud["colors"][:2000]
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "c1_quarter.npz"), **build())
    print("wrote", os.path.join(HERE, "c1_quarter.npz"), os.path.getsize(os.path.join(HERE, "c1_quarter.npz")) // 1024, "KiB")
