#!/usr/bin/env python
"""Generate tests/golden/c1_quarter_prior.npz: the plane-prior term of the cost (libs/MVS/DepthMap.cpp:941-955), live in the authors'
schedule (`--n-photo2geo 1 --n-para_prior 0.4/0.6`, data/*/run.py) from the second outer iteration on:

    score_v = score_v * (1 - para_prior) + 2 * (1 - exp(-(|prior - d| / prior)^2 / (2 sigma^2))) * para_prior      per matching view

computed by the ORACLE on view 2 of the seeded scene C1 at 1/4 scale with a synthetic prior map (the ground-truth depth pulled off the
surface by a smooth +-6 % ripple, 0 = "no prior" in a checkerboard of 16-pixel cells — both branches of `depthMapPrior(x0) != 0`).

    python tests/golden/make_golden_prior.py    # rewrites the fixture (commit the result)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import common  # noqa: E402

REF = 2
PRIOR_PARAMS = dict(para_prior=0.4, fsigmaPrior=0.2)


def prior_map(gt_depth):
    h, w = gt_depth.shape
    ys, xs = np.mgrid[0:h, 0:w]
    ripple = 1.0 + 0.06 * np.sin(xs / 9.0) * np.cos(ys / 7.0)
    p = (gt_depth.astype(np.float64) * ripple).astype(np.float32)
    p[((xs // 16) + (ys // 16)) % 3 == 0] = 0      # cells without a prior
    return p


def build():
    syn, osc, gt, imgs, ok = common.make_scene(1, 0.25)
    out = {}
    prior = prior_map(gt[REF][0])
    out["prior"] = prior
    d, n = common.perturbed_hypotheses(gt[REF][0], gt[REF][1], syn.K[REF], seed=4242, depth_sigma=0.01, angle_deg=6.0)
    out["hyp_depth"] = d; out["hyp_normal"] = n
    # ScorePixel with the prior active: the parity hook scores at outer iteration 0, so photo2geo = 0 switches the term on
    osc.set_params(photo2geo=0, **PRIOR_PARAMS)
    out["score_noprior"] = osc.score_hypotheses(REF, d, n, 0)
    osc.set_prior(REF, prior)
    out["score_prior0"] = osc.score_hypotheses(REF, d, n, 0)
    out["score_prior1"] = osc.score_hypotheses(REF, d, n, 1)
    # the authors' schedule: photo2geo = 1 -> outer iteration 0 ignores the prior, outer iteration 1 uses it
    over = dict(nEstimationIters=2, nEstimationIters_external=2, propagatehalfwin=5, propagatestep=4, photo2geo=1)
    osc.set_params(**over)
    osc.init_depth_sparse(REF)
    init = osc.get_depthmap(REF)
    out["init_depth"] = init[0]; out["init_range"] = np.array(init[3:], np.float32)
    osc.estimate(REF, it_external=0, seed=71, threads=4, mode=2, far_reach=11)
    out["it0_depth"] = osc.get_depthmap(REF)[0]
    osc.estimate(REF, it_external=1, seed=71, threads=4, mode=2, far_reach=11)
    m = osc.get_depthmap(REF)
    out["it1_depth"] = m[0]; out["it1_conf"] = m[2]
    # the same two outer iterations WITHOUT the prior: what the term changes
    osc.set_prior(REF, None)
    osc.set_depthmap(REF, *init)
    osc.estimate(REF, it_external=0, seed=71, threads=4, mode=2, far_reach=11)
    osc.estimate(REF, it_external=1, seed=71, threads=4, mode=2, far_reach=11)
    out["it1_depth_noprior"] = osc.get_depthmap(REF)[0]
    osc.set_params(nEstimationIters=3, nEstimationIters_external=1, propagatehalfwin=1, propagatestep=4, photo2geo=2, para_prior=0.3, fsigmaPrior=0.2)
    return out


if __name__ == "__main__":
    path = os.path.join(HERE, "c1_quarter_prior.npz")
    np.savez_compressed(path, **build())
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
