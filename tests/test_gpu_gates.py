"""The north-star acceptance gates at BASELINE sizes, UNMASKED (tests/gates.py holds the definitions):

* full-size C1 (10 x 640x480), all views, through the end-to-end `DenseReconstruction` call, for both initialisations
  (nMinViewsTrustPoint 1 = splat + random start, 2 = the reference's default triangulated start): the raw maps the reference saves
  (depthNNNN.dmap), the filtered maps and the fused cloud against the RASTER oracle pipeline (1 thread per view, fixed seed);
* one full-size C2 view (1600x1200) against the threaded red-black oracle (same algorithm, same counter RNG) and the raster oracle.

The reference's own criterion is MVS::CompareDepthMaps (DepthMap.cpp:2931-3046): over the whole map, a pixel valid in both maps with
relative error > 0.01 is an error pixel. Raw PatchMatch maps contain, in BOTH implementations, the rim of the image that fewer than two
matching views see: depth there is noise (conf stays high because a single view can always be matched somewhere), so two runs of the
reference itself with different seeds (release builds seed from std::random_device, DepthMap.cpp:395-397) agree with each other only as
far as that rim allows. The gate therefore has three unmasked parts and one geometric one, all asserted:
  G1  raw maps, whole map: the GPU reproduces the reference's valid pixels within 1 % as well as a SECOND RUN OF THE REFERENCE does
      (another seed; -1 pt on the mean over the views, -1.5 pt per view). Measured here (CPU statements, full C1): reference vs
      reference 97.2-97.4 % on the worst view, 98.4-98.5 % on average; red-black vs reference 96.8-97.1 % / 98.0-98.2 % — a flat
      ">= 98 % on every view" is not met by the reference against itself, which is why G2 and G3 exist;
  G2  raw maps restricted to the pixels whose GROUND-TRUTH point projects inside >= 2 matching views (scene geometry only — no
      estimator decides the mask): >= 98 % per view;
  G3  final maps of the benchmarked pipeline (after the geometric-consistency filter), whole map: >= 98 % per view;
  G4  fused cloud vs the raster-oracle cloud: accuracy and completeness >= 98 % at 0.25 % of the scene depth, and against the
      ground-truth surface no worse than the reference cloud by more than 1 pt.
"""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import common
import gates

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

pytestmark = pytest.mark.gpu

CORES = os.cpu_count() or 4


def _oracle_inits(syn, osc, imgs, trust):
    """Initial maps per view as the reference builds them (SceneDensify.cpp:781-812): splat of the sparse points, or the triangulated mesh."""
    out = []
    if trust < 2:
        for i in range(syn.n_views):
            osc.init_depth_sparse(i)
            out.append(osc.get_depthmap(i))
        return out
    import triangulate_init as T
    from test_triangulate_init import _compose_p, _view_points, _avg_depth
    for i in range(syn.n_views):
        idx, xyz = _view_points(syn, i)
        P = _compose_p(syn.K[i], syn.R[i], syn.Cc[i])
        h, w = imgs[i].shape[:2]
        v, t, lo, hi = T.triangulate(P, syn.K[i], w, h, xyz[idx], _avg_depth(P, xyz[idx]))
        d, n = T.rasterize(v, t, syn.K[i], w, h)
        out.append((d, n, np.zeros_like(d), float(np.float32(lo) * np.float32(0.9)), float(np.float32(hi) * np.float32(1.1))))
    return out


def _oracle_pipeline(syn, osc, init, seed, mode=0):
    """Estimate every view with the RASTER oracle (mode 0, one thread per view: deterministic), filter (8 neighbours, adjust), fuse."""
    V = syn.n_views
    for i in range(V):
        osc.set_depthmap(i, *init[i])
    with ThreadPoolExecutor(min(CORES, V)) as ex:
        list(ex.map(lambda i: osc.estimate(i, seed=seed, threads=1 if mode == 0 else 2, mode=mode, far_reach=11), range(V)))
    raw = [osc.get_depthmap(i) for i in range(V)]
    filt = []
    for i in range(V):
        nb = list(range(min(8, len(osc.neighbors(i, 1)["ids"]))))
        filt.append(osc.filter(i, nb, True))
    for i in range(V):
        d, n, c, lo, hi = raw[i]
        osc.set_depthmap(i, filt[i][0], n, filt[i][1], lo, hi)
    cloud = osc.fuse(True, True)
    return raw, filt, cloud


@pytest.mark.parametrize("trust", [1, 2])
def test_c1_full_size_all_views_against_the_raster_oracle(trust, tmp_path):
    from hcmvs_b200 import api, host
    syn, osc, gt, imgs, ok = common.make_scene(1, 1.0)
    V = syn.n_views
    scene_depth = float(np.median(gt[0][0][gt[0][0] > 0]))
    osc.set_params(nMinViewsTrustPoint=trust)
    init = _oracle_inits(syn, osc, imgs, trust)
    t0 = time.time()
    refA = _oracle_pipeline(syn, osc, init, seed=9)
    refB = _oracle_pipeline(syn, osc, init, seed=1234)   # the reference's own seed-to-seed spread
    t_oracle = time.time() - t0
    # ---- the product: ONE end-to-end call (select views, upload, estimate, filter, fuse), raw maps streamed to .dmap files
    params = dict(common.BENCH_PARAMS); params["nMinViewsTrustPoint"] = trust
    ctx = api.Context(0, **params)
    hs = host.HostScene.from_synth(syn, imgs)
    try:
        dm = str(tmp_path)
        st = hs.dense_reconstruction(ctx, seed=9, run_filter=True, dmap_dir=dm)
        cloud = hs.cloud()
        raw = [host.read_dmap(os.path.join(dm, f"depth{i:04d}.dmap")) for i in range(V)]
        # the filtered maps are not an output of DenseReconstruction: reproduce them from the raw maps through the C ABI
        for i in range(V):
            r = raw[i]
            ctx.set_depthmap(i, r["depth"], r["normal"], r["conf"], r["dmin"], r["dmax"])
        filt = [ctx.filter_depthmap(i, list(range(min(8, len(osc.neighbors(i, 1)["ids"])))), True) for i in range(V)]
    finally:
        ctx.close(); hs.close()
    g1, g1_self, g2, g3 = [], [], [], []
    for i in range(V):
        a = gates.compare_depth_maps(raw[i]["depth"], refA[0][i][0])
        g1.append(a)
        g1_self.append(gates.compare_depth_maps(refB[0][i][0], refA[0][i][0]))
        vis = gates.visible_in_views(syn, gt, i, [int(v) for v in osc.match_views(i)]) >= 2
        rv = (refA[0][i][0] > 0) & vis
        d, r = raw[i]["depth"], refA[0][i][0]
        g2.append(float(((d > 0) & rv & (np.abs(d - r) <= 0.01 * r)).sum()) / max(int(rv.sum()), 1))
        g3.append(gates.compare_depth_maps(filt[i][0], refA[1][i][0]))
    th = 0.0025 * scene_depth
    c_ref = gates.cloud_accuracy_completeness(cloud["xyz"], refA[2]["xyz"], th)
    G = gates.gt_cloud(syn, gt, 2)
    c_gt_gpu = gates.cloud_accuracy_completeness(cloud["xyz"], G, th)
    c_gt_ref = gates.cloud_accuracy_completeness(refA[2]["xyz"], G, th)
    w1 = [a["within"] for a in g1]; w3 = [a["within"] for a in g3]
    print(f"\n[C1 full, nMinViewsTrustPoint {trust}] oracle {t_oracle:.0f} s; GPU e2e {st['sec_estimate'] + st['sec_filter'] + st['sec_fuse']:.3f} s, {st['n_points']} points (oracle {len(refA[2]['xyz'])})")
    print(f"  G1 raw, unmasked, within 1 % of the raster oracle: min {min(w1):.4f} mean {np.mean(w1):.4f} | error px max {max(a['error'] for a in g1):.4f} "
          f"missing max {max(a['missing'] for a in g1):.4f} extra max {max(a['extra'] for a in g1):.4f}")
    ws = [a["within"] for a in g1_self]
    print(f"     the raster oracle against itself (seed 1234 vs 9): min {min(ws):.4f} mean {np.mean(ws):.4f}")
    print(f"  G2 raw, GT point inside >= 2 matching views: min {min(g2):.4f} mean {np.mean(g2):.4f}")
    print(f"  G3 filtered, unmasked: min {min(w3):.4f} mean {np.mean(w3):.4f} | error px max {max(a['error'] for a in g3):.4f}")
    print(f"  G4 cloud vs raster-oracle cloud @ {th:.4f}: accuracy {c_ref['accuracy']:.4f} completeness {c_ref['completeness']:.4f} (medians {c_ref['acc_median']:.5f} / {c_ref['comp_median']:.5f})")
    print(f"     vs ground truth: GPU acc {c_gt_gpu['accuracy']:.4f} comp {c_gt_gpu['completeness']:.4f} | oracle acc {c_gt_ref['accuracy']:.4f} comp {c_gt_ref['completeness']:.4f}")
    assert np.mean(w1) >= min(np.mean(ws), 0.98) - 0.01 and np.mean(w1) >= 0.97   # G1
    for i in range(V):
        assert w1[i] >= min(ws[i], 0.98) - 0.015, (i, w1[i], ws[i])
    assert min(g2) >= 0.98                                                # G2
    assert min(w3) >= 0.98                                                # G3
    assert c_ref["accuracy"] >= 0.98 and c_ref["completeness"] >= 0.98    # G4
    assert c_gt_gpu["accuracy"] >= c_gt_ref["accuracy"] - 0.01 and c_gt_gpu["completeness"] >= c_gt_ref["completeness"] - 0.01
    assert abs(st["n_points"] - len(refA[2]["xyz"])) <= 0.01 * len(refA[2]["xyz"])
    osc.set_params(nMinViewsTrustPoint=1)


def test_c2_full_size_view_against_both_oracles():
    """One 1600x1200 view of C2: the GPU against the red-black restatement (>= 0.995, unmasked) and the reference's raster sweep run with
    the reference's own pixel-stealing threads (unmasked, CompareDepthMaps definition)."""
    syn, osc, gt, imgs, ok = common.make_scene(2, 1.0, 12)
    ctx = common.make_context(syn, osc, imgs, ok)
    ref = 5
    try:
        osc.init_depth_sparse(ref)
        init = osc.get_depthmap(ref)
        ctx.init_depthmap(ref, init[0], None, init[3], init[4])
        ctx.estimate_depthmap(ref, 0, seed=17)
        gd = ctx.get_depthmap(ref)[0]
    finally:
        ctx.close()
    t0 = time.time()
    osc.estimate(ref, seed=17, threads=CORES, mode=2, far_reach=11)
    rb = osc.get_depthmap(ref)[0]
    t_rb = time.time() - t0
    osc.set_depthmap(ref, *init)
    t0 = time.time()
    osc.estimate(ref, seed=17, threads=CORES, mode=0)
    ra = osc.get_depthmap(ref)[0]
    t_ra = time.time() - t0
    a_rb = gates.compare_depth_maps(gd, rb)
    a_ra = gates.compare_depth_maps(gd, ra)
    a_rr = gates.compare_depth_maps(rb, ra)
    vis = gates.visible_in_views(syn, gt, ref, [int(v) for v in osc.match_views(ref)]) >= 2
    rv = (ra > 0) & vis
    g2 = float(((gd > 0) & rv & (np.abs(gd - ra) <= 0.01 * ra)).sum()) / max(int(rv.sum()), 1)
    print(f"\n[C2 view {ref}, 1600x1200] oracle red-black {t_rb:.0f} s, raster ({CORES} threads) {t_ra:.0f} s")
    print(f"  GPU vs red-black oracle, unmasked: within {a_rb['within']:.4f} error {a_rb['error']:.4f} missing {a_rb['missing']:.4f} extra {a_rb['extra']:.4f}")
    print(f"  GPU vs raster oracle, unmasked:    within {a_ra['within']:.4f} error {a_ra['error']:.4f} missing {a_ra['missing']:.4f} extra {a_ra['extra']:.4f} (red-black oracle vs raster: {a_rr['within']:.4f})")
    print(f"  GPU vs raster oracle where the GT point is inside >= 2 matching views: {g2:.4f}")
    assert a_rb["within"] >= 0.995
    assert a_ra["within"] >= 0.98
    assert g2 >= 0.98
