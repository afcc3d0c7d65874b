#!/usr/bin/env python
"""Benchmark of the HC-MVS dense-reconstruction hot path (PatchMatch depth estimation + filter + fusion).

One "step" = one pass of the hot path over the whole synthetic scene: for every reference view the
EstimateDepthMap stages (median blur, PASS A, nEstimationIters red-black PatchMatch iterations, PASS C),
then FilterDepthMap for every view, then FuseDepthMaps.

  python bench.py --gpus N --steps K --warmup W        (N>1: launched by torch.distributed.run, one rank per GPU)
  python bench.py --impl reference ...                 (the CPU oracle port of the reference path, host cores)

`value` is BASELINE.json's metric, PatchMatch Mpix*iter/s: pixel-iterations of the whole scene divided by the
whole step time (all stages, inputs resident in HBM). `e2e` is the same metric through the host-facing
DenseReconstruction call with host buffers (H2D of images/initial depth and D2H of maps/cloud inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {1: "C1 synthetic 10-view 640x480 textured plane", 2: "C2 synthetic DTU-shaped 49 views 1600x1200",
             3: "C3 synthetic ETH3D-shaped 20 views 6048x4032", 4: "C4 synthetic video 300 views 1920x1080",
             5: "C5 fusion stress: FuseDepthMaps over 500 precomputed 1920x1080 depth/normal maps"}


def config_dict(args, n_views, width, height):
    """The workload description — IDENTICAL in both arms (the driver compares the two dicts)."""
    if args.config == 5:
        return {"workload": WORKLOADS[5], "scale": args.scale, "views": n_views, "image": [width, height], "neighbours": 12, "patchmatch_iters": 0,
                "stages": "fuse"}
    d = {"workload": WORKLOADS[args.config], "scale": args.scale, "views": n_views, "image": [width, height], "neighbours": 5, "patchmatch_iters": 3,
         "stages": "estimate(A+B+C) + filter + fuse"}
    if args.adapthalfwin != 5:
        d["adapthalfwin"] = args.adapthalfwin
    return d


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=1.0, help="image-size scale of the synthetic scene (1.0 = the named workload)")
    ap.add_argument("--views", type=int, default=0, help="override the number of views (0 = the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sampler", type=int, default=0)
    ap.add_argument("--adapthalfwin", type=int, default=5, help="OPTDENSE::adapthalfwin: 5 = 6x6 texels per patch (BASELINE's configs, the CLI default), 7 = 8x8 (the authors' run.py)")
    ap.add_argument("--e2e-dmap-dir", default="", help="also time the end-to-end call WITH the raw depthNNNN.dmap files written to this directory (streamed behind the GPU)")
    ap.add_argument("--no-split-rows", action="store_true", help="N>1: do not split the views of the incomplete last round into row bands")
    ap.add_argument("--exchange", default="nccl", choices=["nccl", "torch"],
                    help="N>1 map exchange: 'nccl' = hcmvs_exchange_maps (in-place NCCL broadcasts inside the C ABI), 'torch' = torch.distributed all-gather through staging slots")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def make_scene(args, only=None):
    """The seeded synthetic scene; `only`: render just these views (a rank of a multi-GPU run holds the pixels of its share)."""
    from hcmvs_b200.synth import SynthScene
    syn = SynthScene(args.config, args.scale, args.views)
    if only is None:
        return syn, [syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views)]
    return syn, {i: syn.render(i, want_depth=False, want_normal=False)[0] for i in only}


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_sweep launch from the committed `ncu --set full` summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r02b_ncu_k_sweep_tex.txt")
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, found = 0.0, 0
    try:
        with open(path) as f:
            for line in f:
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    if line.startswith(key + " ["):
                        u = line[line.index("[") + 1:line.index("]")]
                        tot += float(line.split("=")[1].split(",")[0]) * unit[u]; found += 1
    except OSError:
        return None
    return tot if found == 2 else None


def flops_per_view_score(texels):
    # 24 flop per texel + ~110 per (hypothesis, view) for H, projections, normalisation
    return 24.0 * texels + 110.0


# ------------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_oracle_sample(args, syn, imgs, seconds_budget=25.0, stages=True):
    """Time the CPU restatement of the reference path (oracle/) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    cores = os.cpu_count() or 1
    ref = syn.n_views // 2
    osc = O.OracleScene(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=args.adapthalfwin)
    for i in range(syn.n_views):
        osc.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr=imgs[i])
    osc.set_sparse(syn.sparse_xyz, syn.sparse_off, syn.sparse_views)
    if osc.select_views(ref) <= 0 or osc.init_views(ref, 5) <= 0:
        raise RuntimeError("oracle view selection failed")
    # bounded sample: as many PatchMatch iterations of ONE reference view as fit the budget (at least 1)
    h, w = osc.sizes[ref]
    inner_pix = (w - 14) * (h - 14)
    osc.init_depth_sparse(ref)
    osc.set_params(nEstimationIters=1)
    t0 = time.time()
    st = osc.estimate(ref, seed=1, threads=cores, mode=0, run_end=False)
    one = time.time() - t0
    iters = 1
    total_s = st["sec_score"] + st["sec_sweeps"]
    pix_iters = st["n_pixel_iters"]
    if one * 3 < seconds_budget:
        osc.init_depth_sparse(ref)
        osc.set_params(nEstimationIters=3)
        st = osc.estimate(ref, seed=1, threads=cores, mode=0, run_end=True)
        iters = 3
        total_s = st["sec_score"] + st["sec_sweeps"] + st["sec_end"]
        pix_iters = st["n_pixel_iters"]
    value = pix_iters / total_s / 1e6
    sample = f"1 of {syn.n_views} reference views ({w}x{h}, 5 neighbours), PASS A + {iters} raster PatchMatch iteration(s), {cores} threads"
    out = {"value": value, "unit": "Mpix*iter/s", "cores": cores, "kind": "port", "sample": sample,
           "hyp_per_pixel_iter": st["n_hyp"] / max(st["n_pixel_iters"], 1), "seconds": total_s}
    if stages:
        # the other two stages of the scene time, single-threaded as in the reference (FilterDepthMap: one thread per view; FuseDepthMaps:
        # main thread only): FilterDepthMap of that view against 8 neighbours and FuseDepthMaps over the same 9 views, on maps derived
        # from the analytic depth (the CPU cannot estimate 9 views within the sample budget)
        try:
            nb = [int(v) for v in osc.neighbors(ref, 1)["ids"][:8]]
            rng = np.random.default_rng(7)
            for v in [ref] + nb:
                _, d, n = syn.render(v, want_bgr=False)
                dn = (d * (1 + 0.002 * rng.standard_normal(d.shape, dtype=np.float32))).astype(np.float32)
                cf = np.where(d > 0, rng.uniform(0.5, 1, d.shape).astype(np.float32), 0).astype(np.float32)
                if osc.select_views(v) <= 0:
                    continue
                osc.init_views(v, 5)
                osc.set_depthmap(v, dn, n, cf, float(d[d > 0].min() * 0.5), float(d.max() * 2))
            t0 = time.time(); osc.filter(ref, list(range(len(nb))), True); t_filter = time.time() - t0
            t0 = time.time(); cloud = osc.fuse(True, True); t_fuse = (time.time() - t0) / (1 + len(nb))
            per_view = total_s * (3.0 / iters) + t_filter + t_fuse
            out["stage_seconds_per_view"] = {"estimate": total_s * (3.0 / iters), "filter": t_filter, "fuse": t_fuse}
            out["scene_seconds_extrapolated"] = per_view * syn.n_views
            out["scene_equivalent_value"] = inner_pix * 3 * syn.n_views / (per_view * syn.n_views) / 1e6
            out["sample"] += f"; + FilterDepthMap of that view (8 neighbours) and FuseDepthMaps over {1 + len(nb)} views ({len(cloud['xyz'])} points), 1 thread each"
        except Exception as e:  # the stage timings are additional information only
            out["stage_seconds_per_view"] = f"failed: {e}"
    return out


def cpu_fusion_sample(args, views=12):
    """C5 on the CPU: FuseDepthMaps of the oracle over a bounded window of the precomputed maps (single thread, like the reference)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from hcmvs_b200.synth import SynthScene, c5_maps, frame_neighbors
    syn = SynthScene(5, args.scale, views)
    osc = O.OracleScene(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
    maps = []
    for i in range(syn.n_views):
        bgr, d, n = syn.render(i)
        osc.add_image(syn.K[i], syn.R[i], syn.Cc[i], bgr=bgr)
        maps.append(c5_maps((d, n), 900 + i))
    for i in range(syn.n_views):
        ids = frame_neighbors(i, syn.n_views)
        osc.set_neighbors(i, ids, min(5, len(ids)))
        osc.set_depthmap(i, *maps[i])
    t0 = time.time(); cloud = osc.fuse(True, True); sec = time.time() - t0
    mpix = syn.n_views * syn.width * syn.height / 1e6
    return {"value": mpix / sec, "unit": "Mpix/s", "cores": 1, "kind": "port", "seconds": sec,
            "sample": f"FuseDepthMaps over {syn.n_views} of the {args.views or 500} maps ({syn.width}x{syn.height}), {len(cloud['xyz'])} points, 1 thread (the reference fuses on its main thread)"}, syn


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be built here) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == 5:
        vals, secs, base, syn = [], [], None, None
        for s in range(args.warmup + args.steps):
            base, syn = cpu_fusion_sample(args)
            if s >= args.warmup:
                vals.append(base["value"]); secs.append(base["seconds"])
            if sum(secs) > 120:
                break
        value = float(np.mean(vals)) if vals else base["value"]
        from hcmvs_b200.synth import SynthScene
        full = SynthScene(5, args.scale, args.views)
        print(json.dumps({"impl": "reference", "metric": "FuseDepthMaps Mpix/s (input depth pixels fused per second)", "value": value, "unit": "Mpix/s", "n_gpus": args.gpus,
                          "steps": len(vals), "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3 if secs else None, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic", "config": config_dict(args, full.n_views, full.width, full.height),
                          "cpu_baseline": {**base, "value": value},
                          "e2e": {"value": value, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    syn, imgs = make_scene(args)
    vals, secs = [], []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_oracle_sample(args, syn, imgs, seconds_budget=20.0 if args.steps + args.warmup <= 3 else 8.0)
        if s >= args.warmup:
            # the metric of both arms is scene pixel-iterations over the WHOLE scene time (estimate + filter + fuse): the sample's
            # per-view stage times extrapolated to the scene; the estimate-only figure stays in cpu_baseline.estimate_only
            vals.append(base.get("scene_equivalent_value", base["value"])); secs.append(base["seconds"])
        if sum(secs) > 150:  # keep the whole run within a few minutes
            break
    value = float(np.mean(vals)) if vals else base.get("scene_equivalent_value", base["value"])
    line = {
        "impl": "reference", "metric": "PatchMatch Mpix*iter/s", "value": value, "unit": "Mpix*iter/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3 if secs else None,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, syn.n_views, syn.width, syn.height),
        "cpu_baseline": {"value": value, "unit": "Mpix*iter/s", "cores": base["cores"], "kind": "port", "sample": base["sample"], "estimate_only": base["value"],
                         **{k: base[k] for k in ("stage_seconds_per_view", "scene_seconds_extrapolated", "scene_equivalent_value") if k in base}},
        "e2e": {"value": value, "unit": "Mpix*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
def sweep_roofline(tm, P, sms, peaks, peak_src, n_sweep_launches):
    """Roofline of the dominant kernel (k_sweep): FP32-pipe bound, not HBM / tensor."""
    fp32_peak = 2.0 * 128 * sms * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    texels = (int(P.adapthalfwin) + 1) ** 2
    flops = tm["n_view_scores"] * flops_per_view_score(texels) + tm["n_smooth_terms"] * 60.0
    sweep_s = tm["ms_sweeps"] / 1e3
    achieved = flops / sweep_s / 1e12 if sweep_s > 0 else 0.0
    hbm_bytes = tm["n_pixel_iters"] * 64.0
    roofline = {
        "kernel": "k_sweep (red-black PatchMatch half-sweep)", "bound": "fp32",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak if fp32_peak else None,
        "peak_source": f"2*128 lanes*{sms} SM*sm_max_mhz ({peak_src} MEASURED_PEAKS.json clock)",
        "avg_launch_ms": tm["ms_sweeps"] / max(n_sweep_launches, 1),
        "algorithmic_flops_per_view_score": flops_per_view_score(texels),
        "hbm": {"achieved": hbm_bytes / sweep_s / 1e9 if sweep_s > 0 else 0.0, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": (hbm_bytes / sweep_s / 1e9) / peaks.get("hbm_gbs", 1.0) if sweep_s > 0 else None, "bytes_per_pixel_iter": 64},
        "traffic": ncu_traffic_per_launch(),
        "traffic_note": "DRAM bytes of one k_sweep launch (ncu --set full, profiles/r02b_ncu_k_sweep_tex.txt; iteration-3 launch of a C2 view) vs 0.96 Mpix x 64 B = 61 MB algorithmic",
        "sweep_mpix_iter_s": tm["n_pixel_iters"] / sweep_s / 1e6 if sweep_s > 0 else None,
        "hyp_per_pixel_iter": tm["n_hypotheses"] / max(tm["n_pixel_iters"], 1),
    }
    # the unit that actually bounds k_sweep: a 4-tap texture gather costs 16 cycles of L1TEX write-back per warp
    # => 2 bilinear samples / clk / SM; samples = view scores x texels
    tex_peak = 2.0 * sms * peaks.get("sm_max_mhz", 1965.0) * 1e6
    tex_rate = tm["n_view_scores"] * texels / sweep_s if sweep_s > 0 else 0.0
    roofline["tex_wall"] = {"achieved": tex_rate / 1e9, "peak": tex_peak / 1e9, "unit": "Gsample/s", "frac": tex_rate / tex_peak if tex_peak else None,
                            "note": "texture write-back 32 B/clk/SM, 16 B per bilinear sample (ncu l1tex__tex_writeback_active 80.1 % in an iteration-3 launch, profiles/r02b_ncu_k_sweep_tex.txt)"}
    return roofline


def stage_rooflines(tm, steps, npoints, hbm_peak):
    """The HBM-bound stages against the measured copy bandwidth (SURVEY §8d's algorithmic bytes; this rank's share)."""
    out = {}
    if tm["ms_filter"] > 0:
        gbs = tm["filter_bytes"] / (tm["ms_filter"] / 1e3) / 1e9  # filter_bytes and ms_filter both accumulate over the timed steps
        out["filter"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                         "bytes": "(24 n + 16) B per reference pixel, n = neighbour maps (8)"}
    if tm["ms_fuse"] > 0 and npoints:
        # per seed 27 B (depth, conf, normal, colour, claim) + 8 B per probe; per merged view 23 B; ~40 B per emitted point (last fusion)
        merged = max(int(tm.get("fuse_view_refs", 0)) - npoints, 0)
        fb = tm["fuse_seeds"] * 27.0 + tm["fuse_probes"] * 8.0 + merged * 23.0 + npoints * 40.0
        gbs = fb / (tm["ms_fuse"] / steps / 1e3) / 1e9
        out["fuse"] = {"bound": "hbm (random 32-byte sectors: latency, not bandwidth)", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                       "seeds": int(tm["fuse_seeds"]), "probes": int(tm["fuse_probes"]),
                       "bytes": "27 B per seed + 8 B per probe + 40 B per point (merged views' 23 B not counted)"}
    return out


def run_fusion_stress(args, torch, local):
    """--config 5: FuseDepthMaps alone over precomputed maps (SURVEY §8d C5 recipe). A step = one fusion of all maps; the maps are restored
    on the device between steps (fusion zeroes occluded depths in place). Single GPU: the fusion is sequential over views by definition."""
    from hcmvs_b200 import api
    from hcmvs_b200.synth import SynthScene, c5_maps, frame_neighbors
    syn = SynthScene(5, args.scale, args.views)
    V, H, W = syn.n_views, syn.height, syn.width
    ctx = api.Context(local, nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
    h2d = 0
    t_up = 0.0
    for i in range(V):
        bgr, d, n = syn.render(i)
        m = c5_maps((d, n), 900 + i)
        gray = np.zeros(d.shape, np.float32)  # fusion reads no gray image
        t0 = time.time()
        ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], gray, bgr)
        ctx.set_depthmap(i, *m)
        t_up += time.time() - t0
        h2d += bgr.nbytes + m[0].nbytes + m[1].nbytes + m[2].nbytes
    for i in range(V):
        ids = frame_neighbors(i, V)
        ctx.set_neighbors(i, ids, min(5, len(ids)))
        ctx.set_fuse_priority(i, len(ids))
    ctx.snapshot_maps()
    ctx.sync()
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    for _ in range(args.warmup):
        ctx.restore_snapshot(); ctx.fuse_depthmaps_device(True, True)
    ctx.sync(); ctx.reset_timers()
    clocks = ClockSampler(local); clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t_dev, npoints, nrefs = 0.0, 0, 0
    for _ in range(args.steps):
        ctx.restore_snapshot(); ctx.sync()
        ev0.record(lib_stream)
        npoints, nrefs = ctx.fuse_depthmaps_device(True, True)
        ev1.record(lib_stream)
        torch.cuda.synchronize()
        t_dev += ev0.elapsed_time(ev1) / 1e3
    clk = clocks.stop()
    tm = ctx.timers()
    sec = t_dev / args.steps
    mpix = V * W * H / 1e6
    peaks, peak_src = load_peaks()
    hbm_peak = peaks.get("hbm_gbs")
    merged = max(nrefs - npoints, 0)
    fb = tm["fuse_seeds"] * 27.0 + tm["fuse_probes"] * 8.0 + merged * 23.0 + npoints * 40.0
    gbs = fb / sec / 1e9
    roofline = {"kernel": "k_fuse_scene (persistent cooperative FuseDepthMaps) + k_fuse_build", "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": gbs / hbm_peak if hbm_peak else None, "peak_source": f"{peak_src} MEASURED_PEAKS.json copy bandwidth",
                "avg_launch_ms": sec * 1e3, "traffic": None,
                "bytes": "SURVEY §8d: 27 B per seed + 8 B per probe + 23 B per merged view + 40 B per point",
                "seeds": int(tm["fuse_seeds"]), "probes": int(tm["fuse_probes"]), "merged": merged, "points": npoints,
                "note": "every probe is one random 32-byte sector; the kernel is bound by the latency chain of dependent sector reads between grid barriers, not by bandwidth (profiles/r02_notes.md)"}
    # e2e: maps from HOST buffers (one set_depthmap per view) + fusion + cloud download
    e2e = None
    if not args.no_e2e:
        maps = []
        for i in range(min(V, 64)):  # bounded host memory: the first 64 views' maps are re-created and re-used round-robin
            _, d, n = syn.render(i, want_bgr=False)
            maps.append(c5_maps((d, n), 900 + i))
        ctx.restore_snapshot(); ctx.fuse_depthmaps_device(True, True); ctx.download_fused_pinned()  # sizes the arena
        torch.cuda.synchronize()
        t0 = time.time()
        up = 0
        for i in range(V):
            m = maps[i % len(maps)]
            ctx.set_depthmap(i, *m); up += m[0].nbytes + m[1].nbytes + m[2].nbytes
        ctx.fuse_depthmaps_device(True, True)
        n_, d2h = ctx.download_fused_pinned()
        torch.cuda.synchronize()
        te = time.time() - t0
        e2e = {"value": mpix / te, "unit": "Mpix/s", "h2d_bytes_per_step": int(up), "d2h_bytes_per_step": int(d2h), "seconds_per_scene": te, "points": int(n_),
               "api": "C ABI: hcmvs_set_depthmap per view from host memory, hcmvs_fuse_depthmaps, hcmvs_download_fused_pinned",
               "note": "views beyond the 64th re-use the maps of view i % 64 (host memory bound), so the e2e cloud differs from the device-timed one"}
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu, _ = cpu_fusion_sample(args)
        except Exception as e:
            cpu = {"value": None, "unit": "Mpix/s", "cores": 1, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps({
        "metric": "FuseDepthMaps Mpix/s (input depth pixels fused per second)", "value": mpix / sec, "unit": "Mpix/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
        "config": config_dict(args, V, W, H), "run": {"parallelism": "single GPU (FuseDepthMaps is sequential over views)", "l2": "20 B/px maps + 32 B/px fusion records of 500 views = 54 GB > 126 MB L2"},
        "scene_seconds": sec, "fused_points": npoints, "fuse_rounds": int(tm["n_fuse_rounds"]), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(tm["n_launches"]), "clocks": clk}))
    ctx.close()


def run_b200(args):
    import torch
    import torch.distributed as dist
    from hcmvs_b200 import api, host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — hcmvs_b200 has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if args.config == 5:
        if rank == 0:
            run_fusion_stress(args, torch, local)
        return
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    params = dict(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=args.adapthalfwin, sampler=args.sampler)
    ctx = api.Context(local, **params)
    P = ctx.params
    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    peaks, peak_src = load_peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    e2e, cpu = None, None

    if world == 1:
        # ------------------------------------------------------------------ one GPU: the C ABI driven view by view
        syn, imgs = make_scene(args)
        V, H, W = syn.n_views, syn.height, syn.width
        hs = host.HostScene.from_synth(syn, imgs)
        # host-side scene preparation (untimed for `value`): view selection, initial depth from the sparse points
        valid = [i for i in range(V) if hs.select_views(P, i) > 0]
        nbs = {i: hs.neighbors(i, 1) for i in valid}
        nall = {i: len(hs.neighbors(i, 0)["ids"]) for i in valid}
        for i in range(V):
            ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], hs.gray(i), imgs[i])
        for i in valid:
            ctx.set_neighbors(i, nbs[i]["ids"], min(5, len(nbs[i]["ids"])), nbs[i]["score"])
            ctx.set_fuse_priority(i, nall[i])
        init = {i: hs.init_depth(i) for i in valid}  # sparse splat, SceneDensify.cpp:783-808
        filtered_views = [v for v in valid if min(8, len(nbs[v]["ids"])) >= 2]
        n_est = len(valid)

        def upload_initial():
            # H2D of the rough depth maps: done before the clock starts (`value` = inputs resident in HBM)
            for v in valid:
                ctx.init_depthmap(v, init[v][0], None, init[v][1], init[v][2])
            ctx.sync()

        def hot_path():
            for v in valid:
                ctx.estimate_depthmap(v, 0, 1)
            for v in filtered_views:  # FilterDepthMap: neighbours with maps, at most 8 (SceneDensify.cpp:4117-4130)
                ctx.filter_depthmap(v, list(range(min(8, len(nbs[v]["ids"])))), adjust=True, download=False)
            ctx.commit_filtered()
            n = ctx.fuse_depthmaps_device(True, True)[0]  # the fused cloud stays in HBM; e2e below downloads it
            ctx.sync()
            return n
        parallelism = "single GPU"
    else:
        # ------------------------------------------------------------------ N GPUs: hcmvs_host's DistributedReconstruction (C++) does everything;
        # Python launches the ranks, carries the communicator id and holds the clock. Every rank renders / holds the pixels of ITS share
        # of the images only (index % world == rank): the others reach its GPU over NVLink.
        syn, imgs = make_scene(args, only=[i for i in range(_n_views(args)) if i % world == rank])
        V, H, W = syn.n_views, syn.height, syn.width
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        hs = host.HostScene.from_synth(syn, imgs)
        hs.dist_prepare(ctx, rank, world)   # selection (sharded), image uploads (1/world over PCIe + NVLink), neighbour lists, plan, initial maps
        info = hs.dist_info()
        n_est = info["n_valid"]
        upload_initial = lambda: (hs.dist_upload_initial(), ctx.sync())
        hot_path = lambda: hs.dist_run(seed=1, run_filter=True, download=False)["n_points"] if rank != 0 else _points_after(hs, ctx)
        parallelism = f"view-sharded x{world} (C++ host: hcmvs_host::DistributedReconstruction), in-place NCCL broadcasts, {info['n_split']} view(s) row-split over all ranks"

    inner = (W - 14) * (H - 14)
    pix_iters_step = inner * int(P.nEstimationIters) * n_est
    for _ in range(args.warmup):
        upload_initial()
        hot_path()
    barrier()
    ctx.reset_timers()
    clocks = ClockSampler(local)
    clocks.start()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    t_dev = 0.0
    npoints = 0
    for _ in range(args.steps):
        upload_initial()
        barrier()
        ev0.record(lib_stream)
        npoints = hot_path()
        ev1.record(lib_stream)
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / 1e3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev += float(t.item())
    clk = clocks.stop()
    tm = ctx.timers()
    sec_step = t_dev / args.steps
    value = pix_iters_step / sec_step / 1e6
    n_mine = n_est if world == 1 else (info["n_mine_whole"] + info["n_split"])
    roofline = sweep_roofline(tm, P, sms, peaks, peak_src, 2 * int(P.nEstimationIters) * n_mine * args.steps)
    stages = {k: tm[k] / args.steps for k in ("ms_prep", "ms_score", "ms_sweeps", "ms_end", "ms_filter", "ms_fuse", "ms_exchange")}
    launches = tm["n_launches"]
    roofline["stages"] = stage_rooflines(tm, args.steps, npoints, peaks.get("hbm_gbs"))

    # ---- e2e: the host-facing DenseReconstruction call with HOST buffers (uploads + downloads inside the timed region)
    if not args.no_e2e and world == 1:
        # one long-lived context (== one DepthMapsData object); every run re-uploads all images and initial maps from host memory
        ctx2 = api.Context(local, **params)
        host.HostScene.from_synth(syn, imgs).dense_reconstruction(ctx2, seed=1, run_filter=True)  # warm-up (allocations)
        ts, st = [], None
        for _ in range(max(1, min(args.steps, 3))):
            hs2 = host.HostScene.from_synth(syn, imgs)
            hs2.pin_images()  # inputs in page-locked host memory, as the bench contract asks
            torch.cuda.synchronize()
            t0 = time.time()
            st = hs2.dense_reconstruction(ctx2, seed=1, run_filter=True)  # returns with the fused cloud in host memory (Scene::pointcloud)
            torch.cuda.synchronize()
            ts.append(time.time() - t0)
            hs2.close()
        with_dmaps = None
        if args.e2e_dmap_dir:
            os.makedirs(args.e2e_dmap_dir, exist_ok=True)
            td = []
            for _ in range(2):
                hs2 = host.HostScene.from_synth(syn, imgs)
                torch.cuda.synchronize()
                t0 = time.time()
                sd = hs2.dense_reconstruction(ctx2, seed=1, run_filter=True, dmap_dir=args.e2e_dmap_dir)
                torch.cuda.synchronize()
                td.append(time.time() - t0)
                hs2.close()
            with_dmaps = {"seconds_per_scene": float(td[-1]), "d2h_bytes_per_step": sd["d2h_bytes"], "dir": args.e2e_dmap_dir,
                          "note": "maps read back through the page-locked download slots and written by a host thread while later views are estimated"}
        ctx2.close()
        # the same call with the reference's DEFAULT initialisation (nMinViewsTrustPoint = 2: the sparse points are triangulated on the
        # host and rasterised on the device) — reported next to the splat start that both arms of this bench time
        default_init = None
        if args.config == 2:
            try:
                p3 = dict(params); p3["nMinViewsTrustPoint"] = 2
                ctx3 = api.Context(local, **p3)
                host.HostScene.from_synth(syn, imgs).dense_reconstruction(ctx3, seed=1, run_filter=True)
                hs3 = host.HostScene.from_synth(syn, imgs)
                torch.cuda.synchronize()
                t0 = time.time()
                s3 = hs3.dense_reconstruction(ctx3, seed=1, run_filter=True)
                torch.cuda.synchronize()
                t3 = time.time() - t0
                hs3.close(); ctx3.close()
                default_init = {"value": pix_iters_step / t3 / 1e6, "seconds_per_scene": t3, "points": s3["n_points"], "h2d_bytes_per_step": s3["h2d_bytes"],
                                "note": "nMinViewsTrustPoint = 2 (InitDepthMap / TriangulatePoints2DepthMap): host Delaunay + device rasteriser"}
            except Exception as e:  # informational only
                default_init = {"value": None, "note": f"failed: {e}"}
        e2e = {"value": pix_iters_step / float(np.mean(ts)) / 1e6, "unit": "Mpix*iter/s", "h2d_bytes_per_step": st["h2d_bytes"],
               "d2h_bytes_per_step": st["d2h_bytes"], "seconds_per_scene": float(np.mean(ts)), "points": st["n_points"],
               "seconds": {k: round(float(st[k]), 4) for k in ("sec_select", "sec_upload", "sec_estimate", "sec_filter", "sec_fuse")},
               "api": "hcmvs_host.DenseReconstruction (select views, upload, estimate, filter, fuse, download cloud)"}
        if with_dmaps:
            e2e["with_dmaps"] = with_dmaps
        if default_init:
            e2e["default_init"] = default_init
    elif not args.no_e2e:
        # N > 1: ONE call per rank of the same C++ host function, hcmvs_host.DenseReconstructionDistributed, with HOST buffers inside the
        # timed region: every rank selects 1/N of the views, uploads 1/N of the images over PCIe (the rest arrives over NVLink), its
        # initial maps; rank 0 fuses and downloads the cloud. Wall clock between barriers, max over ranks.
        def e2e_step():
            hs2 = host.HostScene.from_synth(syn, imgs)
            hs2.pin_images()  # inputs in page-locked host memory, as the bench contract asks
            barrier()
            t0 = time.time()
            st_ = hs2.dense_reconstruction_distributed(ctx, rank, world, seed=1, run_filter=True)
            barrier()
            dt = time.time() - t0
            if os.environ.get("HCMVS_DIST_DEBUG"):
                print(f"[bench rank {rank}] e2e step {dt:.3f} s", file=sys.stderr)
            hs2.close()
            return dt, st_
        e2e_step()  # warm-up: sizes the page-locked arena
        ts, st = [], None
        for _ in range(max(1, min(args.steps, 3))):
            dt, st = e2e_step()
            ts.append(dt)
        tt = torch.tensor([float(np.mean(ts))], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        io = torch.tensor([float(st["h2d_bytes"]), float(st["d2h_bytes"]), float(st["n_points"])], device=dev, dtype=torch.float64)
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
        e2e = {"value": pix_iters_step / float(tt.item()) / 1e6, "unit": "Mpix*iter/s", "h2d_bytes_per_step": int(io[0].item()), "d2h_bytes_per_step": int(io[1].item()),
               "seconds_per_scene": float(tt.item()), "points": int(io[2].item()),
               "seconds_rank0": {k: round(float(st[k]), 4) for k in ("sec_select", "sec_estimate", "sec_filter", "sec_fuse")},
               "api": "hcmvs_host.DenseReconstructionDistributed on every rank (sharded view selection + all-gather, 1/N of the images over PCIe and the rest "
                      "over NVLink, estimate / exchange / filter, fuse + cloud download on rank 0); wall clock between barriers, max over ranks, all ranks' bytes summed"}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_oracle_sample(args, syn, imgs)
        except Exception as e:  # the oracle is a reported baseline, never a dependency of the product path
            cpu = {"value": None, "unit": "Mpix*iter/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": "PatchMatch Mpix*iter/s", "value": value, "unit": "Mpix*iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, V, W, H),
            "run": {"parallelism": parallelism, "valid_views": n_est,
                    "l2": f"inputs per view (5 neighbour images + maps, ~{(6 * 4 + 40) * W * H / 1e6:.0f} MB) re-read per launch; scene working set {V * 44 * W * H / 1e9:.1f} GB > 126 MB L2"},
            "scene_seconds": sec_step, "fused_points": npoints, "fuse_rounds": int(tm["n_fuse_rounds"]), "stage_ms_per_step_rank0": stages,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def _n_views(args):
    from hcmvs_b200.synth import SynthScene
    s = SynthScene(args.config, args.scale, args.views)
    n = s.n_views
    s.close()
    return n


def _points_after(hs, ctx):
    """rank 0 of a distributed run: the cloud stays in HBM; its size comes from the context."""
    hs.dist_run(seed=1, run_filter=True, download=False)
    return ctx.fused_counts()[0]


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
