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
             3: "C3 synthetic ETH3D-shaped 20 views 6048x4032", 4: "C4 synthetic video 300 views 1920x1080"}


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


def make_scene(args):
    from hcmvs_b200.synth import SynthScene
    syn = SynthScene(args.config, args.scale, args.views)
    imgs = [syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views)]
    return syn, imgs


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_sweep launch from the committed `ncu --set full` summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_k_sweep_final.txt")
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
    osc = O.OracleScene(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
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


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference itself cannot be built here) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    syn, imgs = make_scene(args)
    vals, secs = [], []
    base = None
    for s in range(args.warmup + args.steps):
        base = cpu_oracle_sample(args, syn, imgs, seconds_budget=20.0 if args.steps + args.warmup <= 3 else 8.0, stages=(s == args.warmup + args.steps - 1))
        if s >= args.warmup:
            vals.append(base["value"]); secs.append(base["seconds"])
        if sum(secs) > 150:  # keep the whole run within a few minutes
            break
    value = float(np.mean(vals)) if vals else base["value"]
    line = {
        "impl": "reference", "metric": "PatchMatch Mpix*iter/s", "value": value, "unit": "Mpix*iter/s", "n_gpus": args.gpus,
        "steps": len(vals), "warmup": args.warmup, "ms_per_step": float(np.mean(secs)) * 1e3 if secs else None,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "scale": args.scale, "views": syn.n_views, "neighbours": 5, "patchmatch_iters": 3},
        "cpu_baseline": {"value": value, "unit": "Mpix*iter/s", "cores": base["cores"], "kind": "port", "sample": base["sample"],
                         **{k: base[k] for k in ("stage_seconds_per_view", "scene_seconds_extrapolated", "scene_equivalent_value") if k in base}},
        "e2e": {"value": value, "unit": "Mpix*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ B200 arm
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    syn, imgs = make_scene(args)
    V, H, W = syn.n_views, syn.height, syn.width
    params = dict(nNumViews=5, nEstimationIters=3, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5, sampler=args.sampler)
    ctx = api.Context(local, **params)
    P = ctx.params
    hs = host.HostScene.from_synth(syn, imgs)
    # ---- host-side scene preparation (untimed for `value`): view selection, initial depth from the sparse points
    valid = [i for i in range(V) if hs.select_views(P, i) > 0]
    nbs = {i: hs.neighbors(i, 1) for i in valid}
    nall = {i: len(hs.neighbors(i, 0)["ids"]) for i in valid}
    for i in range(V):
        ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], hs.gray(i), imgs[i])
    for i in valid:
        ctx.set_neighbors(i, nbs[i]["ids"], min(5, len(nbs[i]["ids"])), nbs[i]["score"])
        ctx.set_fuse_priority(i, nall[i])
    # initial depth maps (sparse splat, SceneDensify.cpp:783-808) computed once on the host
    init = {i: hs.init_depth(i) for i in valid}
    # views are dealt round-robin in fusion (connection) order — SURVEY §8(e)
    from hcmvs_b200 import shard
    use_lib_nccl = world > 1 and args.exchange == "nccl"
    plan = shard.make_plan(valid, nall, world, split_rows=use_lib_nccl and not args.no_split_rows)
    order, mine = plan.order, plan.views_of(rank)
    mine_whole, split_views = plan.whole_views_of(rank), plan.split_views()
    inner = (W - 14) * (H - 14)
    pix_iters_step = inner * int(P.nEstimationIters) * len(valid)

    lib_stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    filtered_views = {v for v in valid if min(8, len(nbs[v]["ids"])) >= 2}
    if use_lib_nccl:
        # the library's own NCCL communicator: rank 0 creates the id, torch.distributed only carries its 128 bytes
        ids = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0], rank, world)
        owner_rounds = plan.round_owner_arrays(V)
        owner_split = plan.split_owner_array(V)
        owner_filtered = plan.owner_array(V, only=filtered_views)
    elif world > 1:
        # exchange buffers (one slot per view, replicated on every rank)
        send_dn = torch.zeros((plan.slots, H, W, 4), dtype=torch.float32, device=dev)
        send_cf = torch.zeros((plan.slots, H, W), dtype=torch.float32, device=dev)
        recv_dn = torch.zeros((world, plan.slots, H, W, 4), dtype=torch.float32, device=dev)
        recv_cf = torch.zeros((world, plan.slots, H, W), dtype=torch.float32, device=dev)

    def exchange():
        """torch path: all-gather every rank's (normal, depth) and confidence maps over NCCL; import the others' maps."""
        if world == 1:
            return
        shard.exchange_maps(
            plan, rank, send_dn, send_cf, recv_dn, recv_cf,
            export_fn=lambda v, s: ctx.export_maps_d(v, send_dn[s].data_ptr(), send_cf[s].data_ptr()),
            import_fn=lambda v, r, s: ctx.import_maps_d(v, recv_dn[r, s].data_ptr(), recv_cf[r, s].data_ptr(), init[v][1], init[v][2]),
            sync_fn=ctx.sync, dist=dist, post_sync=torch.cuda.synchronize)

    def upload_initial():
        # H2D of the rough depth maps: done before the clock starts (`value` = inputs resident in HBM)
        for v in sorted(set(mine) | set(split_views)):  # every rank estimates a band of the row-split views
            ctx.init_depthmap(v, init[v][0], None, init[v][1], init[v][2])
        ctx.sync()

    def hot_path():
        if use_lib_nccl:
            # round s: every rank estimates its s-th view, then those views are broadcast in place (one NCCL group) on the
            # communication stream while round s+1 is being estimated
            for s_, own in enumerate(owner_rounds):
                if s_ < len(mine_whole):
                    ctx.estimate_depthmap(mine_whole[s_], 0, 1)
                ctx.exchange_maps(own, 0, overlap=True)
            # the views of the incomplete last round: every rank estimates its band of rows, the bands are broadcast in place
            for v in split_views:
                r0, r1 = plan.rows_of(rank, H)
                ctx.estimate_depthmap_rows(v, r0, r1, 0, 1)
            if split_views:
                ctx.exchange_maps(owner_split, 0, overlap=True)
            ctx.exchange_wait()
        else:
            for v in mine:
                ctx.estimate_depthmap(v, 0, 1)
            exchange()
        # FilterDepthMap: neighbours with maps, at most 8 (SceneDensify.cpp:4117-4130)
        for v in mine:
            if v in filtered_views:
                ctx.filter_depthmap(v, list(range(min(8, len(nbs[v]["ids"])))), adjust=True, download=False)
        if use_lib_nccl:
            ctx.exchange_maps(owner_filtered, 1)     # the pending filter output (8 B/px); committed on every rank below
            ctx.commit_filtered()
        else:
            ctx.commit_filtered()
            exchange()
        n = 0
        if rank == 0:
            n = ctx.fuse_depthmaps_device(True, True)[0]  # the fused cloud stays in HBM; e2e below downloads it
        ctx.sync()
        return n

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

    # ---- roofline of the dominant kernel (k_sweep): FP32-pipe bound, not HBM / tensor
    peaks, peak_src = load_peaks()
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    fp32_peak = 2.0 * 128 * sms * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
    texels = (int(P.adapthalfwin) + 1) ** 2
    flops = tm["n_view_scores"] * flops_per_view_score(texels) + tm["n_smooth_terms"] * 60.0
    sweep_s = tm["ms_sweeps"] / 1e3
    n_sweep_launches = 2 * int(P.nEstimationIters) * len(mine) * args.steps
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
        "traffic_note": "DRAM bytes of one k_sweep launch (ncu --set full, profiles/r01_ncu_k_sweep_final.txt) vs 0.96 Mpix x 64 B = 61 MB algorithmic",
        "sweep_mpix_iter_s": tm["n_pixel_iters"] / sweep_s / 1e6 if sweep_s > 0 else None,
        "hyp_per_pixel_iter": tm["n_hypotheses"] / max(tm["n_pixel_iters"], 1),
    }
    # the unit that actually bounds k_sweep: a 4-tap texture gather costs 16 cycles of L1TEX write-back per warp
    # => 2 bilinear samples / clk / SM; samples = view scores x texels
    tex_peak = 2.0 * sms * peaks.get("sm_max_mhz", 1965.0) * 1e6
    tex_rate = tm["n_view_scores"] * texels / sweep_s if sweep_s > 0 else 0.0
    roofline["tex_wall"] = {"achieved": tex_rate / 1e9, "peak": tex_peak / 1e9, "unit": "Gsample/s", "frac": tex_rate / tex_peak if tex_peak else None,
                            "note": "texture write-back 32 B/clk/SM, 16 B per bilinear sample (ncu l1tex__tex_writeback_active 79.4 %, profiles/r01_ncu_k_sweep_final.txt)"}
    stages = {k: tm[k] / args.steps for k in ("ms_prep", "ms_score", "ms_sweeps", "ms_end", "ms_filter", "ms_fuse", "ms_exchange")}
    launches = tm["n_launches"]
    # ---- the HBM-bound stages against the measured copy bandwidth (SURVEY §8d's algorithmic bytes; rank 0's share)
    hbm_peak = peaks.get("hbm_gbs")
    stage_rooflines = {}
    if tm["ms_filter"] > 0:
        gbs = tm["filter_bytes"] / (tm["ms_filter"] / 1e3) / 1e9  # filter_bytes and ms_filter both accumulate over the timed steps
        stage_rooflines["filter"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                                     "bytes": "(24 n + 16) B per reference pixel, n = neighbour maps (8)"}
    if tm["ms_fuse"] > 0 and npoints:
        # per seed 27 B (depth, conf, normal, colour, claim) + 8 B per probe; per merged view 23 B; ~40 B per emitted point (last fusion)
        merged = max(int(tm.get("fuse_view_refs", 0)) - npoints, 0)
        fb = tm["fuse_seeds"] * 27.0 + tm["fuse_probes"] * 8.0 + merged * 23.0 + npoints * 40.0
        gbs = fb / (tm["ms_fuse"] / args.steps / 1e3) / 1e9
        stage_rooflines["fuse"] = {"bound": "hbm (gather latency)", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak if hbm_peak else None,
                                   "seeds": int(tm["fuse_seeds"]), "probes": int(tm["fuse_probes"]),
                                   "bytes": "27 B per seed + 8 B per probe + 40 B per point (merged views' 23 B not counted)"}
    roofline["stages"] = stage_rooflines

    # ---- e2e: the host-facing DenseReconstruction call with HOST buffers (uploads + downloads inside the timed region)
    e2e = None
    if not args.no_e2e and world == 1:
        # one long-lived context (== one DepthMapsData object); every run re-uploads all images and initial maps from host memory
        ctx2 = api.Context(local, **params)
        host.HostScene.from_synth(syn, imgs).dense_reconstruction(ctx2, seed=1, run_filter=True)  # warm-up (allocations)
        ts, st = [], None
        for _ in range(max(1, min(args.steps, 3))):
            hs2 = host.HostScene.from_synth(syn, imgs)
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
        e2e["default_init"] = default_init
    elif world > 1 and not args.no_e2e and use_lib_nccl:
        # N > 1: the same sharded job driven through the C ABI with HOST buffers inside the timed region — every rank re-uploads all
        # images (each rank holds every image) and its initial maps from page-locked host memory, rank 0 downloads the fused cloud
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        grays = [pin(hs.gray(i)) for i in range(V)]
        bgrs = [pin(imgs[i]) for i in range(V)]
        inits = {v: pin(init[v][0]) for v in sorted(set(mine) | set(split_views))}
        h2d = sum(g.nbytes + b.nbytes for g, b in zip(grays, bgrs)) + sum(a.nbytes for a in inits.values())

        def e2e_step():
            for i in range(V):
                ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], grays[i], bgrs[i])
            for v, a in inits.items():
                ctx.init_depthmap(v, a, None, init[v][1], init[v][2])
            n_ = hot_path()
            return ctx.download_fused_pinned() if rank == 0 else (n_, 0)

        e2e_step()  # warm-up: sizes the page-locked arena
        ts, res = [], (0, 0)
        for _ in range(max(1, min(args.steps, 3))):
            barrier()
            t0 = time.time()
            res = e2e_step()
            barrier()
            ts.append(time.time() - t0)
        tt = torch.tensor([float(np.mean(ts))], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        hh = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
        dist.all_reduce(hh, op=dist.ReduceOp.SUM)
        e2e = {"value": pix_iters_step / float(tt.item()) / 1e6, "unit": "Mpix*iter/s", "h2d_bytes_per_step": int(hh.item()), "d2h_bytes_per_step": int(res[1]),
               "seconds_per_scene": float(tt.item()), "points": int(res[0]),
               "api": "C ABI per rank: hcmvs_set_view (all images) + hcmvs_init_depthmap from page-locked host memory, estimate / exchange / filter, "
                      "fuse + hcmvs_download_fused_pinned on rank 0; max over ranks, all ranks' uploads summed"}
    elif world > 1:
        e2e = {"value": None, "unit": "Mpix*iter/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "note": "measured with --exchange nccl only"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cpu = cpu_oracle_sample(args, syn, imgs)
        except Exception as e:  # the oracle is a reported baseline, never a dependency of the product path
            cpu = {"value": None, "unit": "Mpix*iter/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}

    if rank == 0:
        line = {
            "metric": "PatchMatch Mpix*iter/s", "value": value, "unit": "Mpix*iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "scale": args.scale, "views": V, "image": [W, H], "neighbours": 5,
                       "patchmatch_iters": int(P.nEstimationIters), "stages": "estimate(A+B+C) + filter + fuse", "parallelism": f"view-sharded x{world}" + (f", map exchange: {args.exchange}" if world > 1 else "") + (f", {len(split_views)} view(s) row-split over all ranks" if split_views else ""),
                       "l2": "inputs per view (5 neighbour images + maps, ~77 MB) re-read per launch; 49-view working set 2.3 GB > 126 MB L2"},
            "scene_seconds": sec_step, "fused_points": npoints, "fuse_rounds": int(tm["n_fuse_rounds"]), "stage_ms_per_step_rank0": stages,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
