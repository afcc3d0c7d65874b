"""Run under torchrun on N GPUs: the fused cloud of the sharded run must equal the single-GPU cloud (rank 0 computes both)."""
import os, sys, hashlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from hcmvs_b200 import api, host, shard
from hcmvs_b200.synth import SynthScene

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
syn = SynthScene(2, 0.5, 13)  # 13 views: every world size leaves an incomplete last round -> row-split views
imgs = [syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views)]
params = dict(nNumViews=5, nEstimationIters=2, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
hs = host.HostScene.from_synth(syn, imgs)
plan_info = []


def run(ctx, world, rank, use_comm):
    P = ctx.params
    V = syn.n_views
    valid = [i for i in range(V) if hs.select_views(P, i) > 0]
    nbs = {i: hs.neighbors(i, 1) for i in valid}; nall = {i: len(hs.neighbors(i, 0)["ids"]) for i in valid}
    for i in range(V):
        ctx.set_view(i, syn.K[i], syn.R[i], syn.Cc[i], hs.gray(i), imgs[i])
    for i in valid:
        ctx.set_neighbors(i, nbs[i]["ids"], min(5, len(nbs[i]["ids"])), nbs[i]["score"]); ctx.set_fuse_priority(i, nall[i])
    plan = shard.make_plan(valid, nall, world, split_rows=use_comm)
    mine = plan.views_of(rank)
    mine_whole, split_views = plan.whole_views_of(rank), plan.split_views()
    filt = {v for v in valid if min(8, len(nbs[v]["ids"])) >= 2}
    if use_comm:  # per-round exchange overlapped with the next round's estimation, as bench.py does
        for s_, own in enumerate(plan.round_owner_arrays(V)):
            if s_ < len(mine_whole):
                v = mine_whole[s_]
                d, lo, hi = hs.init_depth(v)
                ctx.init_depthmap(v, d, None, lo, hi); ctx.estimate_depthmap(v, 0, 1)
            ctx.exchange_maps(own, 0, overlap=True)
        for v in split_views:  # the incomplete last round: every rank estimates its band of rows of every such view
            d, lo, hi = hs.init_depth(v)
            ctx.init_depthmap(v, d, None, lo, hi)
            r0, r1 = plan.rows_of(rank, syn.height)
            ctx.estimate_depthmap_rows(v, r0, r1, 0, 1)
        if split_views:
            ctx.exchange_maps(plan.split_owner_array(V), 0, overlap=True)
        ctx.exchange_wait()
    else:
        for v in mine:
            d, lo, hi = hs.init_depth(v)
            ctx.init_depthmap(v, d, None, lo, hi); ctx.estimate_depthmap(v, 0, 1)
    for v in mine:
        if v in filt: ctx.filter_depthmap(v, list(range(min(8, len(nbs[v]["ids"])))), True, download=False)
    if use_comm: ctx.exchange_maps(plan.owner_array(V, only=filt), 1)
    ctx.commit_filtered()
    global plan_info
    plan_info = plan.split_views()
    return ctx.fuse_depthmaps(True, True) if rank == 0 else None


ctx = api.Context(local, **params)
ids = [api.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.comm_init(ids[0], rank, world)
sharded = run(ctx, world, rank, True)
n_split = len(plan_info)
t = ctx.timers()
if rank == 0:
    single = run(api.Context(local, **params), 1, 0, False)
    same = all(np.array_equal(sharded[k], single[k]) for k in ("xyz", "views", "weights", "colors", "normals", "n_views"))
    print(f"world {world}: {n_split} row-split view(s); {len(sharded['xyz'])} points, identical to the single-GPU cloud: {same}, exchange {t['ms_exchange']:.2f} ms", flush=True)
    assert same and len(sharded["xyz"]) > 100000
dist.barrier()
dist.destroy_process_group()
