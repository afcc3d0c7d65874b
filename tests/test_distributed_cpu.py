"""world_size-2 gloo test of the multi-rank host plumbing (the N>1 path of bench.py): view sharding + map all-gather.
The CUDA library is replaced here by per-rank numpy 'maps' — the collective, the slot layout and the import/export
order are exactly those the GPU run uses (hcmvs_b200.shard.exchange_maps)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _expected(view, h, w):
    rng = np.random.default_rng(1000 + view)
    return rng.standard_normal((h, w, 4)).astype(np.float32), rng.uniform(0, 1, (h, w)).astype(np.float32)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from hcmvs_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w = 6, 9
    valid = list(range(11)); nall = {v: 12 - (v % 4) for v in valid}
    plan = shard.make_plan(valid, nall, world)
    have = {v: _expected(v, h, w) for v in plan.views_of(rank)}      # this rank "estimated" only its own views
    send_dn = torch.zeros((plan.slots, h, w, 4)); send_cf = torch.zeros((plan.slots, h, w))
    recv_dn = torch.zeros((world, plan.slots, h, w, 4)); recv_cf = torch.zeros((world, plan.slots, h, w))

    def export_fn(v, s):
        send_dn[s] = torch.from_numpy(have[v][0]); send_cf[s] = torch.from_numpy(have[v][1])

    def import_fn(v, r, s):
        have[v] = (recv_dn[r, s].numpy().copy(), recv_cf[r, s].numpy().copy())

    shard.exchange_maps(plan, rank, send_dn, send_cf, recv_dn, recv_cf, export_fn, import_fn, lambda: None, dist)
    ok = sorted(have) == valid and all(np.array_equal(have[v][0], _expected(v, h, w)[0]) and np.array_equal(have[v][1], _expected(v, h, w)[1]) for v in valid)
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    q.put((rank, ok, float(t.item())))
    dist.destroy_process_group()


def test_two_rank_map_exchange_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), "a rank did not end up with every view's maps"
    assert all(r[2] == 2.0 for r in res)


def test_owner_arrays_are_consistent_across_ranks():
    """hcmvs_exchange_maps takes one owner list that every rank must build identically: owner[v] = rank estimating view v."""
    sys.path.insert(0, ROOT)
    from hcmvs_b200 import shard
    valid = [0, 1, 2, 4, 5, 7, 8, 9, 10]
    nall = {v: 12 - (v % 5) for v in valid}
    for world in (1, 2, 4, 8):
        plan = shard.make_plan(valid, nall, world)
        owner = plan.owner_array(11)
        assert owner.dtype == np.int32 and list(np.where(owner < 0)[0]) == [3, 6]
        for r in range(world):
            assert sorted(plan.views_of(r)) == sorted(int(v) for v in np.where(owner == r)[0])
        counts = np.bincount(owner[owner >= 0], minlength=world)
        assert counts.max() - counts.min() <= 1                       # round-robin in connection order balances the ranks
        sub = plan.owner_array(11, only={0, 5, 9})
        assert set(np.where(sub >= 0)[0]) == {0, 5, 9} and all(sub[v] == owner[v] for v in (0, 5, 9))


def _worker_split(rank, world, port, q):
    """The lib-NCCL schedule of bench.py on CPU tensors: whole rounds broadcast per view from their owner, the views of the incomplete
    last round estimated in row bands by every rank and assembled by in-place broadcasts of the bands (HCMVS_OWNER_SPLIT_ROWS)."""
    sys.path.insert(0, ROOT)
    from hcmvs_b200 import shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    h, w = 13, 5                                                      # 13 rows over 2 ranks: bands of 6 and 7 rows
    valid = list(range(7)); nall = {v: 9 - (v % 3) for v in valid}
    plan = shard.make_plan(valid, nall, world, split_rows=True)
    maps = {v: torch.zeros((h, w, 4)) for v in valid}                 # every rank holds a buffer for every view
    for v in plan.whole_views_of(rank):
        maps[v] = torch.from_numpy(_expected(v, h, w)[0])
    r0, r1 = plan.rows_of(rank, h)
    for v in plan.split_views():
        maps[v][r0:r1] = torch.from_numpy(_expected(v, h, w)[0])[r0:r1]   # only this rank's band is valid
        maps[v][:r0] = -7.0; maps[v][r1:] = -7.0                           # halo by-products: must be overwritten by the owners
    def exchange(owner):
        for v in range(len(owner)):
            if owner[v] == shard.OWNER_SPLIT_ROWS:
                for r in range(world):
                    a, b = plan.rows_of(r, h)
                    band = maps[v][a:b].contiguous()
                    dist.broadcast(band, src=r)
                    maps[v][a:b] = band
            elif owner[v] >= 0:
                dist.broadcast(maps[v], src=int(owner[v]))
    for own in plan.round_owner_arrays(len(valid)):
        exchange(own)
    exchange(plan.split_owner_array(len(valid)))
    ok = all(np.array_equal(maps[v].numpy(), _expected(v, h, w)[0]) for v in valid)
    q.put((rank, ok, len(plan.split_views()), (r0, r1)))
    dist.destroy_process_group()


def test_two_rank_row_split_exchange_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_split, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "a rank did not end up with every view's full maps"
    assert all(r[2] == 1 for r in res) and [r[3] for r in res] == [(0, 6), (6, 13)]   # 7 views on 2 ranks: 3 whole rounds + 1 split view
