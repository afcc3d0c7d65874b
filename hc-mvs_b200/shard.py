"""View sharding plan for N ranks (one process per GPU) — a Python VIEW of the C++ plan.

The schedule itself lives in the C++ host (hcmvs_host::MakeShardPlan / DistributedReconstruction, hc-mvs_b200/host/densify_dist.cpp):
make_plan() below asks the host library for the order and the row-split decision, so the CPU tests of this module test the C++ logic.
exchange_maps() is the torch.distributed slot exchange used by the gloo tests only (the GPU path broadcasts in place through
hcmvs_exchange_maps).

The scene shards by reference view: every rank holds all images and cameras, estimates the
depth maps of its own views, and the ranks all-gather the (normal, depth) / confidence maps so that
filtering and fusion see every neighbour. Views are dealt round-robin in FuseDepthMaps' connection order
(decreasing scored-neighbour count, ties by index — SceneDensify.cpp:3286-3303) which balances the load
and keeps neighbouring views on different ranks.
"""
from dataclasses import dataclass
from typing import Dict, List


OWNER_SPLIT_ROWS = -2         # HCMVS_OWNER_SPLIT_ROWS (include/hcmvs_b200.h)


@dataclass
class ShardPlan:
    order: List[int]          # views in connection order
    world: int
    slots: int                # exchange slots per rank
    split_rows: bool = False  # the views left over after the full rounds are estimated in row bands by ALL ranks

    def whole_rounds(self) -> int:
        return len(self.order) // self.world if self.split_rows else self.slots

    def split_views(self) -> List[int]:
        """Views of the last, incomplete round: with split_rows every rank estimates the band rows_of(rank, H) of each of them."""
        return self.order[self.whole_rounds() * self.world:] if self.split_rows else []

    def rows_of(self, rank: int, height: int):
        return rank * height // self.world, (rank + 1) * height // self.world

    def whole_views_of(self, rank: int) -> List[int]:
        n = self.whole_rounds() * self.world
        return [v for k, v in enumerate(self.order[:n]) if k % self.world == rank]

    def split_owner_array(self, n_views: int):
        import numpy as np
        o = np.full(n_views, -1, np.int32)
        for v in self.split_views():
            o[v] = OWNER_SPLIT_ROWS
        return o

    def owner(self, k: int) -> int:
        return k % self.world

    def slot(self, k: int) -> int:
        return k // self.world

    def views_of(self, rank: int) -> List[int]:
        return [v for k, v in enumerate(self.order) if k % self.world == rank]

    def owner_array(self, n_views: int, only=None):
        """owner[i] = rank that estimates view i, -1 for views outside the plan (or outside `only`): the list every rank hands to
        hcmvs_exchange_maps."""
        import numpy as np
        o = np.full(n_views, -1, np.int32)
        for k, v in enumerate(self.order):
            if only is None or v in only:
                o[v] = k % self.world
        return o

    def round_owner_arrays(self, n_views: int):
        """One owner list per estimation round s (the s-th view of every rank): the exchange of round s overlaps round s+1."""
        import numpy as np
        out = []
        for s in range(self.whole_rounds()):
            o = np.full(n_views, -1, np.int32)
            for k in range(s*self.world, min((s+1)*self.world, len(self.order))):
                o[self.order[k]] = k % self.world
            out.append(o)
        return out

    def location(self) -> Dict[int, tuple]:
        """view -> (rank, slot) of its maps in the gathered buffer."""
        return {v: (k % self.world, k // self.world) for k, v in enumerate(self.order)}


def make_plan(valid_views, n_scored_neighbors, world, split_rows=False):
    """split_rows: 49 views on 8 ranks are 6 full rounds + 1 view; instead of one rank estimating a 7th view while seven wait, every
    rank estimates one eighth of its rows (+ the halo its dependencies reach, hcmvs_estimate_depthmap_rows) and the bands are
    exchanged in place. Filtering of those views stays with owner(k) = k % world."""
    from . import host
    import numpy as np
    valid = [int(v) for v in valid_views]
    n_views = (max(valid) + 1) if valid else 1
    ns = np.zeros(n_views, np.uint32)
    for v in valid:
        ns[v] = int(n_scored_neighbors[v])
    p = host.shard_plan(valid, ns, world, split_rows=bool(split_rows))
    order = [int(v) for v in p["order"]]
    slots = (len(order) + world - 1) // world if order else 0
    return ShardPlan(order=order, world=world, slots=slots, split_rows=p["n_split"] > 0)


def exchange_maps(plan, rank, send_dn, send_cf, recv_dn, recv_cf, export_fn, import_fn, sync_fn, dist, post_sync=None):
    """All-gather the ranks' maps. export_fn(view, slot) fills the send buffers, import_fn(view, rank, slot)
    consumes a received slot; `dist` is torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""
    mine = plan.views_of(rank)
    for s, v in enumerate(mine):
        export_fn(v, s)
    sync_fn()
    dist.all_gather_into_tensor(recv_dn.view(-1), send_dn.view(-1))
    dist.all_gather_into_tensor(recv_cf.view(-1), send_cf.view(-1))
    if post_sync:
        post_sync()
    for v, (r, s) in plan.location().items():
        if r != rank:
            import_fn(v, r, s)
    sync_fn()
