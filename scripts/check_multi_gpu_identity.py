"""Run under torchrun on N GPUs: the fused cloud of hcmvs_host.DenseReconstructionDistributed (the C++ multi-GPU host: sharded view
selection, 1/N of the images uploaded per rank and the rest received over NVLink, whole rounds + row-split views, filter on the owners,
fusion on rank 0) must equal the cloud of the single-GPU hcmvs_host.DenseReconstruction bit for bit (rank 0 computes both)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from hcmvs_b200 import api, host
from hcmvs_b200.synth import SynthScene

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
syn = SynthScene(2, 0.5, 13)  # 13 views: every world size leaves an incomplete last round -> row-split views
mine = {i: syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views) if i % world == rank}  # this rank's pixels only
params = dict(nNumViews=5, nEstimationIters=2, nEstimationIters_external=1, nMinViewsTrustPoint=1, adapthalfwin=5)
for trust in (1, 2):  # splat + random start, and the reference's default triangulated start
    params["nMinViewsTrustPoint"] = trust
    ctx = api.Context(local, **params)
    ids = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)
    hs = host.HostScene.from_synth(syn, mine)
    st = hs.dense_reconstruction_distributed(ctx, rank, world, seed=1, run_filter=True)
    info = None
    if rank == 0:
        sharded = hs.cloud()
        sharded = {k: v.copy() for k, v in sharded.items()}
        allimgs = [syn.render(i, want_depth=False, want_normal=False)[0] for i in range(syn.n_views)]
        ctx1 = api.Context(local, **params)
        hs1 = host.HostScene.from_synth(syn, allimgs)
        hs1.dense_reconstruction(ctx1, seed=1, run_filter=True)
        single = hs1.cloud()
        same = all(np.array_equal(sharded[k], single[k]) for k in ("xyz", "views", "weights", "colors", "normals", "n_views"))
        t = ctx.timers()
        print(f"world {world}, nMinViewsTrustPoint {trust}: {len(sharded['xyz'])} points, identical to the single-GPU cloud: {same}; "
              f"h2d of rank 0 {st['h2d_bytes'] / 1e6:.1f} MB (all {syn.n_views} images are {syn.n_views * syn.width * syn.height * 7 / 1e6:.1f} MB)", flush=True)
        assert same and len(sharded["xyz"]) > 100000
        hs1.close(); ctx1.close()
    dist.barrier()
    hs.close(); ctx.close()
dist.destroy_process_group()
