#!/usr/bin/env python3
"""N-GPU check of the optional voxel-block merge over NCCL (plvs_b200/parallel.py):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/merge_2gpu_check.py
Every rank maps the same synthetic world from its own poses, the ranks merge, and the union of the owned partitions is
compared with the weighted sum of all per-rank maps gathered on rank 0."""
import os, pathlib, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
from plvs_b200 import parallel, synth, tsdf as T      # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
w, h = 160, 120
K = synth.intrinsics(w, h)
p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=0)
g = T.ChiselServer(p, device=local); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
for f in (rank * 3, rank * 3 + 1, rank * 3 + 4):
    g.integrate(synth.depth_frame(f, w, h), synth.pose(f))
before = g.download()
sent, recv = parallel.merge_maps(g)
after = g.download()
own = parallel.owner_of(torch.from_numpy(after[0]), world) if len(after[0]) else torch.zeros(0)
ok = bool((own == rank).all())
gathered = [None] * world
dist.all_gather_object(gathered, (before[0], before[1], before[2], after[0], after[1], after[2]))
if rank == 0:
    from tests.merge_expect import fold, compare
    exp = fold([(bk, bs, bw) for (bk, bs, bw, _, _, _) in gathered])     # the owner receives the sources in rank order
    ak = np.concatenate([g_[3] for g_ in gathered]); as_ = np.concatenate([g_[4] for g_ in gathered]); aw = np.concatenate([g_[5] for g_ in gathered])
    disjoint = len({tuple(k) for k in ak}) == len(ak)
    bad = compare(exp, ak, as_, aw, atol=2e-6)
    print(f"merge over NCCL, world {world}: {len(exp)} distinct blocks; partitions disjoint: {disjoint}; mismatching blocks: {bad}; owner rule held here: {ok}")
    print("PASS" if bad == 0 and ok and disjoint else "FAIL")
dist.barrier()
dist.destroy_process_group()
