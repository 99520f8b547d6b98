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
p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
g = T.ChiselServer(p, device=local); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
for f in (rank * 3, rank * 3 + 1, rank * 3 + 4):
    g.integrate(synth.depth_frame(f, w, h), synth.pose(f), synth.bgr_frame(f, w, h))
before = g.download()
sent, recv = parallel.merge_maps(g)
after = g.download()
own = parallel.owner_of(torch.from_numpy(after[0]), world) if len(after[0]) else torch.zeros(0)
ok = bool((own == rank).all())
gathered = [None] * world
dist.all_gather_object(gathered, (before[0], before[1], before[2], after[0], after[1], after[2], before[3], after[3]))
if rank == 0:
    from tests.merge_expect import fold, compare
    exp = fold([(g_[0], g_[1], g_[2], g_[6]) for g_ in gathered])     # the owner receives the sources in rank order
    ak = np.concatenate([g_[3] for g_ in gathered]); as_ = np.concatenate([g_[4] for g_ in gathered]); aw = np.concatenate([g_[5] for g_ in gathered])
    ac = np.concatenate([g_[7] for g_ in gathered])
    disjoint = len({tuple(k) for k in ak}) == len(ak)
    bad = compare(exp, ak, as_, aw, ac, atol=2e-6)
    print(f"merge over NCCL, world {world}: {len(exp)} distinct blocks; partitions disjoint: {disjoint}; mismatching blocks: {bad}; owner rule held here: {ok}")
    print("PASS" if bad == 0 and ok and disjoint else "FAIL")
# bandwidth of the exchange at a map of realistic size: VGA scans at 1 cm (thousands of 48 KiB blocks per rank)
w2, h2 = 640, 480
K2 = synth.intrinsics(w2, h2)
p2 = T.default_params(voxel_resolution=0.01, use_carving=1, near_plane=0.1, far_plane=5.0, max_blocks=32768, use_color=1)
g2 = T.ChiselServer(p2, device=local); g2.SetDepthCameraInfo(K2["fx"], K2["fy"], K2["cx"], K2["cy"], w2, h2)
for f in (rank * 2, rank * 2 + 1):
    g2.integrate(synth.depth_frame(f, w2, h2), synth.pose(f), synth.bgr_frame(f, w2, h2))
g2.stats()
rep = {}
parallel.merge_maps(g2, report=rep)
reps = [None] * world
dist.all_gather_object(reps, rep)
if rank == 0:
    import json
    tot = sum(r["sent_bytes"] for r in reps); t = max(r["exchange_s"] for r in reps)
    print(json.dumps({"merge_exchange": {"world": world, "blocks_sent_per_rank": [r["sent_blocks"] for r in reps], "payload_bytes_total": tot,
                                          "exchange_s_max_over_ranks": t, "aggregate_GBps": tot / t / 1e9,
                                          "note": "two torch.distributed.all_to_all_single calls (keys, then 3 x 16 KiB planes per block) over NCCL; wall clock with device synchronisation on both sides; first call of the process (includes NCCL channel set-up)"}}))
    rep2 = {}
dist.barrier()
# second exchange of the same size: NCCL warmed up
g3 = T.ChiselServer(p2, device=local); g3.SetDepthCameraInfo(K2["fx"], K2["fy"], K2["cx"], K2["cy"], w2, h2)
for f in (rank * 2 + 8, rank * 2 + 9):
    g3.integrate(synth.depth_frame(f, w2, h2), synth.pose(f), synth.bgr_frame(f, w2, h2))
g3.stats()
rep = {}
parallel.merge_maps(g3, report=rep)
reps = [None] * world
dist.all_gather_object(reps, rep)
if rank == 0:
    import json
    tot = sum(r["sent_bytes"] for r in reps); t = max(r["exchange_s"] for r in reps)
    print(json.dumps({"merge_exchange_warm": {"world": world, "payload_bytes_total": tot, "exchange_s_max_over_ranks": t, "aggregate_GBps": tot / t / 1e9}}))
dist.barrier()
dist.destroy_process_group()
