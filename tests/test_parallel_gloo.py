"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path -- stream sharding (one stream per rank) and
the owner routing of the optional voxel-block merge."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plvs_b200 import parallel, synth
    rng = np.random.default_rng(100 + rank)
    # overlapping key sets: both ranks see chunks 0..39, each has 20 private ones
    shared = np.stack(np.meshgrid(np.arange(-2, 3), np.arange(0, 4), np.arange(5, 7), indexing="ij"), -1).reshape(-1, 3)
    private = rng.integers(-50, 50, (20, 3)) + (1000 * (rank + 1))
    keys = torch.from_numpy(np.concatenate([shared, private]).astype(np.int32))
    w = torch.from_numpy(rng.random((len(keys), 4096)).astype(np.float32))
    wsdf = w * 0.01 * (rank + 1)
    rk, rp = parallel.exchange_blocks(keys, torch.cat([wsdf, w], 1))
    rs, rw = rp[:, :4096], rp[:, 4096:]
    own = parallel.owner_of(rk, world)
    ok_owner = bool((own == rank).all())
    tot_sent = torch.tensor([float(w.double().sum()), float(len(keys))], dtype=torch.float64)
    tot_recv = torch.tensor([float(rw.double().sum()), float(len(rk))], dtype=torch.float64)
    dist.all_reduce(tot_sent); dist.all_reduce(tot_recv)
    # shared keys arrive twice at their owner (once per source rank)
    uniq, counts = np.unique(rk.numpy(), axis=0, return_counts=True)
    dup_ok = set(counts.tolist()) <= {1, 2} and int((counts == 2).sum()) == int((parallel.owner_of(torch.from_numpy(shared.astype(np.int32)), world) == rank).sum())
    # stream sharding: every rank generates a different stream from the documented seed rule
    img = synth.gray_frame(0, 160, 120, stream=rank)
    sig = torch.tensor([float(img.astype(np.float64).sum())], dtype=torch.float64)
    sigs = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sigs, sig)
    ret[rank] = dict(ok_owner=ok_owner, conserved=bool(torch.allclose(tot_sent, tot_recv)), dup_ok=bool(dup_ok),
                     streams_differ=len({float(s) for s in sigs}) == world)
    dist.destroy_process_group()


def test_block_routing_and_stream_sharding_world2():
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        assert ret[r] == dict(ok_owner=True, conserved=True, dup_ok=True, streams_differ=True), ret[r]


def test_owner_hash_is_stable():
    from plvs_b200 import parallel
    k = torch.tensor([[0, 0, 0], [1, 2, 3], [-1, -2, -3], [100, -7, 42]], dtype=torch.int32)
    o8 = parallel.owner_of(k, 8)
    assert o8.tolist() == [0, ((1 * 73856093) ^ (2 * 19349663) ^ (3 * 83492791)) % 8,
                           (((-1 * 73856093) ^ (-2 * 19349663) ^ (-3 * 83492791)) & 0xFFFFFFFF) % 8,
                           (((100 * 73856093) ^ (-7 * 19349663) ^ (42 * 83492791)) & 0xFFFFFFFF) % 8]



def _merge_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    from plvs_b200 import _lib as ABI, parallel, synth, tsdf as T
    from oracle import tsdf as OT
    from tests.native_build import build_emulated_library
    from tests.test_emulated_kernels import _Partial
    from tests.merge_expect import fold, compare
    ABI._lib = ABI.declare(_Partial(C.CDLL(build_emulated_library())))          # this process runs the library on the CPU model
    w, h = 160, 120
    K = synth.intrinsics(w, h)
    p = T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1)
    frames = ((0, 1, 2), (2, 5, 9))                                              # the two ranks see the same world from different poses
    g = T.ChiselServer(p); g.SetDepthCameraInfo(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
    for f in frames[rank]:
        g.integrate(synth.depth_frame(f, w, h), synth.pose(f), synth.bgr_frame(f, w, h))
    # a pool too small for the blocks this rank will own: refused BEFORE the map is touched
    tiny = T.ChiselServer(T.default_params(voxel_resolution=0.04, use_carving=1, near_plane=0.1, far_plane=4.0, max_blocks=4096, use_color=1))
    before = g.download()
    g.params.max_blocks, keep = 4, g.params.max_blocks
    try:
        parallel.merge_maps(g, device=torch.device("cpu"))
        refused = False
    except RuntimeError:
        refused = True
    g.params.max_blocks = keep
    after = g.download()
    untouched = refused and all(np.array_equal(a, b) for a, b in zip(before, after))
    del tiny
    rep = {}
    sent, received = parallel.merge_maps(g, device=torch.device("cpu"), report=rep)
    ck, cs, cw, cc = g.download()
    # expectation: both ranks' maps from the oracle, folded in rank order, restricted to the blocks this rank owns
    host = []
    for fr in frames:
        o = OT.Map(p, threads=4); o.set_camera(K["fx"], K["fy"], K["cx"], K["cy"], w, h)
        for f in fr:
            o.integrate(synth.depth_frame(f, w, h), synth.pose(f), synth.bgr_frame(f, w, h))
        host.append(o.download())
    exp = fold(host)
    mine = {k: v for k, v in exp.items() if int(parallel.owner_of(torch.tensor([k], dtype=torch.int32), world)[0]) == rank}
    shared = set(map(tuple, host[0][0])) & set(map(tuple, host[1][0]))
    ret[rank] = dict(bad=compare(mine, ck, cs, cw, cc), n=len(ck), expected=len(mine), sent=sent, received=received, shared=len(shared), untouched=untouched,
                     bytes_ok=rep["sent_bytes"] == sent * (12 + 3 * 4096 * 4) and rep["exchange_s"] > 0, coloured=int((cc[..., 3] > 0).sum()))
    dist.destroy_process_group()


def test_merge_maps_world2_with_the_library_on_the_cpu_model():
    """plvs_b200.parallel.merge_maps end to end in two processes over gloo: export kernel -> owner routing -> all-to-all -> merge kernels, with the library's
    translation units on the CPU execution model (tests/native/cuda_emu.hpp); every rank ends up with exactly the blocks it owns, fused as the
    host-side fold of the two oracle maps says"""
    from tests.native_build import build_emulated_library
    build_emulated_library()                                                   # once, before the workers race for it
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_merge_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    total = 0
    for r in range(world):
        assert ret[r]["bad"] == 0 and ret[r]["n"] == ret[r]["expected"] > 10, ret[r]
        assert ret[r]["untouched"] and ret[r]["bytes_ok"] and ret[r]["coloured"] > 1000, ret[r]
        assert ret[r]["shared"] > 10
        total += ret[r]["n"]
    assert sum(ret[r]["sent"] for r in range(world)) == sum(ret[r]["received"] for r in range(world)) >= total


def _time_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plvs_b200 import parallel
    mine = 10.0 + 2.5 * rank                                  # rank 1 is the slow one
    job, per_rank = parallel.job_time(torch.tensor([mine], dtype=torch.float32))
    ret[rank] = dict(job=job, per_rank=per_rank, stream_same=parallel.stream_of_rank(rank), stream_distinct=parallel.stream_of_rank(rank, "distinct"))
    dist.destroy_process_group()


def test_bench_timing_rule_world2():
    """bench.py's N > 1 rule on two gloo ranks: the job's time is the MAX over the ranks' own times, every rank reports the same list; by default every
    rank processes the same synthetic stream (identical work per GPU), `distinct` gives rank r stream r"""
    world, port = 2, _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_time_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r]["job"] == 12.5 and ret[r]["per_rank"] == [10.0, 12.5], ret[r]
        assert ret[r]["stream_same"] == 0 and ret[r]["stream_distinct"] == r


def test_bench_timing_rule_single_process():
    from plvs_b200 import parallel
    assert parallel.job_time(torch.tensor([3.25])) == (3.25, [3.25])
