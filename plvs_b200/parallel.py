"""Multi-GPU plumbing.  The hot path shards by camera stream (one process per GPU, no collective on the data
path -- SURVEY.md §8e, bench.py).  The single optional exchange is the global voxel-block merge: every block key
is owned by rank `hash(key) mod R`; ranks pack their live blocks as (key, w*sdf, w, rgba), route them to the owners with
an all-to-all (torch.distributed: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests) and the owner folds
them (plvs_tsdf_merge_packed_rgba -- CUDA kernels, not torch): distances with the commutative weighted sum, colours as
ColorVoxel::Integrate would fold the incoming colour weight (ColorVoxel.h:34-127).
The reference has nothing comparable (single process)."""
import ctypes as C
import time
import numpy as np
import torch
import torch.distributed as dist

VOX = 4096
ITEM_WORDS = 3 * VOX          # one packed block travels as 3 x 4096 32-bit words: w*sdf (f32), w (f32), rgba (u8x4)


def owner_of(keys, world):
    """Teschner hash of the chunk id (the same primes the map's hash table uses) modulo the world size."""
    k = keys.to(torch.int64)
    h = ((k[:, 0] * 73856093) ^ (k[:, 1] * 19349663) ^ (k[:, 2] * 83492791)) & 0xFFFFFFFF
    return (h % world).to(torch.int64)


def exchange_blocks(keys, payload, group=None):
    """Route packed blocks to their owners.  keys [n,3] int32, payload [n, k] (any 32-bit dtype; same device).
    Returns the blocks this rank owns, from all ranks (a key may appear several times: one per source rank, in rank order)."""
    world = dist.get_world_size(group)
    own = owner_of(keys, world) if len(keys) else torch.zeros(0, dtype=torch.int64, device=keys.device)
    order = torch.argsort(own, stable=True)
    keys, payload = keys[order].contiguous(), payload[order].contiguous()
    send_counts = torch.bincount(own, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    n_recv = int(sum(rc))
    rk = torch.empty((n_recv, 3), dtype=keys.dtype, device=keys.device)
    rp = torch.empty((n_recv, payload.shape[1]), dtype=payload.dtype, device=keys.device)
    dist.all_to_all_single(rk, keys, output_split_sizes=rc, input_split_sizes=sc, group=group)
    dist.all_to_all_single(rp, payload, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return rk, rp


def merge_maps(server, group=None, device=None, report=None):
    """Global voxel-block merge of the per-rank TSDF maps (plvs_b200.tsdf.ChiselServer).  Afterwards each rank
    holds exactly the blocks it owns, fused over all ranks (distance, weight and -- for a colour map -- the colour state).
    Returns (#blocks sent, #blocks received).  The rank's map is only replaced once the incoming blocks are known to fit its pool;
    otherwise an error is raised and the map is left as it was.
    `device`: where the exchange buffers live -- the current CUDA device by default (NCCL); the CPU tests run the library on the
    CPU execution model of tests/native/cuda_emu.hpp, whose "device" memory is host memory, over gloo.
    `report`: optional dict, filled with payload bytes and the seconds the all-to-all took (for a bandwidth figure)."""
    lib, h = server._lib, server._h
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    n = C.c_int()
    rc = lib.plvs_tsdf_export_packed_rgba(h, None, None, None, None, 0, C.byref(n))
    assert rc == 0, lib.plvs_last_error()
    n = n.value
    keys = torch.empty((n, 3), dtype=torch.int32, device=dev)
    payload = torch.empty((n, ITEM_WORDS), dtype=torch.int32, device=dev)       # [w*sdf | w | rgba], bit patterns
    if n:
        m = C.c_int()
        base = payload.data_ptr()
        # the three planes of an item are strided inside one row: export into contiguous planes, then interleave once
        planes = torch.empty((3, n, VOX), dtype=torch.int32, device=dev)
        rc = lib.plvs_tsdf_export_packed_rgba(h, C.c_void_p(keys.data_ptr()), C.c_void_p(planes[0].data_ptr()), C.c_void_p(planes[1].data_ptr()),
                                              C.c_void_p(planes[2].data_ptr()), n, C.byref(m))
        assert rc == 0 and m.value == n, lib.plvs_last_error()
        payload.view(n, 3, VOX).copy_(planes.permute(1, 0, 2))
        del planes, base
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    rk, rp = exchange_blocks(keys, payload, group)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    if report is not None:
        report.update(sent_blocks=n, received_blocks=len(rk), sent_bytes=n * (12 + 4 * ITEM_WORDS), received_bytes=len(rk) * (12 + 4 * ITEM_WORDS),
                      exchange_s=dt)
    n_unique = int(torch.unique(rk, dim=0).shape[0]) if len(rk) else 0
    if n_unique > server.params.max_blocks:
        raise RuntimeError(f"merge_maps: {n_unique} blocks are owned by this rank but its pool holds {server.params.max_blocks}; map left unchanged")
    server.Reset()
    if len(rk):
        planes = rp.view(len(rk), 3, VOX).permute(1, 0, 2).contiguous()
        rc = lib.plvs_tsdf_merge_packed_rgba(h, C.c_void_p(rk.data_ptr()), C.c_void_p(planes[0].data_ptr()), C.c_void_p(planes[1].data_ptr()),
                                             C.c_void_p(planes[2].data_ptr()), len(rk))
        assert rc == 0, lib.plvs_last_error()
    return n, len(rk)


def job_time(t_ms, group=None):
    """The multi-rank timing rule of bench.py: t_ms = this rank's own device time of the timed region (a one-element tensor on the rank's device: CUDA
    under NCCL, CPU under gloo).  Returns (job time = MAX over ranks, [every rank's own time]) -- the same on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return float(t_ms.item()), [float(t_ms.item())]
    each = [torch.zeros_like(t_ms) for _ in range(world)]
    dist.all_gather(each, t_ms, group=group)
    mx = t_ms.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return float(mx.item()), [float(x.item()) for x in each]


def stream_of_rank(rank, rank_streams="same"):
    """which synthetic stream a rank processes: weak scaling gives every GPU identical work (stream 0); `distinct` gives rank r stream r"""
    return rank if rank_streams == "distinct" else 0
