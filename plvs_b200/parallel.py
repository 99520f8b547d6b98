"""Multi-GPU plumbing.  The hot path shards by camera stream (one process per GPU, no collective on the data
path -- SURVEY.md §8e, bench.py).  The single optional exchange is the global voxel-block merge: every block key
is owned by rank `hash(key) mod R`; ranks pack their live blocks as (key, w*sdf, w), route them to the owners with
an all-to-all (torch.distributed: NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests) and the owner folds
them with the commutative weighted sum (plvs_tsdf_merge_packed -- a CUDA kernel, not torch).
The reference has nothing comparable (single process)."""
import ctypes as C
import numpy as np
import torch
import torch.distributed as dist

VOX = 4096


def owner_of(keys, world):
    """Teschner hash of the chunk id (the same primes the map's hash table uses) modulo the world size."""
    k = keys.to(torch.int64)
    h = ((k[:, 0] * 73856093) ^ (k[:, 1] * 19349663) ^ (k[:, 2] * 83492791)) & 0xFFFFFFFF
    return (h % world).to(torch.int64)


def exchange_blocks(keys, wsdf, w, group=None):
    """Route packed blocks to their owners.  keys [n,3] int32, wsdf/w [n,4096] float32 (same device).
    Returns the blocks this rank owns, from all ranks (a key may appear several times: one per source rank)."""
    world = dist.get_world_size(group)
    own = owner_of(keys, world) if len(keys) else torch.zeros(0, dtype=torch.int64, device=keys.device)
    order = torch.argsort(own, stable=True)
    keys, wsdf, w = keys[order].contiguous(), wsdf[order].contiguous(), w[order].contiguous()
    send_counts = torch.bincount(own, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    n_recv = int(sum(rc))
    rk = torch.empty((n_recv, 3), dtype=keys.dtype, device=keys.device)
    rs = torch.empty((n_recv, VOX), dtype=wsdf.dtype, device=keys.device)
    rw = torch.empty((n_recv, VOX), dtype=w.dtype, device=keys.device)
    dist.all_to_all_single(rk, keys, output_split_sizes=rc, input_split_sizes=sc, group=group)
    dist.all_to_all_single(rs, wsdf, output_split_sizes=rc, input_split_sizes=sc, group=group)
    dist.all_to_all_single(rw, w, output_split_sizes=rc, input_split_sizes=sc, group=group)
    return rk, rs, rw


def merge_maps(server, group=None, device=None):
    """Global voxel-block merge of the per-rank TSDF maps (plvs_b200.tsdf.ChiselServer).  Afterwards each rank
    holds exactly the blocks it owns, fused over all ranks.  Returns (#blocks sent, #blocks received).
    `device`: where the exchange buffers live -- the current CUDA device by default (NCCL); the CPU tests run the library on the
    CPU execution model of tests/native/cuda_emu.hpp, whose "device" memory is host memory, over gloo."""
    lib, h = server._lib, server._h
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    n = C.c_int()
    rc = lib.plvs_tsdf_export_packed(h, None, None, None, 0, C.byref(n))
    assert rc == 0
    n = n.value
    keys = torch.empty((n, 3), dtype=torch.int32, device=dev)
    wsdf = torch.empty((n, VOX), dtype=torch.float32, device=dev)
    w = torch.empty((n, VOX), dtype=torch.float32, device=dev)
    if n:
        m = C.c_int()
        rc = lib.plvs_tsdf_export_packed(h, C.c_void_p(keys.data_ptr()), C.c_void_p(wsdf.data_ptr()), C.c_void_p(w.data_ptr()), n, C.byref(m))
        assert rc == 0 and m.value == n
    rk, rs, rw = exchange_blocks(keys, wsdf, w, group)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    server.Reset()
    if len(rk):
        rc = lib.plvs_tsdf_merge_packed(h, C.c_void_p(rk.data_ptr()), C.c_void_p(rs.data_ptr()), C.c_void_p(rw.data_ptr()), len(rk))
        assert rc == 0, lib.plvs_last_error()
    return n, len(rk)
