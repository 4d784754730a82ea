"""Multi-GPU plumbing: the path shards by independent gates / expressions, every GPU
holds a replica of the cloud key, and the only collective is the one-time broadcast of that key.

`torch.distributed` is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of `total` independent units owned by `rank` (weak or strong scaling alike)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_params(params, src: int = 0):
    """Broadcast an ieache Params struct (as bytes) from rank `src`."""
    import ctypes

    import torch
    import torch.distributed as dist

    raw = np.frombuffer(bytes(params), dtype=np.uint8).copy()
    t = torch.from_numpy(raw)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = t.to(dev)
    dist.broadcast(t, src=src)
    out = type(params).from_buffer_copy(t.cpu().numpy().tobytes())
    assert ctypes.sizeof(out) == len(raw)
    return out


def broadcast_cloud_key(engine, key, params, src: int = 0, timings: dict | None = None):
    """One-time replication of the device-resident cloud key: rank `src` owns `key`; every other rank allocates
    tensors of the same size, receives the transform-domain BK and the packed KSK with two broadcasts (NCCL over
    NVLink on the GPU box), and wraps them with ieache_cloudkey_adopt_device.  Returns (CloudKey, keepalive).

    `timings` (optional) receives broadcast_ms (device time of the two broadcasts, max over ranks; the communicator is
    assumed to be up — bench.py times its bring-up separately) and bytes.  Under the gloo backend (CPU tests) the
    tensors live in host memory and `engine` / `key` may be stand-ins with the same methods: the protocol is the
    same."""
    import ctypes
    import time

    import torch
    import torch.distributed as dist

    from . import lib

    rank = dist.get_rank()
    on_gpu = dist.get_backend() == "nccl"
    params = broadcast_params(params, src)
    bkb, ksb = ctypes.c_size_t(), ctypes.c_size_t()
    rc = lib().ieache_cloudkey_device_sizes(ctypes.byref(params), ctypes.byref(bkb), ctypes.byref(ksb))
    if rc:
        raise RuntimeError(lib().ieache_last_error().decode())
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    bk_t = torch.empty(bkb.value // 8, dtype=torch.float64, device=dev)
    ks_t = torch.empty(ksb.value // 4, dtype=torch.int32, device=dev)
    if rank == src:
        bk_ptr, bk_bytes, ks_ptr, ks_bytes = key.device_arrays()
        assert bk_bytes == bkb.value and ks_bytes == ksb.value
        engine.device_copy(bk_t.data_ptr(), bk_ptr, bk_bytes)
        engine.device_copy(ks_t.data_ptr(), ks_ptr, ks_bytes)
    if on_gpu:
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    t0 = time.perf_counter()
    dist.broadcast(bk_t, src=src)
    dist.broadcast(ks_t, src=src)
    if on_gpu:
        ev1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    else:
        ms = torch.tensor([1e3 * (time.perf_counter() - t0)], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if timings is not None:
        timings["broadcast_ms"] = float(ms.item())
        timings["bytes"] = bkb.value + ksb.value
    if rank == src:
        return key, (bk_t, ks_t)
    return engine.cloud_key_adopt(params, bk_t.data_ptr(), ks_t.data_ptr()), (bk_t, ks_t)
