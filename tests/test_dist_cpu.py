"""N>1 host logic on CPU (gloo, world_size 2): the sharding of independent gates/expressions across
ranks, the one-time parameter broadcast, and the key replication protocol itself (broadcast_cloud_key: sizes from the
library, copy-in on the source rank, two broadcasts, adopt on the receivers) with host-memory stand-ins for the
device arrays."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    m = g.load_package()
    from ieache_b200 import dist as idist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    p = m.Params.default(630) if rank == 0 else m.Params()
    p = idist.broadcast_params(p, src=0)
    lo, hi = idist.shard_range(1 << 20, rank, world)
    # a stand-in for the per-rank device work: every rank "processes" its shard, max time is reduced
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    q.put((rank, p.n, p.bk_l, p.ks_stdev, lo, hi, t.item()))
    dist.destroy_process_group()


class _HostKey:
    """stand-in for CloudKey on the source rank: the key arrays live in host memory"""

    def __init__(self, bk, ks):
        self.bk, self.ks = bk, ks

    def device_arrays(self):
        return self.bk.ctypes.data, self.bk.nbytes, self.ks.ctypes.data, self.ks.nbytes


class _HostEngine:
    """stand-in for Engine: device_copy is a memmove, cloud_key_adopt records what it was given"""

    def device_copy(self, dst, src, nbytes):
        import ctypes
        ctypes.memmove(dst, src, nbytes)

    def cloud_key_adopt(self, params, bk_ptr, ks_ptr):
        return ("adopted", params.n, bk_ptr, ks_ptr)


def _key_worker(rank, world, port, q):
    import ctypes

    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    m = g.load_package()
    from ieache_b200 import dist as idist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 12                                                     # a small LWE dimension keeps the arrays at a few MB
    p = m.Params.default(n) if rank == 0 else m.Params()
    key = None
    if rank == 0:
        bkb, ksb = ctypes.c_size_t(), ctypes.c_size_t()
        assert m.lib().ieache_cloudkey_device_sizes(ctypes.byref(p), ctypes.byref(bkb), ctypes.byref(ksb)) == 0
        rng = np.random.default_rng(5)
        key = _HostKey(rng.standard_normal(bkb.value // 8), rng.integers(-2 ** 31, 2 ** 31, ksb.value // 4, dtype=np.int64).astype(np.int32))
    timings = {}
    got, keep = idist.broadcast_cloud_key(_HostEngine(), key, p, src=0, timings=timings)
    bk_t, ks_t = keep
    q.put((rank, got[0] if rank else "source", float(bk_t.double().sum()), int(ks_t.long().sum()), bk_t.numel(), ks_t.numel(),
           timings["bytes"], timings["broadcast_ms"] > 0, (got[2], got[3]) == (bk_t.data_ptr(), ks_t.data_ptr()) if rank else True))
    dist.destroy_process_group()


def test_gloo_world2_cloud_key_replication():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_key_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    src, dst = res
    assert dst[1] == "adopted" and dst[8]                       # the receiver wraps exactly the tensors it received into
    assert src[2] == dst[2] and src[3] == dst[3]                # same contents on both ranks
    assert src[4] == dst[4] == 12 * 6 * 2 * 512 * 2 and src[5] == dst[5] == 1024 * 8 * 3 * 632
    assert src[6] == dst[6] == src[4] * 8 + src[5] * 4 and src[7] and dst[7]


def test_shard_range_partitions_exactly():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from ieache_b200.dist import shard_range
    for total in (0, 1, 7, 1 << 20, 65536 * 3 + 1):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_param_broadcast_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [630, 630] and [r[2] for r in res] == [3, 3]
    assert res[0][3] == 2.0 ** -15 == res[1][3]
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 1 << 19, 1 << 19, 1 << 20)
    assert res[0][6] == float(1 << 20)
