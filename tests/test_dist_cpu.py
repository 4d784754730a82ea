"""N>1 host logic on CPU (gloo, world_size 2): the sharding of independent gates/expressions across
ranks and the one-time parameter broadcast that precedes the key broadcast."""
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
