"""World-size-2 gloo tests of the population-sharded CEM host logic: shard bounds, the
single all-gather of local top-k records per iteration, and that the merged record set yields the same elite set
on every rank as a top-k over the union population."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mbrl_lib_b200 import dist as bd

    N, dims, elite = 64, 6, 10
    lo, hi = bd.shard_bounds(N, rank, world)
    g = torch.Generator().manual_seed(123)
    pop = torch.randn(N, dims, generator=g)
    vals = torch.randn(N, generator=g)
    k = bd.records_per_rank(elite, hi - lo)
    lv, li = vals[lo:hi].topk(k)  # stands in for b200pets_cem_local_topk on this rank's shard
    records = torch.cat([lv[:, None], pop[lo:hi][li]], dim=1).contiguous()
    merged = bd.gather_records(records)  # the one collective of an iteration
    assert merged.shape == (world * k, 1 + dims)
    mv, mi = merged[:, 0].topk(elite)
    uv, ui = vals.topk(elite)
    ok = torch.allclose(mv, uv) and torch.allclose(merged[mi, 1:], pop[ui])
    gathered = [torch.zeros_like(merged) for _ in range(world)]
    dist.all_gather(gathered, merged)
    same = all(torch.equal(gathered[0], x) for x in gathered)
    out[rank] = bool(ok and same)
    dist.destroy_process_group()


def test_sharded_cem_record_exchange_world2():
    world = 2
    mgr = mp.get_context("spawn").Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_shard_bounds_cover_population():
    from mbrl_lib_b200 import dist as bd

    for total in (500, 501, 7, 64000):
        for world in (1, 2, 3, 8):
            spans = [bd.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert bd.records_per_rank(50, 500) == 50 and bd.records_per_rank(6554, 4000) == 4000
