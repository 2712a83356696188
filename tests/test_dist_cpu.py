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


def test_peer_setup_agrees_across_ranks_or_falls_back(monkeypatch):
    """ShardedCEMOptimizer._setup_peers (peer-memory exchange): every rank must reach the SAME decision.  The CUDA / IPC calls
    and the process group are mocked (no GPU here): (1) all ranks allocate and map -> peer path with world pointers; (2) one
    rank fails to map a peer's buffer -> everybody closes what it opened and keeps the collective."""
    import ctypes as C

    import torch
    from mbrl_lib_b200 import dist as bd

    world = 2

    class FakeLib:
        def __init__(self, fail_open):
            self.fail_open, self.closed, self.opened = fail_open, [], 0

        def b200pets_peer_buffer_bytes(self, *a):
            return 4096

        def b200pets_peer_alloc(self, nbytes, ptr_ref, handle):
            C.cast(ptr_ref, C.POINTER(C.c_void_p))[0] = 0x1000
            return 0

        def b200pets_peer_open(self, hb, ptr_ref):
            if self.fail_open:
                return 7
            self.opened += 1
            C.cast(ptr_ref, C.POINTER(C.c_void_p))[0] = 0x2000
            return 0

        def b200pets_peer_close(self, ptr, owned):
            self.closed.append((getattr(ptr, "value", ptr), owned))
            return 0

        def b200pets_last_error(self):
            return b"mock failure"

    def make(rank, fail_open, peer_ok_flags):
        opt = bd.ShardedCEMOptimizer.__new__(bd.ShardedCEMOptimizer)
        opt.peer_exchange, opt._peer, opt.world, opt.rank, opt.group = True, None, world, rank, None
        opt.population_size, opt.local_population, opt.elite_num = 1000, 500, 100
        opt.device = torch.device("cpu")
        opt.lib = FakeLib(fail_open)
        calls = {"n": 0}

        def all_gather_object(out, obj, group=None):
            calls["n"] += 1
            for r in range(world):
                if calls["n"] == 1:  # (hostname, handle, ok)
                    out[r] = obj if r == rank else ("samehost", b"h" * 64, True)
                else:  # second round: every rank's "all my peers mapped" flag
                    out[r] = obj if r == rank else peer_ok_flags
        monkeypatch.setattr(bd.dist, "is_initialized", lambda: True)
        monkeypatch.setattr(bd.dist, "get_backend", lambda g=None: "nccl")
        monkeypatch.setattr(bd.dist, "all_gather_object", all_gather_object)
        monkeypatch.setattr(bd.dist, "barrier", lambda group=None: None)
        monkeypatch.setattr(bd.socket if hasattr(bd, "socket") else __import__("socket"), "gethostname", lambda: "samehost")
        monkeypatch.setattr(torch.cuda, "device", lambda d: __import__("contextlib").nullcontext())
        monkeypatch.setattr(bd._lib, "check", lambda rc, what="": (_ for _ in ()).throw(RuntimeError(what)) if rc else None)
        return opt

    ok = make(0, fail_open=False, peer_ok_flags=True)
    peer = ok._setup_peers(180)
    assert peer is not None and peer["ptrs"][0] == 0x1000 and peer["ptrs"][1] == 0x2000 and ok.peer_exchange

    mine_fails = make(0, fail_open=True, peer_ok_flags=True)
    assert mine_fails._setup_peers(180) is None and mine_fails.peer_exchange is False
    assert (0x1000, 1) in mine_fails.lib.closed  # its own buffer was released

    other_fails = make(1, fail_open=False, peer_ok_flags=False)
    assert other_fails._setup_peers(180) is None and other_fails.peer_exchange is False
    assert (0x2000, 0) in other_fails.lib.closed and (0x1000, 1) in other_fails.lib.closed
