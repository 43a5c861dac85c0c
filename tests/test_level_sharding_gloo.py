"""Multi-rank host logic on the CPU: two gloo ranks run the same radix operators on the mock backend with
level sharding enabled (each rank bootstraps half of every wide level, all-gather completes it) and must
decrypt the same results as a single rank.  The GPU path uses the identical slicing with NCCL."""
import os
import sys

import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    import ctypes as C
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mock_radix import MockRadix
    from fhe_sign_b200.distributed import make_all_gather
    M = MockRadix()
    buf = torch.zeros(4096 * 2049, dtype=torch.int64)
    cb = make_all_gather(buf)
    rc = M.L.fsc_set_level_exchange(M.ctx, rank, world, 8, C.c_void_p(buf.data_ptr()), buf.numel() * 8, cb, None)
    assert rc == 0
    x, y, z = 0xDEADBEEFCAFEF00D, 0xFEDCBA98FFFFFFFF, 0x0123456789ABCDEF
    a, b = M.enc(x, 32), M.enc(y, 32)
    res = {
        "mul": M.dec(a * b) == (x * y) % 2**64,
        "add": M.dec(a + b) == (x + y) % 2**64,
        "wide": M.dec(M.api.mul_wide(a, b, 64)) == x * y,
        "div": M.dec(M.enc(z, 32) // 5) == z // 5,
        "narrow": M.dec(M.enc(7, 4) + M.enc(9, 4)) == 16,
        "violations": M.counters()["violations"] == 0,
        "sharded": M.api.sharded_levels() > 0,
    }
    dist.barrier()
    q.put((rank, res))
    dist.destroy_process_group()


def test_two_rank_level_sharding_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in out:
        assert all(res.values()), (rank, res)


def test_shard_ranges_cover_every_request():
    # the slicing rule of radix.h shard_range, restated
    for count in (1, 7, 8, 9, 100, 1184, 16384):
        for world in (1, 2, 4, 8):
            per = (count + world - 1) // world
            seen = []
            for r in range(world):
                lo = min(count, r * per); hi = min(count, lo + per)
                seen += list(range(lo, hi))
            assert seen == list(range(count))
