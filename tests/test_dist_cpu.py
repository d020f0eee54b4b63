"""CPU, world_size 2, gloo: the host-side plumbing of the one-process-per-GPU
layout (shard bounds, the data-statistics all-reduce, NCCL-id broadcast)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_pad():
    from bnmtf_ard_b200._dist import shard_bounds
    for n in (1, 5, 7, 8, 100, 20000, 20001):
        for world in (1, 2, 3, 4, 8):
            covered, size0 = [], None
            for r in range(world):
                a, b, size = shard_bounds(n, r, world)
                size0 = size if size0 is None else size0
                assert size == size0 == -(-n // world) and 0 <= a <= b <= n and b - a <= size
                assert a == min(n, r * size)
                covered += list(range(a, b))
            assert covered == list(range(n))
            assert size0 * world >= n


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    from bnmtf_ard_b200 import _dist
    out = {}
    out["info"] = _dist.info()
    out["sum"] = _dist.allreduce_sum(np.array([rank + 1.0, 10.0 * (rank + 1)])).tolist()
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    out["bytes_ok"] = _dist.broadcast_bytes(payload, src=0) == bytes(range(128))
    # the shards tile the matrix and the per-shard data statistics add up
    rng = np.random.default_rng(0)
    R = rng.normal(size=(11, 7)); M = (rng.random((11, 7)) < 0.6)
    r0, r1, _ = _dist.shard_bounds(11, rank, world)
    local = np.array([M[r0:r1].sum(), (R * M)[r0:r1].sum()])
    tot = _dist.allreduce_sum(local)
    out["stats_ok"] = bool(np.allclose(tot, [M.sum(), (R * M).sum()]))
    # seeds: every rank has its own numpy stream; next_seed() hands out rank 0's (ranks drew their own)
    from bnmtf_ard_b200.distributions._device import next_seed
    np.random.seed(100 + rank)
    out["seed"] = next_seed()
    np.random.seed(100)
    out["seed_rank0_stream"] = (lambda lo, hi: (hi << 31) | lo)(int(np.random.randint(0, 2 ** 31 - 1)), int(np.random.randint(0, 2 ** 31 - 1)))
    out["bcast"] = _dist.broadcast_int(12345678901234 + rank)
    out["same_yes"] = _dist.same_on_all_ranks([1.5, -2.0])
    out["same_no"] = _dist.same_on_all_ranks([1.5, float(rank)])
    _dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_gloo_world2_plumbing():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in range(2):
        assert res[r]["info"] == (r, 2)
        assert res[r]["sum"] == [3.0, 30.0]
        assert res[r]["bytes_ok"] and res[r]["stats_ok"]
        assert res[r]["seed"] == res[0]["seed"] == res[r]["seed_rank0_stream"]
        assert res[r]["bcast"] == 12345678901234
        assert res[r]["same_yes"] is True and res[r]["same_no"] is False


def test_single_process_defaults():
    from bnmtf_ard_b200 import _dist
    assert _dist.info() == (0, 1)
    assert _dist.allreduce_sum([1.0, 2.0]).tolist() == [1.0, 2.0]
    assert _dist.broadcast_bytes(b"abc") == b"abc"
    assert _dist.broadcast_int(7) == 7 and _dist.same_on_all_ranks([1.0])
