"""N>1 host logic on CPU: partitioning of the shardable paths and the result gather / max-over-ranks timing,
run as two real processes over gloo (the GPU box uses the same code over nccl).  No compute here: the per-job
"tracker" is a deterministic stub that fills lsd_sim3_result PODs."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"))

from lsd_b200 import shard  # noqa: E402
from lsd_b200.binding import Sim3Result  # noqa: E402


def test_partitions_cover_every_unit_once():
    for n in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            for fn in (shard.shard_contiguous, shard.shard_round_robin):
                parts = [list(fn(n, r, world)) for r in range(world)]
                assert sorted(sum(parts, [])) == list(range(n))
                assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert list(shard.shard_contiguous(10, 1, 4)) == [3, 4, 5]
    assert shard.shard_round_robin(10, 1, 4) == [1, 5, 9]


def test_lpt_balances_and_is_deterministic():
    rng = np.random.default_rng(0)
    costs = rng.integers(1000, 60000, size=64).astype(float)
    parts = shard.shard_lpt(costs, 8)
    assert sorted(sum(parts, [])) == list(range(64))
    loads = [costs[p].sum() for p in parts]
    assert max(loads) <= costs.sum() / 8 + costs.max()  # LPT bound
    assert parts == shard.shard_lpt(costs, 8)
    rr = [costs[shard.shard_round_robin(64, r, 8)].sum() for r in range(8)]
    assert max(loads) <= max(rr)


def _fake_result(i):
    r = Sim3Result()
    for k in range(8):
        r.frameToRef[k] = i + 0.125 * k
    for k in range(49):
        r.lastSim3Hessian[k] = i * 100 + k
    r.lastResidual = 0.5 * i
    r.diverged = i % 3 == 0
    return r


def _worker(rank, world, port, n_units, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        calls = []

        def track(idx):
            calls.append(list(idx))
            return [_fake_result(i) for i in idx]

        costs = [float((i * 37) % 11 + 1) for i in range(n_units)]
        out = {}
        for name, cst in (("rr", None), ("lpt", costs)):
            res, mine = shard.constraint_search(track, n_units, Sim3Result, rank, world, costs=cst)
            out[name] = (None if res is None else [bytes(r) for r in res], mine)
        out["max"] = shard.max_over_ranks([1.0 + rank, 10.0 - rank])
        sub = shard.gather_records(np.arange(len(list(shard.shard_contiguous(5, rank, world))), dtype=np.float64) + 100 * rank,
                                   list(shard.shard_contiguous(5, rank, world)), 5)
        out["contig"] = None if sub is None else sub.tolist()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_units", [64, 7, 1])
def test_constraint_search_gather_world2(n_units):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    procs = [ctxmp.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [bytes(_fake_result(i)) for i in range(n_units)]
    for name in ("rr", "lpt"):
        res0, mine0 = got[0][name]
        res1, mine1 = got[1][name]
        assert res1 is None and res0 == want            # only rank 0 holds the combined result, placed by unit index
        assert sorted(mine0 + mine1) == list(range(n_units))
    assert got[0]["rr"][1] == list(range(0, n_units, 2))
    assert got[0]["max"] == got[1]["max"] == [2.0, 10.0]
    assert got[0]["contig"] == [0.0, 1.0, 2.0, 100.0, 101.0] and got[1]["contig"] is None


def test_single_process_path_needs_no_process_group():
    res, mine = shard.constraint_search(lambda idx: [_fake_result(i) for i in idx], 5, Sim3Result)
    assert mine == [0, 1, 2, 3, 4] and [bytes(r) for r in res] == [bytes(_fake_result(i)) for i in range(5)]
    assert shard.max_over_ranks([3.0]) == [3.0]
