"""Partitioning of the paths that shard over the GPUs of one box (DESIGN.md "Multi-GPU").

Only independent units are ever sharded -- SE3 frame pairs (BASELINE configs[1]), Sim3 constraint-search jobs
(configs[3]; upstream: the candidate loop of SlamSystem::findConstraintsForNewKeyFrames, whose per-candidate
results are combined on the host into pose-graph edges) and whole sequences (configs[4]).  One process per GPU;
there is NO data-path collective: a unit's result is a few hundred bytes (lsd_sim3_result / lsd_se3_result), so
the only communication is (1) a gather of those PODs to rank 0 and (2) the max-over-ranks of the device time.
Both run over whatever backend the process group has (nccl on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Sequence

import numpy as np


def shard_contiguous(n: int, rank: int, world: int) -> range:
    """Units [lo, hi) of rank `rank`: sizes differ by at most one, lower ranks take the remainder."""
    assert 0 <= rank < world and n >= 0
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return range(lo, lo + base + (1 if rank < rem else 0))


def shard_round_robin(n: int, rank: int, world: int) -> List[int]:
    """Static round-robin: unit i goes to rank i % world."""
    assert 0 <= rank < world and n >= 0
    return list(range(rank, n, world))


def shard_lpt(costs: Sequence[float], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment by a per-unit cost (e.g. numData of the candidate keyframe).

    Deterministic: ties are broken by unit index, the least-loaded rank with the lowest index wins.  Returns the
    unit indices of every rank, each list in ascending order."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return [sorted(v) for v in out]


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _comm_device():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def max_over_ranks(values: Sequence[float]) -> List[float]:
    """Element-wise max over all ranks (the timing rule: a multi-GPU number is the slowest rank's device time)."""
    import torch
    dist = _dist()
    if dist is None:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=_comm_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_records(local: np.ndarray, local_units: Sequence[int], n_units: int, dst: int = 0):
    """Gathers per-unit POD records to rank `dst`, placed by unit index.

    local: structured / plain array with one record per entry of `local_units`.  Returns the (n_units,) array on
    `dst` and None elsewhere.  Ranks may own different numbers of units (padded to the maximum for the gather)."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local)
    assert len(local) == len(local_units)
    if dist is None:
        out = np.zeros(n_units, dtype=local.dtype)
        out[list(local_units)] = local
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    rec = local.dtype.itemsize
    cap = (n_units + world - 1) // world if n_units else 0
    dev = _comm_device()
    cnt = torch.tensor([len(local_units)], dtype=torch.int64, device=dev)
    cap_t = cnt.clone()
    dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    cap = max(cap, int(cap_t.item()))
    idx = torch.full((cap,), -1, dtype=torch.int64)
    idx[:len(local_units)] = torch.as_tensor(list(local_units), dtype=torch.int64)
    pay = torch.zeros((cap * rec,), dtype=torch.uint8)
    if len(local):
        pay[:len(local) * rec] = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    idx, pay = idx.to(dev), pay.to(dev)
    if rank == dst:
        idxs = [torch.empty_like(idx) for _ in range(world)]
        pays = [torch.empty_like(pay) for _ in range(world)]
    else:
        idxs = pays = None
    dist.gather(idx, idxs, dst=dst)
    dist.gather(pay, pays, dst=dst)
    if rank != dst:
        return None
    out = np.zeros(n_units, dtype=local.dtype)
    seen = np.zeros(n_units, dtype=bool)
    for ii, pp in zip(idxs, pays):
        ii = ii.cpu().numpy()
        k = int((ii >= 0).sum())
        recs = pp.cpu().numpy()[:k * rec].view(local.dtype)
        assert not seen[ii[:k]].any(), "a unit was processed by two ranks"
        out[ii[:k]] = recs
        seen[ii[:k]] = True
    assert seen.all(), "a unit was processed by no rank"
    return out


def ctypes_records(structs, ctype) -> np.ndarray:
    """ctypes array / list of `ctype` structs -> (n,) array of opaque fixed-size records (np.void)."""
    n = len(structs)
    rec = C.sizeof(ctype)
    buf = np.zeros((n, rec), dtype=np.uint8)
    for i in range(n):
        buf[i] = np.frombuffer(bytes(structs[i]), dtype=np.uint8) if not isinstance(structs[i], (bytes, bytearray)) else np.frombuffer(structs[i], dtype=np.uint8)
    return buf.view(np.dtype((np.void, rec))).reshape(n)


def records_to_ctypes(records: np.ndarray, ctype):
    raw = np.ascontiguousarray(records).view(np.uint8).reshape(len(records), -1)
    return [ctype.from_buffer_copy(raw[i].tobytes()) for i in range(len(records))]


def constraint_search(track_jobs: Callable[[List[int]], Sequence], n_candidates: int, result_ctype, rank: int = 0, world: int = 1,
                      costs: Sequence[float] | None = None, dst: int = 0):
    """Sharded Sim3 constraint search (BASELINE configs[3]).

    `track_jobs(candidate_indices)` tracks this rank's candidates on this rank's GPU (both directions of
    trackFrameSim3 are the caller's business) and returns one `result_ctype` struct per index.  Candidates are
    assigned round-robin, or by LPT when per-candidate `costs` are given.  Returns (results on `dst` | None, my units)."""
    mine = shard_lpt(costs, world)[rank] if costs is not None else shard_round_robin(n_candidates, rank, world)
    res = track_jobs(mine) if mine else []
    rec = ctypes_records(res, result_ctype) if len(res) else np.zeros(0, dtype=np.dtype((np.void, C.sizeof(result_ctype))))
    allrec = gather_records(rec, mine, n_candidates, dst)
    return (records_to_ctypes(allrec, result_ctype) if allrec is not None else None), mine
