#!/usr/bin/env python
"""BASELINE.json configs[3]: Sim3 constraint search -- one new keyframe against 64 candidate keyframes, the candidates
sharded round-robin over the ranks (one process per GPU, no data-path collective; results are ~600 B per candidate and
are gathered on rank 0 with torch.distributed only to show the host-side combine).  Run:

  python scripts/bench_sim3_shard.py                       # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_sim3_shard.py

A job = reciprocal trackFrameSim3 (new keyframe -> candidate and candidate -> new keyframe), levels 4 -> 1.  Every rank
holds the new keyframe (pyramids + reference, ~6 MB) and only ITS candidates.  Timing: barrier + synchronize on both
sides, max over ranks; rank 0 prints one JSON line with jobs/s.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200")]

import numpy as np  # noqa: E402

W, H = 640, 480
N_CAND = int(os.environ.get("SIM3_CANDIDATES", "64"))
STEPS = int(os.environ.get("SIM3_STEPS", "10"))


def main():
    import torch
    import torch.distributed as dist

    import lsd_b200
    from lsd_b200 import shard, synth
    from lsd_b200.binding import Sim3Result
    from lsd_b200.pipeline import sim3_inv

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    K = synth.default_K(W, H)
    ctx = lsd_b200.Context(W, H, K, device=local)
    mine = shard.shard_round_robin(N_CAND, rank, world)
    # candidate i = keyframe rendered at a pose near the new keyframe (seed 900 + i): make_pair gives (new KF view, candidate view)
    # with the same "new keyframe" camera for every i when the pair seed only moves the second camera -- here each pair has its own
    # new-KF render, which is the same amount of work per job and keeps the generator shared with bench_extra.py
    kf_imgs, cand_imgs, gts, depths = [], [], [], []
    for i in mine:
        pr = synth.make_pair(900 + i, W, H, K, device=f"cuda:{local}", max_t=0.10, max_r=np.radians(3.0))
        kf_imgs.append(pr["kf_img"].cpu().numpy()); cand_imgs.append(pr["fr_img"].cpu().numpy())
        gts.append(np.concatenate([pr["frameToRef"], [1.0]])); depths.append((pr["kf_depth"], pr["fr_depth"]))
    A = ctx.create_frames(kf_imgs, flags=lsd_b200.BUILD_MAXGRAD0)
    B = ctx.create_frames(cand_imgs, flags=lsd_b200.BUILD_MAXGRAD0)
    for a, b, (da, db) in zip(A, B, depths):
        a.set_idepth(*synth.semidense_idepth(da, a.maxGradients(0)))
        b.set_idepth(*synth.semidense_idepth(db, b.maxGradients(0)))
    refA, refB = ctx.create_refs(A), ctx.create_refs(B)
    rng = np.random.default_rng(1 + rank)
    ab = np.array(gts).reshape(-1, 8)
    ab[:, 4:7] += rng.normal(size=(len(mine), 3)) * 0.01
    ba = np.array([sim3_inv(g) for g in ab]).reshape(-1, 8)
    refs, frames, inits = refA + refB, B + A, np.concatenate([ab, ba])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def search():
        return ctx.sim3_track_batch(refs, frames, inits, 4, 1) if mine else []

    for _ in range(3):
        res = search()
    barrier()
    t0 = time.perf_counter()
    for _ in range(STEPS):
        res = search()
    barrier()
    dt = (time.perf_counter() - t0) / STEPS
    dt = shard.max_over_ranks([dt])[0] if world > 1 else dt
    # host-side combine: per-candidate results (both directions) on rank 0, as findConstraintsForNewKeyFrames would use them
    m = len(mine)
    fwd, _ = shard.constraint_search(lambda idx: [res[k] for k in range(m)], N_CAND, Sim3Result, rank, world)
    if rank == 0:
        scale_err = float(np.median([abs(r.frameToRef[7] - 1.0) for r in fwd]))
        print(json.dumps({"metric": "Sim3 constraint-search jobs/s (new keyframe vs 64 candidates, reciprocal trackFrameSim3, levels 4->1)",
                          "value": N_CAND / dt, "unit": "jobs/s", "n_gpus": world, "ms_per_search": 1e3 * dt, "candidates": N_CAND,
                          "scaling": "strong", "collective": "none on the data path (600 B per candidate gathered on the host)",
                          "diverged": int(sum(r.diverged for r in fwd)), "median_scale_err": scale_err}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
