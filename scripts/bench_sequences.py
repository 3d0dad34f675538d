#!/usr/bin/env python
"""BASELINE.json configs[4]: N independent 1280x960 synthetic sequences (d2_camera-style intrinsics), one per GPU, full
lock-step tracking + mapping through the native driver (csrc/slam.cu) with HOST images.  Replicas only -- the path does
not shard inside a sequence (Multi-GPU) -- so there is no collective: one process per GPU, barrier +
synchronize on both sides of the timed loop, aggregate frames/s = total frames / max-over-ranks time.

  python scripts/bench_sequences.py                                                             # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_sequences.py
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"), os.path.join(ROOT, "scripts"), os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

W, H = 1280, 960
FRAMES = int(os.environ.get("SEQ_FRAMES", "500"))


def main():
    import torch
    import torch.distributed as dist

    import lsd_b200
    from lsd_b200 import shard, synth

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    K = synth.d2_K()
    room = synth.make_room(rank, device=f"cuda:{local}", contrast=130.0)  # sequence `rank`: its own room texture and trajectory
    traj = synth.trajectory(FRAMES, seed=rank)
    frames = []
    for i, (R, t) in enumerate(traj):
        img, depth = synth.render(room, W, H, K, R, t, noise_seed=i)
        frames.append((img.cpu().numpy(), depth.cpu().numpy() if i == 0 else None))
    ctx = lsd_b200.Context(W, H, K, device=local)
    ctx.set_live_tracking(True)  # a context that tracks one live sequence (include/lsd_b200.h)

    def run(n):
        s = lsd_b200.SlamSystem(ctx, keep_keyframes=False)
        s.gtDepthInit(frames[0][0], 0, frames[0][1])
        lost = kfs = 0
        est = [np.zeros(3)]
        for i in range(1, n):
            st = s.nextImage(frames[i][0], i)
            lost += 0 if st.tracked else 1
            kfs += st.isKeyframe
            est.append(np.array(st.camToWorld[4:7]) if st.tracked else est[-1])
        s.close()
        return lost, kfs, np.array(est)

    run(min(20, FRAMES))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lost, kfs, est = run(FRAMES)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    R0, t0_ = traj[0]
    gt = np.array([R0.T @ (t - t0_) for _, t in traj])
    ate = float(np.sqrt(np.mean(np.sum((est - gt) ** 2, axis=1))))
    dt_max, lost_max, ate_max = shard.max_over_ranks([dt, float(lost), ate]) if world > 1 else (dt, float(lost), ate)
    if rank == 0:
        print(json.dumps({"metric": "tracked+mapped frames/s, independent 1280x960 sequences, one per GPU (BASELINE configs[4])",
                          "value": world * (FRAMES - 1) / dt_max, "unit": "frames/s", "n_gpus": world, "frames_per_sequence": FRAMES,
                          "per_gpu_fps": (FRAMES - 1) / dt_max, "scaling": "weak", "collective": "none (replicas)",
                          "lost_max_over_ranks": int(lost_max), "keyframes_rank0": int(kfs), "ate_rmse_m_max_over_ranks": ate_max,
                          "h2d_bytes_per_frame": W * H}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
