#!/usr/bin/env python
"""bench.py -- tracked frames/s of the LSD-SLAM SE3 direct-alignment hot path on B200.

Workload at N=1 = BASELINE.json configs[1]: "SE3Tracker::trackFrame microbench: 640x480, 5-level
pyramid, batch of 1000 frame pairs on 1xB200".  One step = SE3Tracker::trackFrame for every pair of
the batch (synthetic textured-room renders).

  value   frames/s with keyframe references AND new-frame pyramids already resident in HBM
          (lsd_se3_track_batch; inputs 2.3 GB >> 126 MB L2, so no L2 flush is needed)
  e2e     frames/s through the reference-facing C-ABI call with HOST u8 images
          (lsd_se3_track_images_batch: H2D + pyramid build + tracking + D2H of the results)
  roofline  k_se3_track: algorithmic bytes (per LM evaluation at level l
          20 n_l + 16 min(4 n_l, N_l) + 5 n_l [l==1] + 108) / CUDA-event duration of the kernel
  cpu_baseline  the oracle port (-O3 -march=x86-64-v3) on the box's host cores, bounded sample

`--impl reference` times the CPU oracle port alone (the reference's own implementation of this path is
an un-vendored dependency and cannot be built: DESIGN.md).  N>1: pairs are sharded over ranks, no
data-path collective (weak scaling: every rank tracks its own batch).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lsd-slam-pangolin-gui_b200"))

import numpy as np  # noqa: E402

W, H = 640, 480


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1000, help="frame pairs per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0: auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--active", type=int, default=0, help="pairs in flight inside the tracker launch (0: library default)")
    ap.add_argument("--recs", type=int, default=0, help="records per work item (0: library default)")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks" line).

    NVML is polled from a thread every few ms (the timed region of this microbench is tens of ms, shorter than
    one `nvidia-smi -lms` period); if NVML cannot be loaded the nvidia-smi loop from the recipe is used instead."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, index, uuid=None):
        self.index, self.uuid = index, uuid
        self.sm, self.mx, self.reasons, self.power = [], None, set(), []
        self.proc = self.th = self.h = None
        self.stop_flag = False
        self.how = None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByUUID(self.uuid) if self.uuid else N.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.how = "nvml"
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.h = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi -lms 20"
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        N = self.N
        while not self.stop_flag:
            try:
                self.sm.append(float(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)))
                self.power.append(N.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                r = int(N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            if len(r) >= 9 and r[1].replace(".", "").isdigit():
                self.sm.append(float(r[1]))
                self.mx = float(r[2]) if r[2].replace(".", "").isdigit() else self.mx
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.th:
            self.th.join(timeout=2)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.how,
                "power_w_max": max(self.power) if self.power else None}


def make_inputs(n, seed0, device):
    """n config-2 pairs rendered with torch on `device`; returns host u8 arrays + semi-dense keyframe idepth."""
    import torch
    import torch.nn.functional as F

    from lsd_b200 import synth
    K = synth.default_K(W, H)
    kf = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    fr = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    idp = torch.empty((n, H, W), dtype=torch.float32, device=device)
    var = torch.empty((n, H, W), dtype=torch.float32, device=device)
    gt = np.zeros((n, 7))
    for i in range(n):
        pr = synth.make_pair(seed0 + i, W, H, K, device=device)
        kf[i], fr[i] = pr["kf_img"], pr["fr_img"]
        gt[i] = pr["frameToRef"]
        # semi-dense keyframe depth: where a converged DepthMap would hold hypotheses (maxGrad >= 5)
        I = pr["kf_img"].float()[None, None]
        gx = 0.5 * (I[..., 1:-1, 2:] - I[..., 1:-1, :-2])
        gy = 0.5 * (I[..., 2:, 1:-1] - I[..., :-2, 1:-1])
        mag = F.pad(torch.sqrt(gx * gx + gy * gy), (1, 1, 1, 1))
        mg = F.max_pool2d(mag, 3, 1, 1)[0, 0]
        valid = mg >= 5.0
        valid[:3] = False
        valid[-3:] = False
        valid[:, :3] = False
        valid[:, -3:] = False
        idp[i] = torch.where(valid, 1.0 / pr["kf_depth"], torch.full_like(mg, -1.0))
        var[i] = torch.where(valid, torch.full_like(mg, 0.01), torch.full_like(mg, -1.0))
    return K, kf, fr, idp, var, gt


def cpu_baseline(kf_np, fr_np, id_np, var_np, K, sample, threads, reps=1):
    """The oracle port (timing build) on `threads` host threads over `sample` pairs; returns pairs/s."""
    from oracle import pyoracle as O
    O.build()
    batch = O.RawBatch(kf_np[:sample], fr_np[:sample], id_np[:sample], var_np[:sample], K, threads, fast=True)
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (sample, 1))
    best = None
    outs = None
    for _ in range(reps):
        secs, outs = batch.track(inits, mode=0, threads=threads)
        best = secs if best is None else min(best, secs)
    poses = np.array([list(o.frameToRef) for o in outs])
    batch.free()
    return sample / best, poses


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port on all host threads, bounded sample per step."""
    if rank != 0:
        return
    import torch

    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    sample = args.cpu_sample or max(64, 4 * threads)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    K, kf, fr, idp, var, _ = make_inputs(sample, 0, dev)
    kf_np, fr_np, id_np, var_np = kf.cpu().numpy(), fr.cpu().numpy(), idp.cpu().numpy(), var.cpu().numpy()
    batch = O.RawBatch(list(kf_np), list(fr_np), list(id_np), list(var_np), K, threads, fast=True)
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (sample, 1))
    for _ in range(args.warmup):
        batch.track(inits, 0, threads)
    t = 0.0
    for _ in range(args.steps):
        secs, _ = batch.track(inits, 0, threads)
        t += secs
    val = sample * args.steps / t
    line = {"impl": "reference", "metric": "tracked frames/sec @640x480 (SE3Tracker::trackFrame)", "value": val,
            "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SE3Tracker::trackFrame microbench 640x480 5-level pyramid (BASELINE configs[1])",
                       "pairs_per_step": sample, "note": "CPU oracle port (restatement of the un-vendored lsd-slam core)"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} pairs per step, {threads} host threads, -O3 -march=x86-64-v3"},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import lsd_b200
    from lsd_b200.binding import SE3Result

    assert torch.cuda.is_available(), "bench.py needs a B200 (the product has no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.pairs
    K, kf, fr, idp, var, gt = make_inputs(n, 100000 * rank, dev)
    torch.cuda.synchronize()
    # the library launches on THIS torch stream, so torch.cuda.Event brackets see its kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    ctx = lsd_b200.Context(W, H, K, device=local_rank, stream=tstream.cuda_stream)

    if args.active:
        ctx.set_se3_active_pairs(args.active)
    if args.recs:
        ctx.set_se3_work_item_records(args.recs)
    # resident state: keyframes (with depth) -> tracking references; new frames with prebuilt pyramids
    kfs = ctx.create_frames_device(kf.data_ptr(), n)
    ctx.set_idepth_batch_device(kfs, idp.data_ptr(), var.data_ptr())
    refs = ctx.create_refs(kfs)
    frames = ctx.create_frames_device(fr.data_ptr(), n)
    inits = np.tile(np.array([0, 0, 0, 1, 0, 0, 0.0]), (n, 1))
    fr_host = fr.cpu().pin_memory()  # e2e input: pinned host u8 frames
    ip = (lsd_b200.binding.C.c_void_p * n)(*[fr_host[i].data_ptr() for i in range(n)])
    b_res = ctx.prepare_batch(refs, frames, inits)
    b_e2e = ctx.prepare_batch(refs, None, inits)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ---------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        ctx.se3_track_prepared(b_res)
    sampler = ClockSampler(local_rank, "GPU-" + str(torch.cuda.get_device_properties(dev).uuid))
    barrier()
    sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, alg_bytes, evals = [], [], []
    e0.record()
    for _ in range(args.steps):
        ctx.se3_track_prepared(b_res)
        b, e, ms = ctx.se3_last_stats()
        kern_ms.append(ms)
        alg_bytes.append(b)
        evals.append(e)
    e1.record()
    barrier()
    launches = ctx.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    res = b_res["res"]
    poses = np.array([list(res[i].frameToRef) for i in range(n)])
    n_div = sum(res[i].diverged for i in range(n))
    n_good = sum(res[i].trackingWasGood for i in range(n))
    terr = np.linalg.norm(poses[:, 4:] - gt[:, 4:], axis=1)

    # ---- e2e: host u8 frames in, poses out ----------------------------------------------------
    for _ in range(2):
        ctx.se3_track_images_prepared(b_e2e, ip, W)
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        ctx.se3_track_images_prepared(b_e2e, ip, W)
    f1.record()
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t0))  # host work included
    clocks = sampler.stop()  # sampled over both timed regions (value and e2e)
    res2 = b_e2e["res"]
    e2e_same = all(list(res2[i].frameToRef) == list(res[i].frameToRef) for i in range(n))

    # ---- max over ranks ---------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        k_ms = statistics.mean(kern_ms)
        achieved = statistics.mean(alg_bytes) / (k_ms * 1e-3) / 1e9
        line = {
            "metric": "tracked frames/sec @640x480 (SE3Tracker::trackFrame)",
            "value": world * n * args.steps / (ms_total * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "SE3Tracker::trackFrame microbench 640x480 5-level pyramid, batch of %d frame pairs per GPU "
                                   "(BASELINE configs[1])" % n,
                       "pairs_per_gpu": n, "l2": "inputs (2.3 GB of references + frame pyramids) larger than L2, no flush",
                       "parallelism": "pairs sharded over ranks, no collective" if world > 1 else "1 GPU"},
            "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": n * (W * H + 7 * 8), "d2h_bytes_per_step": n * int(lsd_b200.binding.C.sizeof(SE3Result)),
                    "same_poses_as_resident_path": bool(e2e_same)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_se3_track (persistent: all LM evaluations of the batch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src,
                         # dram__bytes_read.sum + dram__bytes_write.sum of k_se3_track from one `ncu --set full` capture of THIS workload
                         # (1000 pairs; profiles/r01z_k_se3_track_full_1000pairs.txt); other batch sizes: not captured -> null
                         "traffic": 13311593712.0 if n == 1000 else None,
                         "traffic_source": "profiles/r01z_k_se3_track_full_1000pairs.txt (12.88 GB read + 0.43 GB written per launch; below the "
                                           "15.55 GB algorithmic figure because L2 absorbs taps shared by neighbouring points: no wasted re-reads)",
                         "algorithmic_bytes_per_launch": statistics.mean(alg_bytes), "kernel_ms": k_ms,
                         "evaluations_per_launch": statistics.mean(evals)},
            "quality": {"diverged": int(n_div), "trackingWasGood": int(n_good),
                        "median_translation_error_vs_gt_m": float(np.median(terr))},
        }
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            sample = args.cpu_sample or min(n, max(64, 8 * threads))
            kf_np, fr_np = kf[:sample].cpu().numpy(), fr[:sample].cpu().numpy()
            id_np, var_np = idp[:sample].cpu().numpy(), var[:sample].cpu().numpy()
            v, cposes = cpu_baseline(list(kf_np), list(fr_np), list(id_np), list(var_np), K, sample, threads, reps=2)
            line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"first {sample} pairs of the same batch, {threads} host threads, best of 2, "
                                              f"oracle port -O3 -march=x86-64-v3 (restatement; reference core is un-vendored)",
                                    "max_pose_diff_vs_gpu": float(np.abs(cposes - poses[:sample]).max())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
