#!/usr/bin/env python
"""bench.py -- the LSD-SLAM direct-alignment + depth-filtering hot path on B200, one JSON line.

Headline (`value`, `e2e`, `roofline`) = BASELINE.json configs[1]: SE3Tracker::trackFrame, 640x480, 5-level pyramid, batch of
1000 frame pairs per GPU.  One step = trackFrame for every pair of the batch (synthetic textured-room renders).

  value     frames/s with keyframe references AND new-frame pyramids resident in HBM (lsd_se3_track_batch; the inputs,
            2.3 GB, are far larger than the 126 MB L2, so no flush is needed between steps)
  e2e       frames/s through the reference-facing C-ABI call with HOST u8 images (lsd_se3_track_images_batch:
            H2D + pyramid build + tracking + D2H of the results inside the timed region)
  roofline  k_se3_track: algorithmic bytes (per LM evaluation at level l
            20 n_l + 16 min(4 n_l, N_l) + 5 n_l [l==1] + 108) / CUDA-event duration of the kernel
  parity    a sample of the SAME batch against the parity-build oracle in EXACT mode; the run FAILS when it breaks
  cpu_baseline  the oracle port (-O3 -march=x86-64-v3) on the box's host cores, the full batch

The other BASELINE configs ride in the same line under `legs` (rank 0 prints; CPU columns at N=1 only):
  track_map     configs[0]: 500-frame 640x480 lock-step track + map through lsd_slam_next_image with host images, the CPU port
                on the SAME 500 frames, ms per keyframe; and N concurrent sequences through lsd_slam_next_image_batch
  depth_stages  configs[2]: per-stage roofline of DepthMap on 64 keyframes x 10 reference frames
  sim3_search   configs[3]: one new keyframe vs 64 candidates, stages [4,3],[2],[1] x both directions, candidates sharded
                over the ranks by LPT on numData (strong scaling, no collective)
  sequences     configs[4]: one 1280x960 sequence (d2 intrinsics) per GPU, full track + map (weak scaling)

`--impl reference` times the CPU oracle port alone on the same config (the reference's own implementation of this path is an
un-vendored dependency and cannot be built: DESIGN.md).
"""
from __future__ import annotations

import argparse
import ctypes
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
IDENT7 = np.array([0, 0, 0, 1, 0, 0, 0.0])
METRIC = "tracked frames/sec @640x480 (SE3Tracker::trackFrame, 1000-pair batch)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1000, help="frame pairs per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline (0: the full batch)")
    ap.add_argument("--no-cpu", action="store_true", help="skip every CPU leg")
    ap.add_argument("--legs", default="parity,track_map,depth_stages,sim3_search,sequences",
                    help="comma list of the extra legs to run (empty: headline only)")
    ap.add_argument("--frames", type=int, default=500, help="frames of the track_map / sequences legs")
    ap.add_argument("--multi", default="1,4,16,32,64", help="concurrent sequences of the track_map leg")
    ap.add_argument("--active", type=int, default=0, help="pairs in flight inside the tracker launch (0: library default)")
    ap.add_argument("--recs", type=int, default=0, help="records per work item (0: library default)")
    return ap.parse_args()


def workload_config(n):
    """Identical in both arms (the driver compares them)."""
    return {"workload": "SE3Tracker::trackFrame microbench 640x480 5-level pyramid, batch of %d frame pairs per GPU "
                        "(BASELINE configs[1])" % n,
            "pairs_per_gpu": n, "l2": "inputs (2.3 GB of references + frame pyramids) larger than L2, no flush"}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, n_pairs):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this workload
    (profiles/traffic.json, written by scripts/ncu_summary.py from the .ncu-rep); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    for e in json.load(open(p)).get(kernel, []):
        if e.get("pairs") == n_pairs:
            return float(e["dram_bytes"]), e.get("source")
    return None, None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks" line): NVML polled from a
    thread every 2 ms (the timed region is tens of ms); nvidia-smi -lms as the fallback."""

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


def bind_to_gpu_numa_node(local_rank):
    """Best effort: run this rank, and allocate its pinned host buffers, on the NUMA node its GPU hangs off (8 ranks feeding
    307 MB per step each out of ONE socket's memory was the e2e limiter of the 8-GPU run).  Returns what was done."""
    info = {"gpu_numa_node": None, "bound": False}
    try:
        import torch
        bus = torch.cuda.get_device_properties(local_rank).pci_bus_id
        dom = torch.cuda.get_device_properties(local_rank).pci_domain_id
        dev_id = torch.cuda.get_device_properties(local_rank).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev_id)
        node = int(open(path).read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus"] = len(use)
        # memory policy: prefer the node for the pinned allocations that follow (set_mempolicy(MPOL_PREFERRED))
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))  # __NR_set_mempolicy on x86-64
        info["mempolicy_preferred"] = rc == 0
    except Exception as e:  # noqa: BLE001
        info["note"] = repr(e)[:120]
    return info


def make_inputs(n, seed0, device):
    """n config-2 pairs rendered with torch on `device`; returns u8 images + the semi-dense keyframe idepth a converged DepthMap
    would hold (maxGrad >= 5)."""
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


# --------------------------------------------------------------------------------------------------------------------
# CPU arm
# --------------------------------------------------------------------------------------------------------------------
def cpu_track_batch(kf_np, fr_np, id_np, var_np, K, threads, reps=1, warm=0):
    """The oracle port (timing build) on `threads` host threads over all given pairs; returns (pairs/s, seconds list, poses)."""
    from oracle import pyoracle as O
    O.build()
    n = len(kf_np)
    batch = O.RawBatch(kf_np, fr_np, id_np, var_np, K, threads, fast=True, trim=True)
    inits = np.tile(IDENT7, (n, 1))
    for _ in range(warm):
        batch.track(inits, 0, threads)
    secs, outs = [], None
    for _ in range(reps):
        s, outs = batch.track(inits, mode=0, threads=threads)
        secs.append(s)
    poses = np.array([list(o.frameToRef) for o in outs])
    batch.free()
    return n / min(secs), secs, poses


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port on all host threads on the SAME config (the full batch per step)."""
    if rank != 0:
        return
    import torch

    from oracle import pyoracle as O
    O.build()
    threads = os.cpu_count() or 1
    n = args.cpu_sample or args.pairs
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    K, kf, fr, idp, var, _ = make_inputs(n, 0, dev)
    kf_np, fr_np, id_np, var_np = list(kf.cpu().numpy()), list(fr.cpu().numpy()), list(idp.cpu().numpy()), list(var.cpu().numpy())
    del kf, fr, idp, var
    batch = O.RawBatch(kf_np, fr_np, id_np, var_np, K, threads, fast=True, trim=True)
    inits = np.tile(IDENT7, (n, 1))
    for _ in range(args.warmup):
        batch.track(inits, 0, threads)
    t = 0.0
    for _ in range(args.steps):
        secs, _ = batch.track(inits, 0, threads)
        t += secs
    val = n * args.steps / t
    sample = (f"the full batch: {n} pairs per step, {threads} host threads (independent pairs over threads), oracle port "
              f"-O3 -march=x86-64-v3 (restatement; the reference core is un-vendored)")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.pairs),
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# parity (checked in every run: SURVEY.md 8d last row)
# --------------------------------------------------------------------------------------------------------------------
def quat_angle(qa, qb):
    qa, qb = np.asarray(qa, np.float64), np.asarray(qb, np.float64)
    if np.dot(qa, qb) < 0:
        qb = -qb
    v = qa[3] * qb[:3] - qb[3] * qa[:3] - np.cross(qa[:3], qb[:3])
    return 2.0 * np.arcsin(min(1.0, float(np.linalg.norm(v))))


def leg_parity(ctx, refs, frames, kf_np, fr_np, id_np, var_np, K, gpu_poses, sample=32):
    """First `sample` pairs of the bench batch vs the PARITY build of the oracle in EXACT mode (fp64 accumulators = the
    order-independent value; the fp32 SCALAR / SSE4 orders differ from it by their own summation noise, which is what the old
    `max_pose_diff_vs_gpu` of the timing build was showing).  Tolerances: north_star -- counts bit-exact, residual 1e-4
    relative, final SE3 1e-5 -- the last one is only meaningful while the LM accept / reject sequences agree (a flipped
    decision, error ~ lastErr, moves the end pose by up to 1e-3 on either side)."""
    from oracle import pyoracle as O
    O.build()
    m = min(sample, len(refs))
    inits = np.tile(IDENT7, (m, 1))
    res, traces = ctx.se3_track_batch(refs[:m], frames[:m], inits, want_trace=True)
    same_as_bench = all(list(res[i].frameToRef) == list(gpu_poses[i]) for i in range(m))
    counts_exact, first_res_ok, within, flips, worst_same, worst_flip = 0, 0, 0, 0, 0.0, 0.0
    eval_counts_exact = 0
    oflips = [0, 0]
    for i in range(m):
        okf, ofr = O.Frame(2 * i, kf_np[i], K), O.Frame(2 * i + 1, fr_np[i], K)
        okf.build_pyramids()
        ofr.build_pyramids()
        okf.set_idepth(id_np[i], var_np[i])
        oref = O.Ref(okf)
        # one fused evaluation at the initial pose, level 1: bufSize / good / bad must be bit-exact
        _, _, gs = ctx.se3_eval(refs[i], frames[i], IDENT7, 1, 1.0, 0.0)
        _, _, es = O.se3_eval(oref, ofr, IDENT7, 1, 1.0, 0.0, 2)
        eval_counts_exact += int(gs[2] == es[2] and gs[3] == es[3] and gs[4] == es[4])
        eres, etrace = O.se3_track(oref, ofr, inits[i], 2)
        gtr = traces[i]
        counts_exact += int(gtr[0][4] == etrace[0][4])
        first_res_ok += int(abs(gtr[0][2] - etrace[0][2]) <= 1e-4 * abs(etrace[0][2]))
        same = len(gtr) == len(etrace) and all((a[0], a[1]) == (b[0], b[1]) for a, b in zip(gtr, etrace))
        gp, ep = np.array(res[i].frameToRef), np.array(eres.frameToRef)
        d = max(float(np.linalg.norm(gp[4:] - ep[4:])), quat_angle(gp[:4], ep[:4]))
        within += int(d <= 1e-5)
        if same:
            worst_same = max(worst_same, d)
        else:
            flips += 1
            worst_flip = max(worst_flip, d)
        for mode in (0, 1):
            _, otrace = O.se3_track(oref, ofr, inits[i], mode)
            osame = len(otrace) == len(etrace) and all((a[0], a[1]) == (b[0], b[1]) for a, b in zip(otrace, etrace))
            oflips[mode] += 0 if osame else 1
    ok = (counts_exact == m and eval_counts_exact == m and first_res_ok == m and worst_same <= 1e-5 and within >= m // 2
          and same_as_bench)
    return {"ok": bool(ok), "pairs": m, "oracle": "parity build (-O2 -ffp-contract=off), EXACT mode",
            "first_evaluation_bufSize_bit_exact": counts_exact, "level1_bufSize_good_bad_bit_exact": eval_counts_exact,
            "first_residual_within_1e-4": first_res_ok, "final_pose_within_1e-5": within,
            "lm_sequence_flips_vs_exact": {"gpu": flips, "oracle_scalar": oflips[0], "oracle_sse4": oflips[1]},
            "worst_pose_diff_same_sequence": worst_same, "worst_pose_diff_flipped": worst_flip,
            "same_poses_as_timed_batch": bool(same_as_bench),
            "tolerances": "counts bit-exact; first residual 1e-4 rel; SE3 1e-5 (scene units / rad) while the accept/reject "
                          "sequences agree"}


# --------------------------------------------------------------------------------------------------------------------
# configs[0]: lock-step track + map
# --------------------------------------------------------------------------------------------------------------------
def render_sequence(w, h, K, n_frames, seed, device, contrast=60.0):
    from lsd_b200 import synth
    room = synth.make_room(seed, device=device, contrast=contrast)
    traj = synth.trajectory(n_frames, seed=seed)
    frames = []
    for i, (R, t) in enumerate(traj):
        img, depth = synth.render(room, w, h, K, R, t, noise_seed=i)
        frames.append((img.cpu().numpy(), depth.cpu().numpy() if i == 0 else None))
    R0, t0 = traj[0]
    gt = np.array([R0.T @ (t - t0) for _, t in traj])
    return frames, gt


def run_native_sequence(lsd, ctx, frames, n):
    """lsd_slam_next_image over n frames (host images); returns (seconds, stats dict)."""
    import torch
    s = lsd.SlamSystem(ctx, keep_keyframes=False)
    s.gtDepthInit(frames[0][0], 0, frames[0][1])
    lost = kfs = 0
    t_kf = 0.0
    est = {0: np.zeros(3)}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(1, n):
        t1 = time.perf_counter()
        st = s.nextImage(frames[i][0], i)
        if not st.tracked:
            lost += 1
            continue
        est[i] = np.array(st.camToWorld[4:7])
        if st.isKeyframe:
            kfs += 1
            t_kf += time.perf_counter() - t1
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    stages = s.stage_seconds()
    s.close()
    return dt, dict(lost=lost, keyframes=kfs, kf_seconds=t_kf, est=est, stages=stages)



def set_live_records(ctx):
    """Record sizes of a context that tracks live sequences (include/lsd_b200.h: lsd_ctx_set_live_tracking).  Experiments:
    LSD_B200_BENCH_LIVE_REC = one size for every level, or four comma-separated sizes for levels 1..4."""
    env = os.environ.get("LSD_B200_BENCH_LIVE_REC")
    if not env:
        ctx.set_live_tracking(True)
    elif "," in env:
        ctx.set_se3_record_points_per_level([0] + [int(v) for v in env.split(",")])
    else:
        ctx.set_se3_record_points(int(env))


def leg_track_map(lsd, dev, local_rank, args, cpu):
    """configs[0] shape on the device: single live sequence (latency-bound: the GPU is mostly idle) and N concurrent sequences
    through the batched driver (what fills the GPU); CPU port (1 tracking + 4 mapping threads) on the same 500 frames."""
    from lsd_b200 import synth
    K = synth.default_K(W, H)
    n = args.frames
    frames, gt = render_sequence(W, H, K, n, 0, dev)
    ctx = lsd.Context(W, H, K, device=local_rank)
    set_live_records(ctx)  # live sequences: small per-level records (include/lsd_b200.h)
    run_native_sequence(lsd, ctx, frames, min(20, n))  # warm-up: pools, lazy allocations
    dt, stt = run_native_sequence(lsd, ctx, frames, n)
    ids = sorted(stt["est"])
    ate = float(np.sqrt(np.mean([np.sum((stt["est"][i] - gt[i]) ** 2) for i in ids])))
    out = {"frames": n, "fps": (n - 1) / dt, "ms_per_frame": 1e3 * dt / (n - 1), "keyframes": stt["keyframes"], "lost": stt["lost"],
           "ms_per_keyframe_switch_frame": 1e3 * stt["kf_seconds"] / max(1, stt["keyframes"]),
           "stage_ms_per_frame": {k: 1e3 * v / (n - 1) for k, v in stt["stages"].items()},
           "ate_rmse_m": ate, "path_m": float(np.linalg.norm(np.diff(gt, axis=0), axis=1).sum()), "h2d_bytes_per_frame": W * H,
           "driver": "lsd_slam_next_image (csrc/slam.cu), blocking, host images, one sequence"}
    # ms per updateKeyframe / createKeyFrame as blocking calls on the live map
    # ---- N concurrent sequences, every stage batched over the sequences
    multi = {}
    counts = [int(x) for x in args.multi.split(",") if x.strip()]
    nmax = max(counts) if counts else 0
    if nmax > 1:
        # batches of sequences run on the work-queue tracker: one record size for every level (the per-level sizes of a live
        # context would cut the coarse levels into more work items than a batch needs)
        ctx.set_live_tracking(False)
        ctx.set_se3_record_points(int(os.environ.get("LSD_B200_BENCH_MULTI_REC", "512")))
        seqs = [frames] + [render_sequence(W, H, K, n, 100 + s, dev)[0] for s in range(1, min(nmax, 8))]
        import torch
        for m in counts:
            if m < 2:
                continue
            systems = [lsd.SlamSystem(ctx, keep_keyframes=False) for _ in range(m)]
            # sequence k replays render k % 8 (8 different rooms / trajectories)
            for k, s in enumerate(systems):
                src = seqs[k % len(seqs)]
                s.gtDepthInit(src[0][0], 0, src[0][1])
            lost = kfs = 0
            warm = 3  # untimed steps: the first batch call of a new size allocates its frame / reference slabs and staging
            steps = min(n - 1, 200) - warm
            for j in range(warm):
                imgs = [seqs[k % len(seqs)][j + 1][0] for k in range(m)]
                sts = lsd.SlamSystem.nextImageBatch(systems, imgs, [j + 1] * m)
                lost += sum(0 if st.tracked else 1 for st in sts)
                kfs += sum(st.isKeyframe for st in sts)
            stage0 = dict(systems[0].stage_seconds())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for j in range(warm, warm + steps):
                # lock-step: frame j+1 of every sequence (its own render, in its own order: consecutive frames)
                imgs = [seqs[k % len(seqs)][j + 1][0] for k in range(m)]
                sts = lsd.SlamSystem.nextImageBatch(systems, imgs, [j + 1] * m)
                lost += sum(0 if st.tracked else 1 for st in sts)
                kfs += sum(st.isKeyframe for st in sts)
            torch.cuda.synchronize()
            dtm = time.perf_counter() - t0
            stage_ms = {k: 1e3 * (v - stage0[k]) * m / steps for k, v in systems[0].stage_seconds().items()}  # per step, all sequences
            for s in systems:
                s.close()
            multi[str(m)] = {"sequences": m, "frames_per_sequence": steps, "fps_total": m * steps / dtm, "fps_per_sequence": steps / dtm,
                             "ms_per_step": 1e3 * dtm / steps, "lost": lost, "keyframes": kfs, "stage_ms_per_step": stage_ms}
        out["concurrent_sequences"] = multi
        out["concurrent_sequences_note"] = ("lsd_slam_next_image_batch: N lock-step sequences on ONE context; per sequence the results "
                                            "are bit-identical to N separate systems (tests/test_gpu_pipeline.py)")
    ctx.close()
    if cpu:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from lsd_b200.pipeline import LockStepSlam
        from oracle import pyoracle as O
        from oracle_backend import OracleBackend
        O.build()
        slam = LockStepSlam(OracleBackend(W, H, K, mode=0, threads=4, fast=True))
        slam.first_frame(frames[0][0], 0, frames[0][1])
        t_kf = 0.0
        t0 = time.perf_counter()
        for i in range(1, n):
            k0 = slam.stats["keyframes"]
            t1 = time.perf_counter()
            slam.next_image(frames[i][0], i)
            if slam.stats["keyframes"] != k0:
                t_kf += time.perf_counter() - t1
        odt = time.perf_counter() - t0
        opose = {i: p[4:7] for i, p in slam.world_poses}
        common = [i for i in ids if i in opose]
        out["cpu_port"] = {"fps": (n - 1) / odt, "ms_per_frame": 1e3 * odt / (n - 1), "frames": n, "kind": "port",
                           "threads": "1 tracking + 4 mapping (upstream MAPPING_THREADS)", "keyframes": slam.stats["keyframes"],
                           "lost": slam.stats["lost"],
                           "ms_per_keyframe_switch_frame": 1e3 * t_kf / max(1, slam.stats["keyframes"]),
                           "max_translation_diff_vs_gpu_m": float(max(np.abs(opose[i] - stt["est"][i]).max() for i in common)),
                           "sample": "the same 500 frames, the same driver logic (lsd_b200/pipeline.py) on the oracle port"}
        out["speedup_vs_cpu_port"] = out["fps"] / out["cpu_port"]["fps"]
        if multi:
            best = max(multi.values(), key=lambda v: v["fps_total"])
            out["speedup_vs_cpu_port_concurrent"] = best["fps_total"] / out["cpu_port"]["fps"]
    return out


# --------------------------------------------------------------------------------------------------------------------
# configs[2]: DepthMap stages
# --------------------------------------------------------------------------------------------------------------------
def hyp_from_idepth(lsd, idepth, var, validity=20):
    h, w = idepth.shape
    m = np.zeros((h, w), lsd.binding.HYP_DTYPE)
    valid = var > 0
    m["isValid"] = valid
    m["validity_counter"] = np.where(valid, validity, 0)
    for k, src in (("idepth", idepth), ("idepth_var", var), ("idepth_smoothed", idepth), ("idepth_var_smoothed", var)):
        m[k] = np.where(valid, src, 0)
    return m


def leg_depth_stages(lsd, dev, local_rank, cpu, B=64, n_refs=10, reps=5):
    from lsd_b200 import synth
    K = synth.default_K(W, H)
    N = W * H
    ctx = lsd.Context(W, H, K, device=local_rank)
    scenes = [synth.make_depth_scene(500 + i, W, H, n_refs, K=K, device=dev) for i in range(B)]
    kfs, refs, maps0, dms = [], [], [], []
    for i, sc in enumerate(scenes):
        kf = ctx.create_frame(sc["kf_img"].cpu().numpy(), 1000, flags=lsd.BUILD_MAXGRAD0 | lsd.BUILD_GRAD0)
        idv, vv = synth.semidense_idepth(sc["kf_depth"], kf.maxGradients(0), var=0.01, noise=0.05, seed=i)
        rr = []
        for j, r in enumerate(sc["refs"]):
            f = ctx.create_frame(r["img"].cpu().numpy(), 1001 + j, flags=lsd.BUILD_MAXGRAD0)
            f.set_tracking_meta(1000, np.concatenate([r["toKf"], [1.0]]), 1.0)
            rr.append(f)
        kfs.append(kf)
        refs.append(rr)
        maps0.append(hyp_from_idepth(lsd, idv, vv))
        dms.append(ctx.create_depthmap())
    valid_frac = float(np.mean([m["isValid"].mean() for m in maps0]))

    def reset():
        for dm, kf, m0, rr in zip(dms, kfs, maps0, refs):
            dm.initializeFromMap(kf, m0)
            dm.prepare(rr)

    def timed(stage, a1=0, a2=0, frames=None):
        ts = []
        for _ in range(reps):
            reset()
            ctx.depth_stage_batch(dms, stage, a1, a2, frames)
            ts.append(ctx.last_stage_ms())
        return float(np.median(ts))

    reset()
    ctx.depth_stage_batch(dms, lsd.STAGE_OBSERVE)
    after = dms[0].read()
    s = float(((after["idepth_var"] != maps0[0]["idepth_var"]) & (maps0[0]["isValid"] > 0)).mean())
    bpp = {"observeDepth": 29 + 29 + 4 + 4 + 4 * 1 + 16 * s + 0.25,  # R_used = 1: nextStereoFrameMinID = 0 selects the oldest frame
           "fillHoles": 46.0, "regularize(false)": 30.0, "regularize(true)": 30.0,
           "propagateDepth": 29 * valid_frac + 12 + 29 * valid_frac + 5, "setDepth+pyramids": 9 + 8 * (409200 / 307200)}
    ms = {"observeDepth": timed(lsd.STAGE_OBSERVE), "fillHoles": timed(lsd.STAGE_FILL_HOLES),
          "regularize(false)": timed(lsd.STAGE_REGULARIZE, 0, 24), "regularize(true)": timed(lsd.STAGE_REGULARIZE, 1, 24),
          "setDepth+pyramids": timed(lsd.STAGE_SET_DEPTH), "propagateDepth": timed(lsd.STAGE_PROPAGATE, frames=[rr[-1] for rr in refs])}
    pk, src = peak_hbm()
    out = {"B": B, "n_refs": n_refs, "valid_fraction": valid_frac, "stereo_success_fraction": s, "peak": pk, "peak_source": src,
           "stages": {k: {"ms_batch": v, "us_per_keyframe": 1e3 * v / B, "alg_bytes_per_px": bpp[k],
                          "roofline": {"bound": "hbm", "achieved": bpp[k] * N * B / (v * 1e-3) / 1e9, "peak": pk, "unit": "GB/s",
                                       "frac": bpp[k] * N * B / (v * 1e-3) / 1e9 / pk}} for k, v in ms.items()}}
    lat_u, lat_c = [], []
    for _ in range(reps):
        reset()
        kfs[0].set_depth_updated_flag(0)
        t0 = time.perf_counter()
        dms[0].updateKeyframe(refs[0])
        lat_u.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        dms[0].createKeyFrame(refs[0][-1])
        lat_c.append(time.perf_counter() - t0)
    out["ms_per_keyframe"] = {"updateKeyframe (10 reference frames, blocking call)": 1e3 * float(np.median(lat_u)),
                              "createKeyFrame (blocking call)": 1e3 * float(np.median(lat_c))}
    if cpu:
        from oracle import pyoracle as O
        O.build()
        sc = scenes[0]
        okf = O.Frame(1000, sc["kf_img"].cpu().numpy(), K, fast=True)
        okf.build_pyramids()
        ofr = []
        for j, r in enumerate(sc["refs"]):
            f = O.Frame(1001 + j, r["img"].cpu().numpy(), K, fast=True)
            f.build_pyramids()
            f.set_track_meta(1.0, 1000, np.concatenate([r["toKf"], [1.0]]))
            ofr.append(f)
        odm = O.DepthMap(W, H, K, threads=4, fast=True)

        def ctimed(stage, a1=0, a2=0, frame=None):
            ts = []
            for _ in range(3):
                odm.init_map(okf, maps0[0])
                odm.prepare(ofr)
                ts.append(odm.stage(stage, a1, a2, frame))
            return 1e3 * float(np.median(ts))
        out["cpu_port_ms_per_keyframe"] = {"threads": 4, "kind": "port", "observeDepth": ctimed(O.STAGE_OBSERVE),
                                           "fillHoles": ctimed(O.STAGE_FILL_HOLES), "regularize(false)": ctimed(O.STAGE_REGULARIZE, 0, 24),
                                           "regularize(true)": ctimed(O.STAGE_REGULARIZE, 1, 24),
                                           "setDepth+pyramids": ctimed(O.STAGE_SET_DEPTH),
                                           "propagateDepth": ctimed(O.STAGE_PROPAGATE, frame=ofr[-1])}
    for dm in dms:
        dm.destroy()
    ctx.close()
    return out


# --------------------------------------------------------------------------------------------------------------------
# configs[3]: Sim3 constraint search, strong scaling over the ranks
# --------------------------------------------------------------------------------------------------------------------
def leg_sim3_search(lsd, dev, local_rank, rank, world, cpu, n_cand=64, steps=10):
    import torch

    from lsd_b200 import shard, synth
    from lsd_b200.pipeline import sim3_inv
    K = synth.default_K(W, H)
    ctx = lsd.Context(W, H, K, device=local_rank)
    if os.environ.get("LSD_B200_BENCH_SIM3_REC"):  # experiments: points per partial record of the Sim3 tracker (default 1024)
        ctx.set_sim3_record_points(int(os.environ["LSD_B200_BENCH_SIM3_REC"]))
    sc = synth.make_constraint_scene(4, W, H, n_cand, K=K, device=dev)
    # every rank builds the new keyframe and all candidates (setup, untimed): LPT needs every candidate's numData
    new = ctx.create_frame(sc["new"][0].cpu().numpy(), 0, flags=lsd.BUILD_MAXGRAD0)
    new_id = synth.semidense_idepth(sc["new"][1], new.maxGradients(0))
    new.set_idepth(*new_id)
    cands = ctx.create_frames([c["img"].cpu().numpy() for c in sc["cands"]], ids=list(range(1, n_cand + 1)), flags=lsd.BUILD_MAXGRAD0)
    cand_id = []
    for f, c in zip(cands, sc["cands"]):
        idv = synth.semidense_idepth(c["depth"], f.maxGradients(0))
        f.set_idepth(*idv)
        cand_id.append(idv)
    ref_new = ctx.create_refs([new])[0]
    ref_c = ctx.create_refs(cands)
    costs = [sum(r.num_data(l) for l in (1, 2, 3, 4)) for r in ref_c]
    mine = shard.shard_lpt(costs, world)[rank]
    rng = np.random.default_rng(1)
    c2n = np.array([np.concatenate([c["candToNew"], [1.0]]) for c in sc["cands"]])
    c2n[:, 4:7] += rng.normal(size=(n_cand, 3)) * 0.01  # initial estimate: GT + ~1-2 cm
    c2n[:, 7] = 1.0 + rng.uniform(-0.05, 0.05, size=n_cand)
    m = len(mine)

    # testConstraint for my candidates: tryTrackSim3 at [4,3], [2], [1]; each stage tracks candidate->new on the new keyframe's
    # reference and new->candidate on the candidate's reference, starting from the previous stage's results
    CtoF0 = c2n[mine].copy() if m else np.zeros((0, 8))
    FtoC0 = np.array([sim3_inv(p) for p in CtoF0]).reshape(-1, 8)
    s_refs = [ref_new] * m + [ref_c[i] for i in mine]
    s_frames = [cands[i] for i in mine] + [new] * m
    s_inits = np.concatenate([CtoF0, FtoC0])

    def search():
        if m == 0:
            return None, 0.0, 0.0, 0
        # ONE launch: every track runs its [4,3] -> [2] -> [1] chain on the device (lsd_sim3_track_stages_batch)
        per_stage = ctx.sim3_track_stages_batch(s_refs, s_frames, s_inits, ((4, 3), (2, 2), (1, 1)))
        byts, evals, kms = ctx.se3_last_stats()
        return per_stage[-1], byts, kms, evals

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):
        search()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        res, byts, kms, evals = search()
    barrier()
    dt = (time.perf_counter() - t0) / steps
    dt, kms_max = shard.max_over_ranks([dt, kms]) if world > 1 else (dt, kms)
    out = None
    if rank == 0:
        pk, src = peak_hbm()
        scale_err = float(np.median([abs(res[i].frameToRef[7] - 1.0) for i in range(m)])) if m else None
        out = {"candidates": n_cand, "job": "tryTrackSim3 x 3 stages ([4,3],[2],[1]) x 2 directions = 6 tracks per candidate",
               "jobs_per_s": n_cand / dt, "ms_per_search": 1e3 * dt, "n_gpus": world, "scaling": "strong",
               "sharding": "LPT on sum of numData(1..4) of the candidate, no collective (results ~600 B per candidate)",
               "candidates_on_rank0": m, "kernel_ms_rank0": kms, "kernel_ms_max_over_ranks": kms_max,
               "diverged_rank0": int(sum(r.diverged for r in res)) if m else 0, "median_scale_err_rank0": scale_err}
        if m and kms > 0:
            out["roofline"] = {"bound": "hbm", "kernel": "k_sim3_track (rank 0's share)", "achieved": byts / (kms * 1e-3) / 1e9,
                               "peak": pk, "unit": "GB/s", "frac": byts / (kms * 1e-3) / 1e9 / pk, "algorithmic_bytes": byts,
                               "evaluations": evals}
    if cpu and rank == 0 and world == 1:
        from oracle import pyoracle as O
        O.build()
        threads = os.cpu_count() or 1
        mc = min(n_cand, max(16, threads))
        onew = O.Frame(0, sc["new"][0].cpu().numpy(), K, fast=True)
        onew.build_pyramids()
        onew.set_idepth(*new_id)
        oc = []
        for i in range(mc):
            f = O.Frame(1 + i, sc["cands"][i]["img"].cpu().numpy(), K, fast=True)
            f.build_pyramids()
            f.set_idepth(*cand_id[i])
            oc.append(f)
        oref_new, oref_c = O.Ref(onew), [O.Ref(f) for f in oc]
        for r in [oref_new] + oref_c:
            for l in (1, 2, 3, 4):
                r.num(l)
        CtoF = c2n[:mc].copy()
        FtoC = np.array([sim3_inv(p) for p in CtoF])
        secs = 0.0
        for (ls, le) in ((4, 3), (2, 2), (1, 1)):
            s_, outs = O.sim3_track_batch([oref_new] * mc + oref_c, oc + [onew] * mc, np.concatenate([CtoF, FtoC]), ls, le, 0, threads)
            secs += s_
            CtoF = np.array([list(outs[i].frameToRef) for i in range(mc)])
            FtoC = np.array([list(outs[mc + i].frameToRef) for i in range(mc)])
        out["cpu_port"] = {"jobs_per_s": mc / secs, "threads": threads, "kind": "port",
                           "sample": f"the first {mc} candidates of the same search, tracks spread over the threads",
                           "max_scale_diff_vs_gpu": float(max(abs(CtoF[i][7] - res[mine.index(i)].frameToRef[7]) for i in range(mc) if i in mine))}
        out["speedup_vs_cpu_port"] = out["jobs_per_s"] / out["cpu_port"]["jobs_per_s"]
    ctx.close()
    return out


# --------------------------------------------------------------------------------------------------------------------
# configs[4]: one 1280x960 sequence per GPU
# --------------------------------------------------------------------------------------------------------------------
def leg_sequences(lsd, dev, local_rank, rank, world, args, cpu):
    import torch

    from lsd_b200 import shard, synth
    w, h = 1280, 960
    K = synth.d2_K()
    n = args.frames
    # the same wall texture seen at twice the resolution: per-pixel gradients halve, so the contrast is raised to keep about the
    # same semi-dense density
    frames, gt = render_sequence(w, h, K, n, rank, dev, contrast=130.0)
    ctx = lsd.Context(w, h, K, device=local_rank)
    set_live_records(ctx)
    run_native_sequence(lsd, ctx, frames, min(20, n))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()
    dt, stt = run_native_sequence(lsd, ctx, frames, n)
    ids = sorted(stt["est"])
    ate = float(np.sqrt(np.mean([np.sum((stt["est"][i] - gt[i]) ** 2) for i in ids])))
    dt_max, lost_max, ate_max = shard.max_over_ranks([dt, float(stt["lost"]), ate]) if world > 1 else (dt, float(stt["lost"]), ate)
    ctx.close()
    if rank != 0:
        return None
    out = {"width": w, "height": h, "intrinsics": "d2_camera.xml scaled (fx=fy=953.4, cx=623.4, cy=495.1)", "frames_per_sequence": n,
           "sequences": world, "fps_total": world * (n - 1) / dt_max, "fps_per_gpu": (n - 1) / dt_max, "scaling": "weak",
           "collective": "none (replicas)", "lost_max_over_ranks": int(lost_max), "keyframes_rank0": stt["keyframes"],
           "ms_per_keyframe_switch_frame": 1e3 * stt["kf_seconds"] / max(1, stt["keyframes"]), "ate_rmse_m_max_over_ranks": ate_max,
           "h2d_bytes_per_frame": w * h}
    if cpu and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from lsd_b200.pipeline import LockStepSlam
        from oracle import pyoracle as O
        from oracle_backend import OracleBackend
        O.build()
        mc = min(n, 100)
        slam = LockStepSlam(OracleBackend(w, h, K, mode=0, threads=4, fast=True))
        slam.first_frame(frames[0][0], 0, frames[0][1])
        t0 = time.perf_counter()
        for i in range(1, mc):
            slam.next_image(frames[i][0], i)
        odt = time.perf_counter() - t0
        out["cpu_port"] = {"fps": (mc - 1) / odt, "kind": "port", "threads": "1 tracking + 4 mapping", "keyframes": slam.stats["keyframes"],
                           "sample": f"the first {mc} frames of the same sequence"}
        out["speedup_vs_cpu_port"] = out["fps_per_gpu"] / out["cpu_port"]["fps"]
    return out


# --------------------------------------------------------------------------------------------------------------------
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
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    legs = [x for x in args.legs.split(",") if x]
    cpu = (not args.no_cpu) and rank == 0 and world == 1

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
    inits = np.tile(IDENT7, (n, 1))
    fr_host = fr.cpu().pin_memory()  # e2e input: pinned host u8 frames
    ip = (lsd_b200.binding.C.c_void_p * n)(*[fr_host[i].data_ptr() for i in range(n)])
    b_res = ctx.prepare_batch(refs, frames, inits)
    b_e2e = ctx.prepare_batch(refs, None, inits)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs ---------------------------------------------------------
    warm = max(args.warmup, 3)
    for _ in range(warm):
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
    # the same with PAGEABLE caller memory (a cv::Mat handed to nextImage, lib/App/InputThread.cpp:58-71, is not pinned)
    pageable_ms = None
    if args.steps > 0:
        fr_pg = fr.cpu().numpy()
        ipg = (lsd_b200.binding.C.c_void_p * n)(*[fr_pg[i].ctypes.data for i in range(n)])
        ctx.se3_track_images_prepared(b_e2e, ipg, W)
        barrier()
        t0 = time.perf_counter()
        reps_pg = max(2, min(args.steps, 5))
        for _ in range(reps_pg):
            ctx.se3_track_images_prepared(b_e2e, ipg, W)
        barrier()
        pageable_ms = 1e3 * (time.perf_counter() - t0) / reps_pg

    # ---- max over ranks ---------------------------------------------------------------------
    t = torch.tensor([ms_total, e2e_ms, pageable_ms or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, pageable_ms = float(t[0]), float(t[1]), float(t[2])

    line = None
    if rank == 0:
        peak, peak_src = peak_hbm()
        k_ms = statistics.mean(kern_ms)
        achieved = statistics.mean(alg_bytes) / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic("k_se3_track", n)
        h2d = n * (W * H + 7 * 8)
        line = {
            "metric": METRIC,
            "value": world * n * args.steps / (ms_total * 1e-3), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n),
            "parallelism": "pairs sharded over ranks, no collective" if world > 1 else "1 GPU",
            "e2e": {"value": world * n * args.steps / (e2e_ms * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": n * int(lsd_b200.binding.C.sizeof(SE3Result)),
                    "same_poses_as_resident_path": bool(e2e_same), "source_memory": "pinned",
                    "h2d_GBs_per_rank": h2d * args.steps / (e2e_ms * 1e-3) / 1e9,
                    "pageable_source_frames_per_s": world * n / (pageable_ms * 1e-3) if pageable_ms else None,
                    "numa": numa},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_se3_track (persistent: all LM evaluations of the batch)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": statistics.mean(alg_bytes), "kernel_ms": k_ms,
                         "evaluations_per_launch": statistics.mean(evals)},
            "quality": {"diverged": int(n_div), "trackingWasGood": int(n_good),
                        "median_translation_error_vs_gt_m": float(np.median(terr))},
            "legs": {},
        }
    kf_np = fr_np = id_np = var_np = None
    if rank == 0 and (cpu or "parity" in legs):
        m = n if cpu else 32
        kf_np, fr_np = list(kf[:m].cpu().numpy()), list(fr[:m].cpu().numpy())
        id_np, var_np = list(idp[:m].cpu().numpy()), list(var[:m].cpu().numpy())
    parity_ok = True
    if rank == 0 and "parity" in legs:
        try:
            line["parity"] = leg_parity(ctx, refs, frames, kf_np, fr_np, id_np, var_np, K, poses)
            parity_ok = line["parity"]["ok"]
        except Exception as e:  # noqa: BLE001
            line["parity"] = {"ok": False, "error": repr(e)[:300]}
            parity_ok = False
    # release the headline's resident state before the other legs allocate theirs
    for f in frames + kfs:
        f.release()
    for r in refs:
        r.release()
    ctx.close()
    del kf, fr, idp, var, fr_host
    torch.cuda.empty_cache()
    if cpu:
        threads = os.cpu_count() or 1
        sample = args.cpu_sample or n
        v, secs, cposes = cpu_track_batch(kf_np[:sample], fr_np[:sample], id_np[:sample], var_np[:sample], K, threads, reps=2)
        line["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": threads, "kind": "port",
                                "sample": f"{'the full batch: ' if sample == n else 'first '}{sample} pairs of the same batch, {threads} host "
                                          f"threads, best of 2, oracle port -O3 -march=x86-64-v3 (restatement; the reference core is "
                                          f"un-vendored)",
                                "seconds": secs}
    kf_np = fr_np = id_np = var_np = None

    def run_leg(name, fn):
        if name not in legs:
            return
        try:
            out = fn()
        except Exception as e:  # noqa: BLE001
            out = {"error": repr(e)[:400]}
        if rank == 0 and out is not None:
            line["legs"][name] = out

    if world == 1:
        run_leg("track_map", lambda: leg_track_map(lsd_b200, dev, local_rank, args, cpu))
        run_leg("depth_stages", lambda: leg_depth_stages(lsd_b200, dev, local_rank, cpu))
    run_leg("sim3_search", lambda: leg_sim3_search(lsd_b200, dev, local_rank, rank, world, cpu))
    run_leg("sequences", lambda: leg_sequences(lsd_b200, dev, local_rank, rank, world, args, cpu))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and not parity_ok:
        sys.stderr.write("bench.py: PARITY CHECK FAILED (see the `parity` key of the line above)\n")
        sys.exit(1)


if __name__ == "__main__":
    main()
