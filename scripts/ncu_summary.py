#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/ (run here, no GPU).

  python scripts/ncu_summary.py launches gpurun_out/launches_r01.csv  > profiles/r01_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof_se3_r01.ncu-rep  > profiles/r01_k_se3_track_full.txt
"""
import csv
import collections
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.reader(lines)
    hdr = None
    for r in rd:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            unit = d.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((d["Kernel Name"].split("(")[0], v * scale))
    agg = collections.OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot:.1f} us total device time (ncu: cold cache, serialised -- compare SHARES)")
    print(f"{'kernel':40s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:8d} {us:12.1f} {us / n:10.1f} {100 * us / tot:6.1f}%")



def srcsum(path, top=40):
    """Compact view of an `ncu --page source --csv` dump (SASS view): the instructions that collect the stall samples."""
    with open(path) as f:
        rows = list(csv.reader(f))
    name = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    body = []
    for r in rows[2:]:  # the dump may hold further views (CUDA-C lines) behind the SASS view: stop at the next header
        if len(r) != len(hdr) or r[0] == "Address" or r[0] == "Kernel Name":
            break
        body.append(r)
    tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
    inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in body)
    print(f"## {name}: {len(body)} SASS instructions, {inst} warp instructions executed, {tot} stall samples")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]] or 0) for r in body) for h in stalls}
    print("stall totals:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
    print(f"{'line':>5s} {'samples':>8s} {'%':>6s} {'executed':>10s} {'thr/inst':>8s}  top stall    instruction")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        smp = int(r[ix["# Samples"]] or 0)
        ts = max(stalls, key=lambda h: int(r[ix[h]] or 0))
        print(f"{i:5d} {smp:8d} {100.0 * smp / max(tot, 1):6.2f} {r[ix['Instructions Executed']]:>10s} {r[ix['Avg. Threads Executed']]:>8s}  {ts[6:]:12s} {r[ix['Source']].strip()[:90]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = collections.Counter()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d['Kernel Name'].split('(')[0]
        seen[name] += 1
        if seen[name] > 1:  # one block per kernel: repeated launches of the same shape add nothing
            continue
        u = dict(zip(hdr, units))
        print(f"## {d['Kernel Name'].split('(')[0]}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in KEYS:
            if k in d:
                print(f"{k:75s} {d[k]:>18s} {u[k]}")
        print("# stall reasons (warps per issue-active cycle)")
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
                try:
                    if float(d[h]) >= 0.05:
                        print(f"{h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {float(d[h]):8.3f}")
                except ValueError:
                    pass


def traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in the report, as JSON (profiles/traffic.json
    entries: bench.py's roofline.traffic reads them)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    res = collections.OrderedDict()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0]
        try:
            b = sum(float(d[k].replace(",", "")) * scale[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
            t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[u["gpu__time_duration.sum"]]
        except (KeyError, ValueError):
            continue
        res.setdefault(name, []).append({"dram_bytes": b, "ncu_ms": t, "grid": d.get("Grid Size")})
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "srcsum":
        srcsum(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
        sys.exit(0)
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
