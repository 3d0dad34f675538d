#!/bin/bash
# One GPU round trip (run under gpurun): parity tests, stage bench, per-kernel launch list, optional library-variant
# sweep, `ncu --set full` captures summarised ON THE BOX (the .ncu-rep files are deleted: gpurun_out/ is capped at 64 MiB).
#   R=r01g VARIANTS="s3b1 s3b3" SWEEP_PARTS=sim3 NCU_K='regex:k_depth_observe|k_prop_' bash scripts/gpu_round.sh
mkdir -p gpurun_out
R=${R:-r01x}
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} 2>&1 | tail -15 > gpurun_out/pytest_gpu_${R}.log; tail -4 gpurun_out/pytest_gpu_${R}.log
fi
summ() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
if "depthmap" in d:
    for k, v in d["depthmap"]["stages"].items():
        print(k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
    print(d["depthmap"]["latency_ms"], d["depthmap"].get("cpu_port_ms_per_keyframe"))
if "sim3" in d:
    s = d["sim3"]; print("sim3", {k: s[k] for k in ("ms_per_search", "kernel_ms", "candidates_per_s", "diverged", "median_scale_err") if k in s}, "frac", [v for k, v in s.items() if k.startswith("frac_")], s.get("cpu_port"))
if "vbo" in d:
    print("vbo", d["vbo"])
PY
}
if [ -z "$SKIP_EXTRA" ]; then
  timeout 1200 python scripts/bench_extra.py > gpurun_out/extra_${R}.json 2> gpurun_out/extra_${R}.err; echo "extra exit $?"; tail -5 gpurun_out/extra_${R}.err
  summ gpurun_out/extra_${R}.json
fi
for v in ${VARIANTS}; do
  EXTRA_PARTS=${SWEEP_PARTS:-sim3} EXTRA_NO_CPU=1 EXTRA_B=${SWEEP_B:-16} LSD_B200_LIB=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$v.so \
    timeout 600 python scripts/bench_extra.py > gpurun_out/extra_${R}_$v.json 2> gpurun_out/extra_${R}_$v.err || { echo "variant $v FAILED"; tail -3 gpurun_out/extra_${R}_$v.err; continue; }
  echo "== variant $v"; summ gpurun_out/extra_${R}_$v.json
done
if [ -z "$SKIP_NCU" ]; then
  export EXTRA_REPS=2 EXTRA_NO_CPU=1 EXTRA_B=16
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:^k_ -c 3000 --csv \
     --log-file gpurun_out/launches_extra_${R}.csv python scripts/bench_extra.py > gpurun_out/ncu_extra.log 2>&1
  echo "ncu list exit $?"
  python scripts/ncu_summary.py launches gpurun_out/launches_extra_${R}.csv > gpurun_out/${R}_launches_extra.txt; rm -f gpurun_out/launches_extra_${R}.csv
  cat gpurun_out/${R}_launches_extra.txt
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function \
     -k "${NCU_K:-regex:k_depth_observe|k_depth_fill_holes|k_depth_regularize|k_prop_|k_depth_set_depth|k_vbo_extract|k_sim3_track}" -c ${NCU_C:-28} \
     -o gpurun_out/prof_extra_${R} -f python scripts/bench_extra.py > gpurun_out/ncu_extra_full.log 2>&1
  echo "ncu full exit $?"; tail -2 gpurun_out/ncu_extra_full.log
  python scripts/ncu_summary.py full gpurun_out/prof_extra_${R}.ncu-rep > gpurun_out/${R}_kernels_full.txt
  for k in ${NCU_SRC_KERNELS}; do
    ncu -i gpurun_out/prof_extra_${R}.ncu-rep --page source --csv --kernel-name-base function -k $k -c 1 > gpurun_out/${R}_src_$k.csv 2>/dev/null
  done
  ls -la gpurun_out/prof_extra_${R}.ncu-rep; rm -f gpurun_out/prof_extra_${R}.ncu-rep
fi
if [ -n "$RUN_BENCH" ]; then
  timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err; echo "bench exit $?"; cat gpurun_out/bench_${R}.json; tail -3 gpurun_out/bench_${R}.err
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${R}.json 2>> gpurun_out/bench_${R}.err; cat gpurun_out/bench_ref_${R}.json
fi
du -sh gpurun_out
