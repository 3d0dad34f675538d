#!/bin/bash
# Round-2 GPU round trip (run under gpurun):  R=r02a [SKIP_TESTS=1] [SKIP_BENCH=1] [SKIP_NCU=1] [VARIANTS="a b"] bash scripts/gpu_r02.sh
# parity tests -> bench (both arms) -> ncu launch list of the bench command -> one `--set full` capture of k_se3_track at the
# full 1000-pair batch (summarised ON THE BOX; gpurun_out/ is capped at 64 MiB, so the .ncu-rep is deleted).
mkdir -p gpurun_out
R=${R:-r02x}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv,noheader | head -2
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS} > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?"
  grep -E "passed|failed|error|parity:|pipeline:|Error|assert" gpurun_out/${R}_pytest_gpu.log | tail -40
fi
if [ -z "$SKIP_BENCH" ]; then
  timeout 1200 python bench.py ${BENCH_ARGS} > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench exit $?"
  tail -c 6000 gpurun_out/${R}_bench.json; echo; tail -5 gpurun_out/${R}_bench.err
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err; echo "reference exit $?"
  tail -c 1500 gpurun_out/${R}_bench_reference.json; echo
fi
for v in ${VARIANTS}; do
  LSD_B200_LIB=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$v.so timeout 900 python bench.py --no-cpu ${VARIANT_ARGS:---legs parity --steps 10} \
     > gpurun_out/${R}_bench_$v.json 2> gpurun_out/${R}_bench_$v.err || { echo "variant $v FAILED"; tail -3 gpurun_out/${R}_bench_$v.err; }
  echo "== variant $v"; python - gpurun_out/${R}_bench_$v.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms")}, "parity", d.get("parity", {}).get("ok"))
    for k, v in d.get("legs", {}).items():
        print(k, json.dumps(v)[:1500])
except Exception as e:
    print("no line", e)
PY
done
if [ -z "$SKIP_NCU" ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:^k_ -c ${NCU_LIST_C:-6000} --csv \
     --log-file gpurun_out/launches_${R}.csv python bench.py --pairs 256 --steps 2 --warmup 1 --no-cpu --frames 60 --multi 1,4 ${NCU_BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
  echo "ncu list exit $?"; tail -2 gpurun_out/ncu_bench.log
  python scripts/ncu_summary.py launches gpurun_out/launches_${R}.csv > gpurun_out/${R}_launches_bench.txt; rm -f gpurun_out/launches_${R}.csv
  cat gpurun_out/${R}_launches_bench.txt
  timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:k_se3_track -s 3 -c 1 \
     -o gpurun_out/prof_se3_${R} -f python bench.py --pairs 1000 --steps 1 --warmup 3 --no-cpu --legs "" > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log
  python scripts/ncu_summary.py full gpurun_out/prof_se3_${R}.ncu-rep > gpurun_out/${R}_k_se3_track_full_1000pairs.txt
  head -50 gpurun_out/${R}_k_se3_track_full_1000pairs.txt
  python scripts/ncu_summary.py traffic gpurun_out/prof_se3_${R}.ncu-rep > gpurun_out/${R}_traffic_se3_1000pairs.json; cat gpurun_out/${R}_traffic_se3_1000pairs.json
  ncu -i gpurun_out/prof_se3_${R}.ncu-rep --page source --csv --kernel-name-base function -k k_se3_track -c 1 > gpurun_out/src_tmp.csv 2>/dev/null
  python scripts/ncu_summary.py srcsum gpurun_out/src_tmp.csv 60 > gpurun_out/${R}_src_k_se3_track.txt; rm -f gpurun_out/src_tmp.csv
  if [ -n "$NCU_EXTRA_K" ]; then
    timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$NCU_EXTRA_K" -c ${NCU_EXTRA_C:-24} \
       -o gpurun_out/prof_extra_${R} -f python bench.py --pairs 64 --steps 1 --warmup 1 --no-cpu --frames 40 --multi 1 --legs ${NCU_EXTRA_LEGS:-depth_stages,sim3_search} > gpurun_out/ncu_extra_full.log 2>&1
    echo "ncu extra exit $?"; tail -2 gpurun_out/ncu_extra_full.log
    python scripts/ncu_summary.py full gpurun_out/prof_extra_${R}.ncu-rep > gpurun_out/${R}_kernels_full.txt
    for k in ${NCU_SRC_KERNELS}; do
      ncu -i gpurun_out/prof_extra_${R}.ncu-rep --page source --csv --kernel-name-base function -k $k -c 1 > gpurun_out/src_tmp.csv 2>/dev/null
      python scripts/ncu_summary.py srcsum gpurun_out/src_tmp.csv ${NCU_SRC_TOP:-40} > gpurun_out/${R}_src_$k.txt; rm -f gpurun_out/src_tmp.csv
    done
    rm -f gpurun_out/prof_extra_${R}.ncu-rep
  fi
  rm -f gpurun_out/prof_se3_${R}.ncu-rep
fi
du -sh gpurun_out
