#!/bin/bash
# A/B of kernel variants on one box (run under gpurun):
#   R=r02e VARIANTS="main notma ..." [LEGS=depth_stages,sim3_search] [TESTS="tests/test_gpu_depth.py"] bash scripts/gpu_ab.sh
# Every variant is build/liblsd_b200_<name>.so (make variant ...; "main" = the shipped library).  Prints one line per variant and stage.
mkdir -p gpurun_out
R=${R:-r02x}
LEGS=${LEGS:-depth_stages}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1
if [ -n "$TESTS" ]; then
  timeout 1500 python -m pytest $TESTS -m gpu -q -x > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?"
  grep -E "passed|failed|error|parity:|pipeline:|Error|assert" gpurun_out/${R}_pytest_gpu.log | tail -15
fi
for vv in ${VARIANTS}; do  # name or name@ENV=VALUE (an environment switch on top of the library variant)
  v=${vv%%@*}; envset=""; [ "$vv" != "$v" ] && envset=${vv#*@}
  lib=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$v.so
  [ "$v" = main ] && lib=$PWD/lsd-slam-pangolin-gui_b200/liblsd_b200.so
  [ -n "$envset" ] && export "$envset"
  if [ -n "$VTESTS" ]; then
    LSD_B200_LIB=$lib timeout 900 python -m pytest $VTESTS -m gpu -q -x > gpurun_out/${R}_pytest_$v.log 2>&1; echo "== $v pytest exit $? $(tail -1 gpurun_out/${R}_pytest_$v.log)"
  fi
  LSD_B200_LIB=$lib timeout 900 python bench.py --no-cpu --legs $LEGS --pairs ${PAIRS:-64} --steps 3 --warmup 3 --frames ${FRAMES:-200} --multi ${MULTI:-1} \
     > gpurun_out/${R}_ab_$v.json 2> gpurun_out/${R}_ab_$v.err || { echo "variant $v FAILED"; tail -3 gpurun_out/${R}_ab_$v.err; }
  [ -n "$envset" ] && unset "${envset%%=*}"
  python - gpurun_out/${R}_ab_$v.json $vv <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = d.get("legs", {})
    if "depth_stages" in L:
        print(sys.argv[2], "depth", {k: round(v["ms_batch"], 3) for k, v in L["depth_stages"]["stages"].items()}, L["depth_stages"]["ms_per_keyframe"])
    if "sim3_search" in L:
        s = L["sim3_search"]
        print(sys.argv[2], "sim3", {k: s[k] for k in ("ms_per_search", "kernel_ms_rank0", "diverged_rank0")}, "frac", round(s["roofline"]["frac"], 3))
    if "track_map" in L:
        t = L["track_map"]
        print(sys.argv[2], "track_map", {k: t[k] for k in ("fps", "keyframes", "lost", "stage_ms_per_frame")},
              {k: round(v["fps_total"]) for k, v in t.get("concurrent_sequences", {}).items()})
    if "sequences" in L:
        print(sys.argv[2], "sequences", {k: L["sequences"][k] for k in ("fps_total", "keyframes_rank0", "lost_max_over_ranks")})
    print(sys.argv[2], "se3", {k: d[k] for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "parity", d.get("parity", {}).get("ok"))
except Exception as e:
    print(sys.argv[2], "no line", repr(e))
PY
done
if [ -n "$NCU_K" ]; then  # one --set full capture of the named kernels with the variant NCU_V (default main)
  v=${NCU_V:-main}
  lib=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$v.so
  [ "$v" = main ] && lib=$PWD/lsd-slam-pangolin-gui_b200/liblsd_b200.so
  LSD_B200_LIB=$lib timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function -k "$NCU_K" -c ${NCU_C:-16} \
     -o gpurun_out/prof_${R} -f python bench.py --pairs 64 --steps 1 --warmup 1 --no-cpu --frames 40 --multi 1 --legs ${NCU_LEGS:-$LEGS} > gpurun_out/ncu_ab.log 2>&1
  echo "ncu exit $?"; tail -2 gpurun_out/ncu_ab.log
  python scripts/ncu_summary.py full gpurun_out/prof_${R}.ncu-rep > gpurun_out/${R}_kernels_full.txt
  for k in ${NCU_SRC_KERNELS}; do
    ncu -i gpurun_out/prof_${R}.ncu-rep --page source --csv --kernel-name-base function -k $k -c 1 > gpurun_out/src_tmp.csv 2>/dev/null
    python scripts/ncu_summary.py srcsum gpurun_out/src_tmp.csv ${NCU_SRC_TOP:-60} > gpurun_out/${R}_src_$k.txt; rm -f gpurun_out/src_tmp.csv
  done
  rm -f gpurun_out/prof_${R}.ncu-rep
  grep -E "^## |gpu__time_duration|smsp__inst_executed.sum|issue_active|warps_active|long_scoreboard|barrier" gpurun_out/${R}_kernels_full.txt | head -80
fi
du -sh gpurun_out
