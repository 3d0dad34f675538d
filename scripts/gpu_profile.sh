#!/bin/bash
# Full-size bench + ncu launch list + one `--set full` capture of the dominant kernel (run under gpurun).
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
echo "bench exit $?"; cat gpurun_out/bench_${R}.json; tail -3 gpurun_out/bench_${R}.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${R}.json 2>> gpurun_out/bench_${R}.err
cat gpurun_out/bench_ref_${R}.json
# launch list of the same command at a reduced step count (per-launch device times, cold cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:^k_ -c 400 --csv \
   --log-file gpurun_out/launches_${R}.csv python bench.py --pairs ${NCU_PAIRS:-256} --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_se3_track -s 1 -c 1 \
   -o gpurun_out/prof_se3_${R} -f python bench.py --pairs ${NCU_PAIRS:-256} --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
