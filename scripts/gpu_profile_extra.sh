#!/bin/bash
# DepthMap / Sim3 stage bench + ncu launch list + `--set full` captures of every depth / sim3 kernel (run under gpurun).
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python scripts/bench_extra.py > gpurun_out/extra_${R}.json 2> gpurun_out/extra_${R}.err; echo "extra exit $?"; tail -5 gpurun_out/extra_${R}.err
python -c "
import json;d=json.load(open('gpurun_out/extra_${R}.json'))
for k,v in d['depthmap']['stages'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print(d['depthmap']['latency_ms'], d['depthmap'].get('cpu_port_ms_per_keyframe'))
print(d['sim3'])"
export EXTRA_REPS=2 EXTRA_NO_CPU=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k regex:^k_ -c 2000 --csv \
   --log-file gpurun_out/launches_extra_${R}.csv python scripts/bench_extra.py > gpurun_out/ncu_extra.log 2>&1
echo "ncu list exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base function \
   -k 'regex:k_depth_observe|k_depth_fill_holes|k_depth_regularize|k_prop_|k_depth_set_depth|k_depth_sums|k_idepth_pyramid|k_sim3_track' -c 60 \
   -o gpurun_out/prof_extra_${R} -f python scripts/bench_extra.py > gpurun_out/ncu_extra_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_extra_full.log
ls -la gpurun_out
