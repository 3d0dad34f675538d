#!/bin/bash
# kernel-variant sweep of the SE3 tracker (bench.py --no-cpu), one line per configuration.
# VARIANTS="base d2 d3" (names of build/liblsd_b200_<name>.so; "base" = the shipped library)
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
timeout 600 python -m pytest tests/test_gpu_se3.py -m gpu -x -q 2>&1 | tail -5
for lib in ${VARIANTS:-base}; do
  for recs in ${RECS:-1}; do
    if [ "$lib" != "base" ]; then export LSD_B200_LIB=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$lib.so; else unset LSD_B200_LIB; fi
    timeout 300 python bench.py --no-cpu --steps 5 --recs $recs ${BENCH_ARGS} > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "lib=$lib recs=$recs FAILED"; tail -2 gpurun_out/sw.err; continue; }
    python -c "
import json;d=json.load(open('gpurun_out/sw.json'))
print('lib=${lib} recs=$recs value',round(d['value']),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']), d['quality']['median_translation_error_vs_gt_m'], d['roofline']['evaluations_per_launch'])" | tee -a gpurun_out/sweep.txt
  done
done
