#!/bin/bash
# kernel-variant x scheduling sweep of the SE3 tracker (bench.py --no-cpu), one line per configuration
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for lib in t64b8r2k t64b8r4k t128b4r4k t128b4r2kp1 t128b4r2kp4; do
  for recs in 1 2 4; do
    if [ -n "$lib" ]; then export LSD_B200_LIB=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$lib.so; else unset LSD_B200_LIB; fi
    timeout 300 python bench.py --no-cpu --steps 5 --recs $recs > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "lib=$lib recs=$recs FAILED"; tail -2 gpurun_out/sw.err; continue; }
    python -c "
import json;d=json.load(open('gpurun_out/sw.json'))
print('lib=${lib:-base} recs=$recs value',round(d['value']),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']), d['quality']['median_translation_error_vs_gt_m'])" | tee -a gpurun_out/sweep.txt
  done
done
