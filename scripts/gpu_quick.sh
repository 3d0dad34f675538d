#!/bin/bash
# quick loop: SE3 parity tests + the headline bench (no CPU leg)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_se3.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --no-cpu --steps ${STEPS:-10} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms/step',round(d['ms_per_step'],3),'frac',round(d['roofline']['frac'],4),'kernel_ms',round(d['roofline']['kernel_ms'],3),'evals',d['roofline']['evaluations_per_launch'], d['quality'], d['e2e']['same_poses_as_resident_path'], d['clocks'])"
tail -3 gpurun_out/bench_quick.err
