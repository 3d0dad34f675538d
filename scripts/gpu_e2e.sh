#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_se3.py -m gpu -x -q 2>&1 | tail -5
for ch in 250 125 200 334 500; do
  LSD_B200_E2E_CHUNK=$ch timeout 300 python bench.py --no-cpu --steps 5 > gpurun_out/e2e.json 2> gpurun_out/e2e.err
  python -c "
import json;d=json.load(open('gpurun_out/e2e.json'))
print('chunk=$ch value',round(d['value']),'e2e',round(d['e2e']['value']), 'ms', round(1000*1000/d['e2e']['value'],2), d['e2e']['same_poses_as_resident_path'])"
  tail -2 gpurun_out/e2e.err
done
