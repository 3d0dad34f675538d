#!/bin/bash
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_${R}.log; tail -3 gpurun_out/pytest_gpu_${R}.log
timeout 900 python scripts/bench_extra.py > gpurun_out/extra_${R}.json 2> gpurun_out/extra_${R}.err; echo "extra exit $?"; tail -5 gpurun_out/extra_${R}.err
python -c "
import json;d=json.load(open('gpurun_out/extra_${R}.json'))
for k,v in d['depthmap']['stages'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print(d['depthmap']['latency_ms'], d['depthmap'].get('cpu_port_ms_per_keyframe'))
print(d['sim3'])"
