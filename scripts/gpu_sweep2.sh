#!/bin/bash
# library-variant sweep over the headline bench (batch of 1000 pairs) AND the live pipeline (one pair at a time)
mkdir -p gpurun_out
for lib in ${VARIANTS:-base}; do
  if [ "$lib" != "base" ]; then export LSD_B200_LIB=$PWD/lsd-slam-pangolin-gui_b200/build/liblsd_b200_$lib.so; else unset LSD_B200_LIB; fi
  timeout 300 python bench.py --no-cpu --steps 5 > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "lib=$lib bench FAILED"; tail -2 gpurun_out/sw.err; }
  EXTRA_PARTS=${SWEEP_PARTS:-pipeline} EXTRA_NO_CPU=1 EXTRA_FRAMES=300 timeout 300 python scripts/bench_extra.py > gpurun_out/sw2.json 2> gpurun_out/sw2.err || { echo "lib=$lib extra FAILED"; tail -2 gpurun_out/sw2.err; }
  python - "$lib" <<'PY'
import json, sys
lib = sys.argv[1]
try:
    d = json.load(open('gpurun_out/sw.json'))
    print('lib=%s batch: value %.0f kernel_ms %.3f frac %.3f e2e %.0f' % (lib, d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value']))
except Exception as e:
    print('lib=%s batch: n/a' % lib, e)
try:
    e = json.load(open('gpurun_out/sw2.json'))
    for k, v in e.items():
        if k.startswith('pipeline'):
            print('lib=%s %s: fps %.0f lost %d keyframes %d stage %s' % (lib, k, v['fps'], v['lost'], v['keyframes'], {a: round(b, 3) for a, b in v['stage_ms_per_frame'].items()}))
        if k == 'sim3':
            print('lib=%s sim3: kernel_ms %.3f candidates/s %.0f diverged %d scale_err %.2e' % (lib, v['kernel_ms'], v['candidates_per_s'], v['diverged'], v['median_scale_err']))
except Exception as e:
    print('lib=%s extra: n/a' % lib, e)
PY
done 2>&1 | tee gpurun_out/sweep2_${R:-r01}.txt
