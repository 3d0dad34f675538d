#!/bin/bash
# Runs on the GPU box under gpurun: parity tests, a short bench, and an ncu launch list.
# Everything is wrapped in `timeout` so a hung persistent kernel cannot eat the whole slot.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --pairs ${PAIRS:-200} --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench exit: $?"
cat gpurun_out/bench_small.json
tail -5 gpurun_out/bench_small.err
