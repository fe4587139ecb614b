#!/bin/bash
# One gpurun call: GPU test tier, smoke, kernel micro-benchmarks. Everything bounded by timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 2>&1 | tail -150 > gpurun_out/pytest_gpu.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 600 python tools/kbench.py --json gpurun_out/kbench.json > gpurun_out/kbench.log 2>&1
echo "kbench exit: $?" >> gpurun_out/kbench.log
tail -120 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -5; tail -30 gpurun_out/kbench.log
