#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/pipes > gpurun_out/pipes.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 180 -k "ext_attn or propagate" 2>&1 | tail -40 > gpurun_out/pytest_v2.log
timeout 300 python tools/kbench.py --json gpurun_out/kbench_v2.json > gpurun_out/kbench_v2.log 2>&1
for v in "" "--no-channels-last" "--cudnn-benchmark 0"; do
  echo "== variant: $v" >> gpurun_out/bench_variants.log
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --skip-e2e $v >> gpurun_out/bench_variants.log 2>&1
done
cat gpurun_out/pipes.log; tail -15 gpurun_out/pytest_v2.log; grep -E "ext_attn|sdpa|propagate" gpurun_out/kbench_v2.log; cut -c1-400 gpurun_out/bench_variants.log
