#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 2>&1 | tail -40 > gpurun_out/pytest_r4.log
timeout 300 python tools/kbench.py --json gpurun_out/kbench_r4.json > gpurun_out/kbench_r4.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r4.json 2> gpurun_out/bench_r4.err
tail -12 gpurun_out/pytest_r4.log; grep -E "ext_attn|sdpa|propagate|nn_field|unit" gpurun_out/kbench_r4.log; cat gpurun_out/bench_r4.json | cut -c1-2500; tail -3 gpurun_out/bench_r4.err
