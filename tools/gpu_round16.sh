#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r16.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 2>&1 | tail -5 >> gpurun_out/r16.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r16.json 2> gpurun_out/bench_r16.err
cat gpurun_out/r16.log; cut -c1-3000 gpurun_out/bench_r16.json; tail -2 gpurun_out/bench_r16.err
