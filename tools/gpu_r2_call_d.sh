#!/bin/bash
# 1-GPU box: dual-stream tests, N=1 dual vs fused, then the single-GPU line of every BASELINE config
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "dual or sharded_cuda or graph" > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2_pytest8.log | cut -c1-300
run() { tag=$1; shift; timeout ${TIMEOUT:-500} python bench.py "$@" > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; echo "rc=$? ($tag)"; tail -2 gpurun_out/r2_bench_$tag.err | cut -c1-300; cut -c1-${CUT:-420} gpurun_out/r2_bench_$tag.json; echo; }
run n1_c2_dual --steps 5 --warmup 3 --dual-stream 1 --no-cpu-baseline --no-gpu-reference --skip-e2e
CUT=3800 run n1_c2 --steps 5 --warmup 3 --verify
run n1_c4 --config C4 --steps 3 --warmup 3 --no-cpu-baseline --gpu-reference-steps 1
run n1_c3 --config C3 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-reference
run n1_c5s16 --config C5s16 --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-reference --skip-e2e
run n1_c5s8 --config C5s8 --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-reference --skip-e2e
TIMEOUT=700 run n1_c5s4 --config C5s4 --steps 2 --warmup 2 --no-cpu-baseline --no-gpu-reference --skip-e2e
