#!/bin/bash
# round-2 infrastructure check on one B200: GPU tests, then the bench in graph / eager mode with --verify
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest3.log
timeout 900 python bench.py --steps 5 --warmup 3 --verify > gpurun_out/r2_bench_graph.json 2> gpurun_out/r2_bench_graph.err; echo "bench graph rc=$?"; tail -3 gpurun_out/r2_bench_graph.err; cut -c1-3000 gpurun_out/r2_bench_graph.json
timeout 600 python bench.py --steps 5 --warmup 3 --graph 0 --no-cpu-baseline --no-gpu-reference --skip-e2e > gpurun_out/r2_bench_eager.json 2> gpurun_out/r2_bench_eager.err; echo "bench eager rc=$?"; tail -3 gpurun_out/r2_bench_eager.err; cut -c1-1500 gpurun_out/r2_bench_eager.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-kernel-events --no-cpu-baseline --no-gpu-reference --skip-e2e > gpurun_out/r2_bench_graph_noev.json 2> gpurun_out/r2_bench_graph_noev.err; echo "bench graph noev rc=$?"; tail -3 gpurun_out/r2_bench_graph_noev.err; cut -c1-600 gpurun_out/r2_bench_graph_noev.json
