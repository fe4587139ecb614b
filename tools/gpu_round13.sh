#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r13.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 2>&1 | tail -6 >> gpurun_out/r13.log
for cfg in 0 1; do
  echo "=== nn cfg $cfg" >> gpurun_out/r13.log
  TF_NN_FIELD_CFG=$cfg timeout 300 python tools/kbench.py 2>&1 | grep -E "^nn_field_S4096|^nn_field_S1024" >> gpurun_out/r13.log
done
timeout 300 python tools/kbench.py --json gpurun_out/kbench_r13.json > gpurun_out/kbench_r13.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r13.json 2> gpurun_out/bench_r13.err
cat gpurun_out/r13.log; grep -E "ext_attn|sdpa_S4096|propagate_S4096|nn_field_S4096" gpurun_out/kbench_r13.log; cut -c1-2600 gpurun_out/bench_r13.json; tail -2 gpurun_out/bench_r13.err
