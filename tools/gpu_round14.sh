#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r14.log
for cfg in 0 1; do
  echo "=== nn cfg $cfg" >> gpurun_out/r14.log
  TF_NN_FIELD_CFG=$cfg timeout 300 python tools/kbench.py 2>&1 | grep -E "^nn_field_S4096|^nn_field_S1024|^cublas_argmax_S4096" >> gpurun_out/r14.log
done
for fpp in 8 40; do
  echo "=== bench frames-per-pass $fpp" >> gpurun_out/r14.log
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --skip-e2e --frames-per-pass $fpp 2>/dev/null | cut -c1-160 >> gpurun_out/r14.log
done
cat gpurun_out/r14.log
