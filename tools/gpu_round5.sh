#!/bin/bash
mkdir -p gpurun_out
for mode in pp128 pp64; do
  echo "=== mode $mode" >> gpurun_out/r5.log
  TF_EXT_ATTN_MODE=$mode timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 180 -k "ext_attn" 2>&1 | tail -8 >> gpurun_out/r5.log
  TF_EXT_ATTN_MODE=$mode timeout 300 python tools/kbench.py 2>&1 | grep -E "ext_attn|sdpa" >> gpurun_out/r5.log
done
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 2>&1 | tail -8 >> gpurun_out/r5.log
cat gpurun_out/r5.log
