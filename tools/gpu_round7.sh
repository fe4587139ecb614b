#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r7.log
for mode in pp128 pp64; do
  echo "=== mode $mode" >> gpurun_out/r7.log
  TF_EXT_ATTN_MODE=$mode timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 180 -k "ext_attn" 2>&1 | tail -4 >> gpurun_out/r7.log
  TF_EXT_ATTN_MODE=$mode timeout 300 python tools/kbench.py 2>&1 | grep -E "ext_attn_S4096|sdpa_S4096" >> gpurun_out/r7.log
done
cat gpurun_out/r7.log
