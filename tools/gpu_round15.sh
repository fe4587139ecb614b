#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r15.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 180 -k "ext_attn" 2>&1 | tail -3 >> gpurun_out/r15.log
for cfg in "pp128 2" "pp128 1" "pp64 0"; do
  set -- $cfg
  echo "=== attn mode $1 handoff $2" >> gpurun_out/r15.log
  TF_EXT_ATTN_MODE=$1 TF_EXT_ATTN_HANDOFF=$2 timeout 300 python tools/kbench.py 2>&1 | grep -E "ext_attn_S4096|sdpa_S4096" >> gpurun_out/r15.log
done
cat gpurun_out/r15.log
