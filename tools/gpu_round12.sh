#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r12.log
for cfg in "pp128 0" "pp128 1" "pp128 2" "pp64 0" "pp64 -1"; do
  set -- $cfg
  echo "=== attn mode $1 handoff $2" >> gpurun_out/r12.log
  TF_EXT_ATTN_MODE=$1 TF_EXT_ATTN_HANDOFF=$2 timeout 300 python tools/kbench.py 2>&1 | grep -E "ext_attn_S4096" >> gpurun_out/r12.log
done
cat gpurun_out/r12.log
