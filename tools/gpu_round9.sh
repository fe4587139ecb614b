#!/bin/bash
mkdir -p gpurun_out
TF_BUILD_TRACE=1 python -m tokenflow_b200._build --force > /dev/null 2>&1
for mode in pp128 pp64; do
  echo "=== $mode"
  TF_EXT_ATTN_MODE=$mode timeout 120 python tools/trace_attn.py 2>&1 | tail -40
done > gpurun_out/trace.log 2>&1
cat gpurun_out/trace.log
