#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r11.log
for mode in pp128 pp64; do
  echo "=== attn mode $mode" >> gpurun_out/r11.log
  TF_EXT_ATTN_MODE=$mode timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -p no:cacheprovider --timeout 180 -k "ext_attn" 2>&1 | tail -3 >> gpurun_out/r11.log
  TF_EXT_ATTN_MODE=$mode timeout 300 python tools/kbench.py 2>&1 | grep -E "ext_attn_S4096|sdpa_S4096" >> gpurun_out/r11.log
done
TF_BUILD_TRACE=1 python -m tokenflow_b200._build --force > /dev/null 2>&1
for mode in pp128 pp64; do
echo "== trace $mode" >> gpurun_out/trace3.log
TF_EXT_ATTN_MODE=$mode timeout 120 python tools/trace_attn.py 2>&1 | tail -40 >> gpurun_out/trace3.log
done
cat gpurun_out/r11.log; grep -A9 "== trace pp128" gpurun_out/trace3.log; grep -A4 "MMA events" gpurun_out/trace3.log
