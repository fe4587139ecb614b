#!/bin/bash
# ncu evidence for the round (one GPU):
#  1. launch list of one eager bench step (every kernel with its device time; compare SHARES, not absolutes)
#  2. --set full captures of the hot kernels at the C2 shapes (tools/prof_kernels.py)
R=${R:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 1 --warmup 1 --graph 0 --no-cpu-baseline --no-gpu-reference --skip-e2e > gpurun_out/${R}_bench_ncu.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${R}_launches.csv)"
for kname in ext_attn_q4_kernel ext_attn_q4d_kernel ext_attn_h2_kernel nn_field_kernel propagate_kernel layernorm_rows_kernel; do
  skip=1; [ $kname = ext_attn_q4_kernel ] && skip=2      # 3rd match = the 15-sample launch of the second loop iteration
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$kname -s $skip -c 1 -f -o gpurun_out/${R}_$kname \
      python tools/prof_kernels.py > gpurun_out/${R}_$kname.log 2>&1
  echo "$kname rc=$?"
done
ls -la gpurun_out | grep -E "${R}_.*ncu-rep|${R}_launches"
