#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/bench_ncu_final.log 2>&1
for kname in ext_attn_pp_kernel nn_field_kernel propagate_kernel layernorm_unit_rows_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kname -s 1 -c 1 -f -o gpurun_out/final_$kname \
      python tools/prof_kernels.py > gpurun_out/final_$kname.log 2>&1
done
ls -la gpurun_out | grep -E "final|launches"
