#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?" >> gpurun_out/bench_n1.err
# every launch of one warm step (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --skip-e2e > gpurun_out/bench_ncu.log 2>&1
# full captures of the three kernels at C2 top-level shapes
for kname in ext_attn_kernel nn_field_kernel propagate_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kname -s 1 -c 2 -f -o gpurun_out/prof_$kname \
      python tools/prof_kernels.py > gpurun_out/prof_$kname.log 2>&1
done
tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-3000; ls -la gpurun_out
