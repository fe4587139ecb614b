#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ext_attn_pp_kernel -s 1 -c 1 -f -o gpurun_out/prof_pp128 \
    python tools/prof_kernels.py > gpurun_out/prof_pp128.log 2>&1
tail -3 gpurun_out/prof_pp128.log; ls -la gpurun_out/*.ncu-rep
