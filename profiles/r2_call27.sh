#!/bin/bash
# round 2, GPU call 27 (1 GPU): ncu --set full of fit_kernel<poisson, K = 4> of the final build (the capture of call 15
# returned no counters for it) and of the wide-table kernel
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --batch 65536 --no-e2e --no-extras --cpu-sample 8"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fit_kernel<\(int\)1, \(int\)4' -c 1 -o gpurun_out/fit_r2_pois4 $B > gpurun_out/ncu_pois4.log 2>&1; tail -2 gpurun_out/ncu_pois4.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:fit_cell_kernel<\(int\)0, \(int\)10' -c 1 -o gpurun_out/cell_r2_large $B > gpurun_out/ncu_cell_large.log 2>&1; tail -2 gpurun_out/ncu_cell_large.log | cut -c1-200
