#!/bin/bash
# round 2, GPU call 27 (1 GPU): ncu --set full of fit_kernel<poisson, K = 4> of the final build (the capture of call 15
# returned no counters for it) and of the wide-table kernel; summarised on the box (the reports exceed the 64 MiB that
# come back from a call)
mkdir -p gpurun_out /tmp/ncu
B="python bench.py --steps 1 --warmup 0 --batch 65536 --no-e2e --no-extras --cpu-sample 8"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:fit_kernel<\(int\)1, \(int\)4' -c 1 -o /tmp/ncu/fit_r2_pois4 $B > gpurun_out/ncu_pois4.log 2>&1; tail -1 gpurun_out/ncu_pois4.log | cut -c1-200
python tools/ncu_summary.py /tmp/ncu/fit_r2_pois4.ncu-rep > gpurun_out/ncu_r2_final_pois4_summary.txt
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:fit_cell_kernel<\(int\)0, \(int\)10' -c 1 -o /tmp/ncu/cell_r2_large $B > gpurun_out/ncu_cell_large.log 2>&1; tail -1 gpurun_out/ncu_cell_large.log | cut -c1-200
python tools/ncu_summary.py /tmp/ncu/cell_r2_large.ncu-rep > gpurun_out/ncu_r2_final_cell_large_summary.txt
grep -E "fp64_cycles_active|issue_active|gpu__time_duration" gpurun_out/ncu_r2_final_*_summary.txt | cut -c1-200
