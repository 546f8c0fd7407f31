#!/bin/bash
# round 2, GPU call 15 (1 GPU): the two ncu passes of call 13 again with a kernel filter that matches the demangled
# names (fit_kernel<(int)0, (int)4, (bool)1>)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --batch 65536 --no-e2e --no-extras --cpu-sample 8"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fit_kernel<\(int\)[01], \(int\)4' -c 2 -o gpurun_out/fit_r2_k4 $B > gpurun_out/ncu_full_k4.log 2>&1; tail -2 gpurun_out/ncu_full_k4.log | cut -c1-200
timeout 900 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:fit_kernel<\(int\)[01], \(int\)4' --csv --log-file gpurun_out/fp64_counts_r2.csv $B > gpurun_out/bench_small_under_ncu.json 2> gpurun_out/ncu_counts.err; tail -4 gpurun_out/fp64_counts_r2.csv | cut -c1-300
