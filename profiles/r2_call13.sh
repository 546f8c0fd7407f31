#!/bin/bash
# round 2, GPU call 13 (1 GPU): ncu evidence for the final build - one --set full capture of the two K = 4 streaming
# kernels (source view, DRAM traffic) and one light pass with the FP64 instruction counters that bench.py's
# roofline_fp64 numerator is cross-checked against (the JSON line of that run holds the device-side tallies)
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --batch 65536 --no-e2e --no-extras --cpu-sample 8"
bash profiles/ab.sh 1
timeout 300 $B > gpurun_out/bench_small_plain.json 2> gpurun_out/bench_small_plain.err; tail -c 300 gpurun_out/bench_small_plain.json
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:fit_kernel<[01], 4, 1>' -c 2 -o gpurun_out/fit_r2_k4 $B > gpurun_out/ncu_full_k4.log 2>&1; tail -2 gpurun_out/ncu_full_k4.log | cut -c1-200
timeout 900 ncu --metrics smsp__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k 'regex:fit_kernel<[01], 4, 1>' --csv --log-file gpurun_out/fp64_counts_r2.csv $B > gpurun_out/bench_small_under_ncu.json 2> gpurun_out/ncu_counts.err; tail -4 gpurun_out/fp64_counts_r2.csv | cut -c1-300
# memory checker over the kernels that are new in round 2 (tables with caps and bit planes, device enumeration, the
# multinomial kernels, result slabs)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_plane or cells_on or multinomial or enumeration or result_on_device" > gpurun_out/memcheck_r2.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/memcheck_r2.log | cut -c1-200
