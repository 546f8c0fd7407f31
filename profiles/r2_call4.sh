#!/bin/bash
# round 2, GPU call 4 (1 GPU): full suite, A/B of exp-table / TMA variants, ncu --set full with source of the two K = 4 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -20 > gpurun_out/pytest_gpu4.log; tail -14 gpurun_out/pytest_gpu4.log
bash profiles/ab.sh 2
A="--steps 1 --warmup 1 --batch 16384 --no-e2e --no-extras --cpu-sample 8"
python bench.py $A > gpurun_out/plain_r2c.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:fit_kernel<\(int\)[01], \(int\)4' -c 2 -o gpurun_out/prof_r2c \
    python bench.py $A > gpurun_out/ncu_full_r2c.log 2>&1
tail -2 gpurun_out/ncu_full_r2c.log
