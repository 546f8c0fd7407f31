#!/bin/bash
# full ncu capture of the dominant kernel alone (Poisson K=4), see ncu_capture.sh
TAG=${1:-r1}
A="--steps 1 --warmup 1 --batch 16384 --no-e2e --cpu-sample 8"
mkdir -p gpurun_out
python bench.py $A > gpurun_out/plain3_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:fit_kernel<\(int\)1, \(int\)4' -c 1 -o gpurun_out/prof_${TAG}_pois4 \
    python bench.py $A > gpurun_out/ncu_full_pois_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_pois_${TAG}.log
# the binomial K=4 kernel alone (same command, other kernel filter):
#   ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
#       -k 'regex:fit_kernel<\(int\)0, \(int\)4' -c 1 -o gpurun_out/prof_${TAG}_bin4 python bench.py $A
