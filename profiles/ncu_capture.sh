#!/bin/bash
# ncu captures used for profiles/: run under gpurun from the repo root.
#   bash profiles/ncu_capture.sh <tag>
# 1) plain run (must exit 0), 2) launch list of the same command, 3) --set full of the two dominant kernels
set -u
TAG=${1:-r1}
A="--steps 1 --warmup 1 --batch 16384 --no-e2e --cpu-sample 8"
mkdir -p gpurun_out
python bench.py $A > gpurun_out/plain_${TAG}.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py $A > gpurun_out/ncu_launches_${TAG}.log 2>&1
python bench.py $A > gpurun_out/plain2_${TAG}.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:fit_kernel<\(int\)[01], \(int\)4' -c 2 -o gpurun_out/prof_${TAG}_k4 \
    python bench.py $A > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out
