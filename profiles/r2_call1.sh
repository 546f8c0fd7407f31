#!/bin/bash
# round 2, GPU call 1: FP64 ceilings, full GPU test suite, parity probe, lean vs previous kernels (A/B), ncu
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt
tools/fp64_peak > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err; cat gpurun_out/fp64_peak.json
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
timeout 900 python tests/probe_parity.py > gpurun_out/probe.log 2>&1; tail -3 gpurun_out/probe.log
bash profiles/ab.sh 2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
bash profiles/ncu_capture.sh r2_lean
