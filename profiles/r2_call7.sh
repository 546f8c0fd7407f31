#!/bin/bash
# round 2, GPU call 7 (1 GPU): one-step exp reduction + one-step reciprocal in the lean passes (A/B against the
# two-step / cubic forms), full suite, the bench line with default arguments
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu7.log; tail -4 gpurun_out/pytest_gpu7.log
bash profiles/ab.sh 2
timeout 1200 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r2_n1b.json 2> gpurun_out/bench_r2_n1b.err; tail -c 1800 gpurun_out/bench_r2_n1b.json; tail -3 gpurun_out/bench_r2_n1b.err
