#!/bin/bash
# round 2, GPU call 3a (1 GPU): full GPU suite, A/B line, the new bench line at N = 1, config 3 at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -30 > gpurun_out/pytest_gpu3.log; tail -22 gpurun_out/pytest_gpu3.log
bash profiles/ab.sh 1
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 3000 gpurun_out/bench_r2_n1.json; tail -5 gpurun_out/bench_r2_n1.err
timeout 600 python tests/run_config.py --config 3 --out gpurun_out/config3_full_r2b.json > gpurun_out/config3b.log 2>&1; tail -c 900 gpurun_out/config3_full_r2b.json
