#!/bin/bash
# round 2, GPU call 8 (1 GPU): lean row ops staged over 4 rows in lockstep (A/B against the row-by-row build),
# rank precheck, whole-cache config 3 through device enumeration, full suite, ncu launch list of the contingency kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu8.log; tail -4 gpurun_out/pytest_gpu8.log
bash profiles/ab.sh 2
timeout 600 python tests/run_config.py --config 3 --check 1500 > gpurun_out/config3_full_r2c.json 2> gpurun_out/config3c.log; tail -c 900 gpurun_out/config3_full_r2c.json; tail -3 gpurun_out/config3c.log
timeout 1200 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_r2_n1c.json 2> gpurun_out/bench_r2_n1c.err; tail -c 1800 gpurun_out/bench_r2_n1c.json; tail -3 gpurun_out/bench_r2_n1c.err
