#!/bin/bash
# round 2, GPU call 9 (1 GPU): contingency tables of all-binary designs from bit planes; full suite, config 3 at full
# size through device enumeration, launch list of that run
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu9.log; tail -5 gpurun_out/pytest_gpu9.log
timeout 600 python tests/run_config.py --config 3 --sample 1500 --out gpurun_out/config3_full_r2e.json > gpurun_out/config3e.log 2>&1; tail -c 1200 gpurun_out/config3_full_r2e.json; tail -3 gpurun_out/config3e.log
ABN_B200_NO_BITPLANES=1 timeout 600 python tests/run_config.py --config 3 --sample 200 --out gpurun_out/config3_full_r2e_noplanes.json > gpurun_out/config3e2.log 2>&1; tail -c 700 gpurun_out/config3_full_r2e_noplanes.json
bash profiles/ab.sh 1
