#!/bin/bash
# round 2, GPU call 10 (1 GPU): bit-plane tables with compile-time subset loops, lean multinomial pass;
# full suite, config 3 with and without planes, multinomial slice of config 5, ncu source capture of the table kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu10.log; tail -5 gpurun_out/pytest_gpu10.log
timeout 600 python tests/run_config.py --config 3 --sample 1500 --out gpurun_out/config3_full_r2f.json > gpurun_out/config3f.log 2>&1; tail -c 500 gpurun_out/config3_full_r2f.json; tail -3 gpurun_out/config3f.log
ABN_B200_NO_BITPLANES=1 timeout 600 python tests/run_config.py --config 3 --sample 200 --out gpurun_out/config3_full_r2f_noplanes.json > gpurun_out/config3f2.log 2>&1; tail -c 300 gpurun_out/config3_full_r2f_noplanes.json
timeout 900 python tests/run_mn_slice.py --out gpurun_out/mn_slice_r2.json > gpurun_out/mn_slice.log 2>&1; tail -c 1500 gpurun_out/mn_slice.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_cell_kernel -c 1 -o gpurun_out/cell_r2 python tests/run_config.py --config 3 --sample 10 --limit 600000 > gpurun_out/ncu_cell.log 2>&1; tail -2 gpurun_out/ncu_cell.log
