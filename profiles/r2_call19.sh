#!/bin/bash
# round 2, GPU call 19 (1 GPU): multinomial kernels compiled for 2 CTAs per SM (128 registers) vs 1 (255 registers)
mkdir -p gpurun_out
timeout 400 python tests/run_mn_slice.py --sample 0 --out gpurun_out/mn_slice_blocks2.json > gpurun_out/mn_b2.log 2>&1; tail -c 400 gpurun_out/mn_b2.log
ABN_B200_LIB=$PWD/variants/libabn_b200_mn1.so timeout 400 python tests/run_mn_slice.py --sample 0 --out gpurun_out/mn_slice_blocks1.json > gpurun_out/mn_b1.log 2>&1; tail -c 400 gpurun_out/mn_b1.log
