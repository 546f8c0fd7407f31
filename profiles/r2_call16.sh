#!/bin/bash
# round 2, GPU call 16 (8 GPUs of one box): the gathered results land in page-locked staging buffers - bench at N = 8
# again, and config 4 at full size with the ranks' scores kept on the device until ONE all-gather (gather time)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 330 $TR --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_r2_n8b.json 2> gpurun_out/bench_r2_n8b.err; tail -c 600 gpurun_out/bench_r2_n8b.json; tail -3 gpurun_out/bench_r2_n8b.err
timeout 270 $TR --master-port 29532 tests/run_config.py --config 4 --sample 1500 --out gpurun_out/config4_full_n8_r2b.json > gpurun_out/config4_n8b.log 2>&1; tail -c 900 gpurun_out/config4_full_n8_r2b.json; tail -3 gpurun_out/config4_n8b.log
