#!/bin/bash
# round 2, GPU call 14 (8 GPUs of one box): the strong-scaling bench at N = 8 and BASELINE configs 4 and 5 at full
# size, one process per GPU under torchrun, results gathered with one NCCL all-gather
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 330 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; tail -c 1200 gpurun_out/bench_r2_n8.json; tail -3 gpurun_out/bench_r2_n8.err
timeout 270 $TR --master-port 29522 tests/run_config.py --config 4 --sample 1500 --out gpurun_out/config4_full_n8_r2.json > gpurun_out/config4_n8.log 2>&1; tail -c 900 gpurun_out/config4_full_n8_r2.json
timeout 270 $TR --master-port 29524 tests/run_config.py --config 4 --method bayes --sample 1500 --out gpurun_out/config4_bayes_full_n8_r2.json > gpurun_out/config4b_n8.log 2>&1; tail -c 700 gpurun_out/config4_bayes_full_n8_r2.json
timeout 420 $TR --master-port 29523 tests/run_config.py --config 5 --sample 600 --out gpurun_out/config5_full_n8_r2.json > gpurun_out/config5_n8.log 2>&1; tail -c 900 gpurun_out/config5_full_n8_r2.json; tail -3 gpurun_out/config5_n8.log
