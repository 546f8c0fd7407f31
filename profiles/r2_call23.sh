#!/bin/bash
# round 2, GPU call 23 (4 GPUs of one box): the strong-scaling bench at N = 4 with the final build
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 3 --warmup 2 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; tail -c 700 gpurun_out/bench_r2_n4.json; tail -3 gpurun_out/bench_r2_n4.err
