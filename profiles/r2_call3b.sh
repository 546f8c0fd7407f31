#!/bin/bash
# round 2, GPU call 3b (2 GPUs): the 2-device C-ABI test and the strong-scaling bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices or slab" 2>&1 | tail -8 > gpurun_out/pytest_gpu3b.log; tail -8 gpurun_out/pytest_gpu3b.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 1500 gpurun_out/bench_r2_n2.json; tail -5 gpurun_out/bench_r2_n2.err
