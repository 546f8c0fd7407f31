#!/bin/bash
# round 2, GPU call 12 (2 GPUs): the 2-device C-ABI test, the slab gather tests, the multinomial tests with the
# one-fit-per-CTA kernel, the strong-scaling bench at N = 2 under torchrun (NCCL all-gather inside the timed region),
# and the multinomial slice of config 5 with the CTA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_devices or slab or multinomial" 2>&1 | tail -8 > gpurun_out/pytest_gpu12.log; tail -8 gpurun_out/pytest_gpu12.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 1500 gpurun_out/bench_r2_n2.json; tail -5 gpurun_out/bench_r2_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_r2_n2_reference.json 2> gpurun_out/bench_r2_n2_reference.err; tail -c 300 gpurun_out/bench_r2_n2_reference.json
timeout 600 python tests/run_mn_slice.py --out gpurun_out/mn_slice_r2_cta.json > gpurun_out/mn_slice_cta.log 2>&1; tail -c 1200 gpurun_out/mn_slice_cta.log
