#!/bin/bash
# round 2, GPU call 25 (1 GPU): all-binary designs in their own kernel instantiation (bit planes, 24 warps per SM) - full
# GPU suite, smoke(), config 3 at full size, the bench line with default arguments
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu25.log; tail -5 gpurun_out/pytest_gpu25.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python tests/run_config.py --config 3 --sample 1500 --out gpurun_out/config3_full_r2i.json > gpurun_out/config3i.log 2>&1; tail -c 500 gpurun_out/config3_full_r2i.json; tail -2 gpurun_out/config3i.log
timeout 1500 python bench.py > gpurun_out/bench_r2_final4.json 2> gpurun_out/bench_r2_final4.err; tail -c 1200 gpurun_out/bench_r2_final4.json; tail -3 gpurun_out/bench_r2_final4.err
