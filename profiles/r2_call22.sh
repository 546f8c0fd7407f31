#!/bin/bash
# round 2, GPU call 22 (1 GPU): the committed build (wide tables at 1 024 counters / 16 warps per SM) - full GPU suite,
# smoke(), the bench line with default arguments, config 3 at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu22.log; tail -5 gpurun_out/pytest_gpu22.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r2_final3.json 2> gpurun_out/bench_r2_final3.err; tail -c 1200 gpurun_out/bench_r2_final3.json; tail -3 gpurun_out/bench_r2_final3.err
timeout 600 python tests/run_config.py --config 3 --sample 1500 --out gpurun_out/config3_full_r2h.json > gpurun_out/config3h.log 2>&1; tail -c 300 gpurun_out/config3_full_r2h.json
