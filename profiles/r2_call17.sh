#!/bin/bash
# round 2, GPU call 17 (1 GPU): the committed build once more - full GPU suite, smoke(), the bench line with default
# arguments (what the driver runs at round end)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu17.log; tail -5 gpurun_out/pytest_gpu17.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err; tail -c 1500 gpurun_out/bench_r2_final2.json; tail -3 gpurun_out/bench_r2_final2.err
