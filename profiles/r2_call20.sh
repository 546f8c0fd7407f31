#!/bin/bash
# round 2, GPU call 20 (1 GPU): full GPU suite and smoke() on the build with the multinomial kernels at one CTA per SM
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu20.log; tail -5 gpurun_out/pytest_gpu20.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
