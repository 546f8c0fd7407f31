#!/bin/bash
# round 2, GPU call 11 (1 GPU): the build that is committed - full suite, smoke(), the bench line with default
# arguments, the reference arm, config 3 at full size, ncu launch list and one --set full capture of the K = 4 kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu11.log; tail -5 gpurun_out/pytest_gpu11.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 2500 gpurun_out/bench_r2_final.json; tail -3 gpurun_out/bench_r2_final.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -c 700 gpurun_out/bench_r2_reference.json
timeout 600 python tests/run_config.py --config 3 --sample 1500 --out gpurun_out/config3_full_r2g.json > gpurun_out/config3g.log 2>&1; tail -c 400 gpurun_out/config3_full_r2g.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --batch 262144 --no-e2e --no-extras --cpu-sample 8 > gpurun_out/ncu_launches.log 2>&1; tail -1 gpurun_out/ncu_launches.log | cut -c1-300
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:fit_kernel -c 4 -o gpurun_out/fit_r2_final python bench.py --steps 1 --warmup 0 --batch 65536 --no-e2e --no-extras --cpu-sample 8 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log | cut -c1-300
