#!/bin/bash
# round 2, GPU call 2: full GPU suite (new parity tests, full-size configs), A/B of the staged / direct lean
# kernels with and without the contingency-table path, config 3 at full size, ncu of the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=12 2>&1 | tail -40 > gpurun_out/pytest_gpu2.log; tail -30 gpurun_out/pytest_gpu2.log
bash profiles/ab.sh 2
ABN_B200_NO_CELLS=1 python bench.py --steps 3 --warmup 2 --no-e2e --cpu-sample 8 > gpurun_out/ab_nocells.json 2> gpurun_out/ab_nocells.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ab_nocells.json').read().strip().splitlines()[-1])
print("default,NO_CELLS value=%.0f"%d["value"], {k:round(v["ms"]) for k,v in d["roofline"]["all_kernels"].items() if v["ms"]>50})
d=json.loads(open('gpurun_out/ab_default_1.json').read().strip().splitlines()[-1])
print("default value=%.0f"%d["value"], {k:round(v["ms"]) for k,v in d["roofline"]["all_kernels"].items() if v["ms"]>20})
PY
timeout 600 python tests/run_config.py --config 3 --out gpurun_out/config3_full_r2.json > gpurun_out/config3.log 2>&1; tail -c 1500 gpurun_out/config3_full_r2.json
A="--steps 1 --warmup 1 --batch 16384 --no-e2e --cpu-sample 8"
python bench.py $A > gpurun_out/plain_r2b.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:fit_kernel<\(int\)[01], \(int\)4|fit_cell_kernel' -c 4 -o gpurun_out/prof_r2b \
    python bench.py $A > gpurun_out/ncu_full_r2b.log 2>&1
tail -2 gpurun_out/ncu_full_r2b.log
