#!/bin/bash
# round 2, GPU call 5 (1 GPU): capped contingency tables (all discrete designs), 256-entry exp table default, TMA A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=5 2>&1 | tail -30 > gpurun_out/pytest_gpu6.log; tail -24 gpurun_out/pytest_gpu5.log
bash profiles/ab.sh 2
ABN_B200_NO_CELL_CAPS=1 python bench.py --steps 3 --warmup 2 --batch 262144 --no-e2e --no-extras --cpu-sample 8 > gpurun_out/ab_nocaps.json 2> gpurun_out/ab_nocaps.err
python - <<'PY'
import json
for tag in ("nocaps", "default_1"):
    d=json.loads(open(f'gpurun_out/ab_{tag}.json').read().strip().splitlines()[-1])
    print(tag, "value=%.0f"%d["value"], {k:(round(v["ms"]), v["fits"]) for k,v in d["roofline"]["all_kernels"].items() if v["ms"]>20})
PY
timeout 600 python tests/run_config.py --config 3 --out gpurun_out/config3_full_r2d.json > gpurun_out/config3d.log 2>&1; tail -c 700 gpurun_out/config3_full_r2c.json
