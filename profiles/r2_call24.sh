#!/bin/bash
# round 2, GPU call 24 (1 GPU): small-table kernel at 16 warps per SM / 2 048 counters ("default") vs 20 and 24 warps per
# SM with 512 counters ("s9x5", "s9x6"): bench (262 144-fit steps) and config 3 at full size per build
mkdir -p gpurun_out
bash profiles/ab.sh 1
for f in abn_b200/csrc/libabn_b200.so variants/libabn_b200_*.so; do
  tag=$(basename $f .so); tag=${tag#libabn_b200_}; [ "$tag" = libabn_b200 ] && tag=default
  ABN_B200_LIB=$PWD/$f timeout 300 python tests/run_config.py --config 3 --sample 300 --out gpurun_out/config3_$tag.json > gpurun_out/config3_$tag.log 2>&1
  python - "$tag" gpurun_out/config3_$tag.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "config3 score_s", round(d["score_s"],3), "fit_ms", round(d["device_ms"]["fit_ms"],1), "maxrel", d["parity"]["mlik"]["max_rel"])
PY
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/ab_*_1.json")):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k.split('<')[1][:16]:round(v['ms']) for k,v in d['roofline']['all_kernels'].items() if 'cell' in k})
PY
