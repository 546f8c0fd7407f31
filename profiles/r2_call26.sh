#!/bin/bash
# round 2, GPU call 26 (1 GPU): config 3 at full size, same box, alternating: the build with the separate bit-plane kernel
# instantiation ("default") vs the commit before it ("prev")
mkdir -p gpurun_out
for rep in 1 2; do
for f in abn_b200/csrc/libabn_b200.so variants/libabn_b200_prev.so; do
  tag=$(basename $f .so); tag=${tag#libabn_b200_}; [ "$tag" = libabn_b200 ] && tag=default
  ABN_B200_LIB=$PWD/$f timeout 300 python tests/run_config.py --config 3 --sample 100 --out gpurun_out/config3_${tag}_$rep.json > gpurun_out/config3_${tag}_$rep.log 2>&1
  python - "$tag" gpurun_out/config3_${tag}_$rep.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], "config3 score_s", round(d["score_s"],3), "fit_ms", round(d["device_ms"]["fit_ms"],1), [k["name"] for k in d["kernels"]])
PY
done
done
