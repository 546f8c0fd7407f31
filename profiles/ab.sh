#!/bin/bash
# A/B timing of differently built copies of the library on ONE box: boxes differ by +-5 % (power cap),
# so variants are only comparable inside one call. usage: profiles/ab.sh <reps> [bench args]
reps=$1; shift
mkdir -p gpurun_out
for r in $(seq 1 $reps); do
for f in abn_b200/csrc/libabn_b200.so variants/libabn_b200_*.so; do
  tag=$(basename $f .so); tag=${tag#libabn_b200_}; [ "$tag" = libabn_b200 ] && tag=default
  ABN_B200_LIB=$PWD/$f python bench.py --steps 3 --warmup 2 --batch 262144 --no-e2e --no-extras --cpu-sample 8 "$@" > gpurun_out/ab_${tag}_$r.json 2> gpurun_out/ab_${tag}_$r.err
  python - "$tag" gpurun_out/ab_${tag}_$r.json <<'PY'
import json,sys
tag,fn=sys.argv[1:3]
try:
    d=json.loads(open(fn).read().strip().splitlines()[-1])
    ks=d["roofline"]["all_kernels"]
    print(f"{tag:10s} value={d['value']:9.0f} fits/s clk={d['clocks']['sm_mhz']:.0f} " + "  ".join(f"{k.split('<')[1][:14]}={v['ms']:.0f}ms" for k,v in ks.items() if 'K=4' in k), "maxrel=%.1e"%d["cpu_baseline"]["max_rel_diff_gpu_vs_oracle_on_sample"])
except Exception as e:
    print(tag,"FAILED",e)
PY
done
done
