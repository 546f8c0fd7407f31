#!/bin/bash
# tuning sweep over differently built copies of the library (variants/libabn_b200_<tag>.so)
mkdir -p gpurun_out
for f in variants/libabn_b200_*.so; do
  tag=$(basename $f .so); tag=${tag#libabn_b200_}
  ABN_B200_LIB=$PWD/$f python bench.py --steps 2 --warmup 1 --no-e2e --cpu-sample 16 "$@" > gpurun_out/sweep_$tag.json 2> gpurun_out/sweep_$tag.err
  python - "$tag" gpurun_out/sweep_$tag.json <<'PY'
import json,sys
tag,fn=sys.argv[1:3]
try:
    d=json.loads(open(fn).read().strip().splitlines()[-1])
    ks=d["roofline"]["all_kernels"]
    print(f"{tag:10s} value={d['value']:9.0f} fits/s  whole_frac={d['roofline']['whole_step_frac']:.3f}  " + "  ".join(f"{k.split('<')[1][:14]}={v['GBps']:.0f}" for k,v in ks.items() if 'K=4' in k), "maxrel=%.1e"%d["cpu_baseline"]["max_rel_diff_gpu_vs_oracle_on_sample"])
except Exception as e:
    print(tag,"FAILED",e)
PY
done
