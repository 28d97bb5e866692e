#!/bin/bash
# Interleaved A/B of env-toggled kernels on one box: usage tools/gpu_ab.sh TAG "ENV=1 ENV2=1" [rounds] [bench args]
set -u
TAG=$1; ENVS=$2; ROUNDS=${3:-3}; shift 3 || shift $#
OUT=gpurun_out; mkdir -p $OUT
for i in $(seq 1 $ROUNDS); do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-encoder "$@" > $OUT/ab_${TAG}_new_$i.json 2>> $OUT/ab_${TAG}.err
  env $ENVS timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-encoder "$@" > $OUT/ab_${TAG}_old_$i.json 2>> $OUT/ab_${TAG}.err
done
python - <<PY
import json, glob
for kind in ("new", "old"):
    for f in sorted(glob.glob("$OUT/ab_${TAG}_%s_*.json" % kind)):
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(kind, f.split("_")[-1], round(d["ms_per_step"], 3), round(d["p50_ms_per_image"], 3), d["clocks"]["sm_mhz"], {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
PY
