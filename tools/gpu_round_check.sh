#!/bin/bash
# One gpurun call: GPU parity tests, the default bench line, A/B of the two kernels changed last (env toggles), the
# reference arm. usage: tools/gpu_round_check.sh TAG
set -u
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_${TAG}.log
timeout 400 python bench.py --steps 10 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
IR_VAE_CONVOUT_LEGACY=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-encoder > $OUT/bench_${TAG}_legacy_convout.json 2>> $OUT/bench_${TAG}.err; echo "bench legacy convout rc=$?"
IR_XATTN_QT=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-encoder > $OUT/bench_${TAG}_xattn_qt1.json 2>> $OUT/bench_${TAG}.err; echo "bench xattn qt1 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>> $OUT/bench_${TAG}.err; echo "reference rc=$?"
python - <<PY
import json
for n in ["", "_legacy_convout", "_xattn_qt1", "_reference"]:
    try:
        d = json.loads(open("$OUT/bench_${TAG}%s.json" % n).read().strip().splitlines()[-1])
        print(n or "default", d.get("ms_per_step"), d.get("value"), (d.get("e2e") or {}).get("value"), {k: round(v["ms_per_step"], 3) for k, v in (d.get("kernels") or {}).items()})
    except Exception as e:
        print(n, "ERR", e)
PY
