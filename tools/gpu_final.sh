#!/bin/bash
# End-of-round verification in one gpurun call: descriptor probe, GPU parity tests, smoke(), the default bench line and the
# tiled workload on one GPU. usage: tools/gpu_final.sh TAG
set -u
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
if [ -x tools/probes/bin/desc_probe ]; then timeout 60 tools/probes/bin/desc_probe > $OUT/desc_probe_${TAG}.log 2>&1; echo "probe rc=$?"; cat $OUT/desc_probe_${TAG}.log; fi
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_${TAG}.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_${TAG}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_${TAG}.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_${TAG}.log
timeout 400 python bench.py > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; echo "bench rc=$?"
timeout 400 python bench.py --workload tiled --steps 5 --warmup 3 --no-cpu-baseline --no-encoder > $OUT/bench_${TAG}_tiled.json 2>> $OUT/bench_${TAG}.err; echo "bench tiled rc=$?"
python - <<PY
import json
for n in ["", "_tiled"]:
    try:
        d = json.loads(open("$OUT/bench_${TAG}%s.json" % n).read().strip().splitlines()[-1])
        print(n or "default", round(d["ms_per_step"], 3), round(d["value"], 2), round(d["e2e"]["value"], 2), d["output_crc32"], d["clocks"], {k: round(v["ms_per_step"], 3) for k, v in (d.get("kernels") or {}).items()})
    except Exception as e:
        print(n, "ERR", e)
PY
