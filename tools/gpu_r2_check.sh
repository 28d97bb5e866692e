#!/bin/bash
# round-2 iteration check: targeted kernel tests, the reference-golden parity tests, then the bench
# usage: tools/gpu_r2_check.sh TAG "<pytest -k expr for test_gpu_kernels>" [bench args...]
set -u
TAG=$1; KEXPR=$2; shift 2
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "$KEXPR" > $OUT/${TAG}_kernels.log 2>&1
echo "kernel tests rc=$?"; tail -6 $OUT/${TAG}_kernels.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -s > $OUT/${TAG}_parity.log 2>&1
echo "parity tests rc=$?"; grep -E "max-abs|PSNR|dB|passed|failed|Error|error" $OUT/${TAG}_parity.log | tail -14
timeout 900 python bench.py --steps 10 --warmup 3 "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "cli", (d.get("e2e_full_cli") or {}).get("value"), "crc", d["output_crc32"])
    print({k:(round(v["ms_per_step"],3), v["tflops"] and round(v["tflops"])) for k,v in d["kernels"].items()})
    t=d.get("tiled_2048")
    if t: print("tiled", t["ms_per_step"], t["value"], "e2e", t["e2e"]["value"], "crc", t["output_crc32"], {k:round(v,2) for k,v in t["phases_ms"].items()})
    for r in (d["roofline"].get("by_shape") or [])[:12]: print("  ", r["class"], r["M"], r["N"], r["K"], "n=%.0f us=%.1f ms=%.3f TF=%s" % (r["launches_per_step"], r["avg_us"], r["ms_per_step"], r["tflops"] and round(r["tflops"])))
except Exception as e:
    print("no bench line:", e)
PY
