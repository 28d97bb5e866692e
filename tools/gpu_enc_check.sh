#!/bin/bash
# encoder iteration check: encode parity tests (reference goldens), VAE decode goldens, per-launch profile
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-enc}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "vae or encode or process or decoder" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -15 $OUT/${TAG}_tests.log
timeout 300 python tools/gpu_encode_profile.py > $OUT/${TAG}_profile.log 2>&1; echo "profile rc=$?"; cat $OUT/${TAG}_profile.log | tail -40
