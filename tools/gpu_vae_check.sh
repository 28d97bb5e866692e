#!/bin/bash
# VAE iteration check: parity tests touching the VAE, encode / decode whole-call times, ncu launch list of the decoder
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-vae}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "vae or encode or process or decoder" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
timeout 300 python tools/gpu_encode_profile.py 2>&1 | grep "ms per call"
timeout 300 python tools/gpu_decode_profile.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/decode_launches.csv python tools/gpu_decode_profile.py > $OUT/decode_ncu.log 2>&1; echo "ncu rc=$?"
