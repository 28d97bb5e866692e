#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gpu_decode_profile.py > $OUT/decode_profile.log 2>&1; echo "profile rc=$?"; tail -3 $OUT/decode_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/decode_launches.csv python tools/gpu_decode_profile.py > $OUT/decode_ncu.log 2>&1; echo "ncu rc=$?"
