#!/bin/bash
# encoder per-launch view: the profile tool, then an ncu launch list (gpu__time_duration) of the same script
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python tools/gpu_encode_profile.py > $OUT/encode_profile.log 2>&1; echo "profile rc=$?"; tail -60 $OUT/encode_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/encode_launches.csv python tools/gpu_encode_profile.py > $OUT/encode_ncu.log 2>&1; echo "ncu rc=$?"
