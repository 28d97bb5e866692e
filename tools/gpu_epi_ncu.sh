#!/bin/bash
# ncu --set full of one isolated DiT GEMM launch per epilogue variant (debug build). usage: tools/gpu_epi_ncu.sh TAG CASE "MASKS"
set -u
TAG=$1; CASE=$2; MASKS=${3:-"0 15"}
OUT=gpurun_out; mkdir -p $OUT
for m in $MASKS; do
  IR_GEMM_DIRECT=$m timeout 120 python tools/gpu_gemm_ncu_case.py $CASE > /dev/null 2>&1 || { echo "plain run failed ($m)"; exit 1; }
  IR_GEMM_DIRECT=$m timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o $OUT/epi_${TAG}_${CASE}_$m -f python tools/gpu_gemm_ncu_case.py $CASE > $OUT/epi_${TAG}_${CASE}_$m.log 2>&1
  echo "ncu $CASE mask $m rc=$?"
done
ls -la $OUT/epi_${TAG}_*.ncu-rep
