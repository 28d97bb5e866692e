#!/bin/bash
# A/B of the direct row-owner GEMM epilogue (debug build: IR_DEBUG=1 python -m instarevive_b200.csrc.build --force)
# usage: tools/gpu_epi_ab.sh TAG [rounds]
set -u
TAG=$1; ROUNDS=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gemm or qkv or attention" > $OUT/${TAG}_kernels.log 2>&1
echo "kernel tests rc=$?"; tail -4 $OUT/${TAG}_kernels.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "swinir or dit_forward" > $OUT/${TAG}_parity.log 2>&1; echo "parity subset rc=$?"; tail -2 $OUT/${TAG}_parity.log
for m in ${MASKS:-0 15}; do
  echo "== ${ABVAR:-IR_GEMM_DIRECT}=$m"; env ${ABVAR:-IR_GEMM_DIRECT}=$m timeout 200 python tools/gpu_epi_ab.py 2>&1 | tee $OUT/${TAG}_epi_$m.log | tail -24
done
tools/gpu_ab.sh $TAG "${ABVAR:-IR_GEMM_DIRECT}=0" $ROUNDS --no-tiled
