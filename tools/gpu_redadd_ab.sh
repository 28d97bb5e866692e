#!/bin/bash
# reduce-add epilogue: correctness (kernel tests + DiT goldens), then interleaved A/B of the bench (debug build: env switches live)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "f32_gate" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "dit" 2>&1 | tail -3
bash tools/gpu_ab.sh redadd "IR_GEMM_REDADD=0" 2 --no-tiled
