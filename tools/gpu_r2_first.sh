#!/bin/bash
# round-2 first check: the new full-size parity tests, then the default bench (with the tiled_2048 block)
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q -s -k "fullsize or 1024 or 2048 or caption_cache or batch8" > $OUT/r2_tests_fullsize.log 2>&1
echo "fullsize tests rc=$?"; tail -15 $OUT/r2_tests_fullsize.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2_bench_first.json 2> $OUT/r2_bench_first.err
echo "bench rc=$?"; tail -c 3000 $OUT/r2_bench_first.json; tail -5 $OUT/r2_bench_first.err
