#!/bin/bash
# ncu evidence for profiles/ (run on a B200 through gpurun): the bench command first exits 0 WITHOUT ncu, then
#  (1) launch list of one resident step (bench.py brackets it with cudaProfilerStart/Stop, outside every timed region)
#  (2) one `--set full` capture of a few launches of the tensor-core kernels (attention + GEMM/conv)
# usage: tools/gpu_profile.sh TAG [bench args...]
set -u
TAG=$1; shift
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $*"
timeout 300 $CMD > $OUT/prof_${TAG}_plain.json 2> $OUT/prof_${TAG}_plain.err || { echo "plain run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/prof_${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'attn_tc_kernel|gemm_tc_kernel|gn_apply_kernel|ln_modulate_kernel' --launch-skip 40 --launch-count 36 \
  -o $OUT/full_${TAG} -f $CMD > $OUT/prof_${TAG}_ncu2.log 2>&1
echo "set full rc=$?"
ls -la $OUT/full_${TAG}.ncu-rep $OUT/launches_${TAG}.csv
