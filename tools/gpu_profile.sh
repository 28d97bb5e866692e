#!/bin/bash
# ncu evidence for profiles/ (run on a B200 through gpurun): the bench command first exits 0 WITHOUT ncu, then
#  (1) launch list of one resident step (bench.py brackets it with cudaProfilerStart/Stop, outside every timed region)
#  (2) `--set full` captures of a few launches per kernel family (attention, DiT GEMMs, convs, GroupNorm apply, LN)
# usage: tools/gpu_profile.sh TAG [bench args...]
set -u
TAG=$1; shift
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-encoder $*"
timeout 300 $CMD > $OUT/prof_${TAG}_plain.json 2> $OUT/prof_${TAG}_plain.err || { echo "plain run failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/prof_${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, launch-skip, launch-count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$2" --launch-skip $3 --launch-count $4 -o $OUT/full_${TAG}_$1 -f $CMD > $OUT/prof_${TAG}_ncu_$1.log 2>&1
  echo "set full $1 rc=$?"
}
# gpurun copies back at most 64 MiB: ~2.5 MB per captured launch, so keep the captures to ~22 launches in total
full attn 'attn_tc_kernel' 20 1
full gemm 'gemm_tc_kernel' 120 6                           # one DiT block: qkv, proj, q_linear, cross proj, fc1, fc2
full conv 'gemm_tc_kernel' 282 6                           # two 3x3 convs of up.2 + the four phase convs of its upsample
full tail 'gemm_tc_kernel|conv_out_gather_kernel' 304 4    # last full-resolution convs, conv_out tap GEMM + gather
full ln 'ln_modulate_kernel|flash_attn_kernel' 40 3
full gn 'gn_apply_kernel' 22 2
if [ "$(du -sm $OUT | cut -f1)" -gt 60 ]; then rm -f $OUT/full_${TAG}_gn.ncu-rep; fi
if [ "$(du -sm $OUT | cut -f1)" -gt 60 ]; then rm -f $OUT/full_${TAG}_ln.ncu-rep; fi
du -sm $OUT
ls -la $OUT/full_${TAG}_*.ncu-rep $OUT/launches_${TAG}.csv
