#!/bin/bash
# ncu --set full WITH source of single launches inside the resident 1024^2 step (plain schedule): self-attention, the two
# conv instantiations, the fp32-residual GEMM. usage: tools/gpu_step_ncu_src.sh TAG
set -u
TAG=${1:-src}
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-encoder --no-tiled --plain-schedule"
timeout 300 $CMD > $OUT/src_${TAG}_plain.json 2> $OUT/src_${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/src_${TAG}_plain.err; exit 1; }
full() {  # name, kernel regex, launch-skip, launch-count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$2" --launch-skip $3 --launch-count $4 -o $OUT/src_${TAG}_$1 -f $CMD > $OUT/src_${TAG}_ncu_$1.log 2>&1
  echo "set full $1 rc=$?"
}
full attn 'attn_tc_kernel' 20 1
full conv256 'gemm_tc_kernelILi256ELi0ELb1ELi2|gemm_tc_kernel<\(int\)256, \(int\)0, \(bool\)1, \(int\)2>' 12 2
full conv128 'gemm_tc_kernel<\(int\)128, \(int\)0, \(bool\)1, \(int\)2>' 2 1
full gnapply 'gn_apply_kernel' 20 1
ls -la $OUT/src_${TAG}_*.ncu-rep
