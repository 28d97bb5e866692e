#!/bin/bash
# round-2 ncu evidence for profiles/ (run on a B200 through gpurun). The bench command first exits 0 WITHOUT ncu, then:
#  (1) launch list of one resident 1024x1024 step (bench.py brackets it with cudaProfilerStart/Stop, outside timed regions)
#  (2) DRAM-bytes pass of the same step + the library's per-launch shape log (tools/ncu_traffic_by_shape.py joins them)
#  (3) `--set full` captures of the top kernels (self-attention, cross-attention, the six linear GEMMs of one DiT block)
# --plain-schedule: one stream, program order, no CUDA graph, so that ncu's launch order equals the shape log's
set -u
TAG=${1:-r02}
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-encoder --no-tiled --plain-schedule"
timeout 300 $CMD --dump-shapes $OUT/shapes_${TAG}.json > $OUT/prof_${TAG}_plain.json 2> $OUT/prof_${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/prof_${TAG}_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/launches_${TAG}.csv $CMD > $OUT/prof_${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $OUT/traffic_${TAG}.csv $CMD > $OUT/prof_${TAG}_ncu2.log 2>&1
echo "traffic pass rc=$?"
full() {  # name, kernel regex, launch-skip, launch-count
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"$2" --launch-skip $3 --launch-count $4 -o $OUT/full_${TAG}_$1 -f $CMD > $OUT/prof_${TAG}_ncu_$1.log 2>&1
  echo "set full $1 rc=$?"
}
full attn 'attn_tc_kernel' 20 1
full xattn 'xattn_tc_kernel' 20 1
full gemm 'gemm_tc_kernel' 120 6      # one DiT block: qkv, proj, q_linear, cross proj, fc1, fc2
du -sm $OUT
ls -la $OUT/full_${TAG}_*.ncu-rep $OUT/launches_${TAG}.csv $OUT/traffic_${TAG}.csv $OUT/shapes_${TAG}.json
