#!/bin/bash
# attention A/B (debug build): the probe under two environments, interleaved: usage gpu_attn_ab.sh "ENV_A=.." "ENV_B=.." [rounds]
set -u
for i in $(seq 1 ${3:-2}); do
  env $1 timeout 200 python tools/gpu_attn_probe.py 2>&1 | grep "qscale 1.0\|PROBE" | sed "s/^/A $i: /"
  env $2 timeout 200 python tools/gpu_attn_probe.py 2>&1 | grep "qscale 1.0\|PROBE" | sed "s/^/B $i: /"
done
