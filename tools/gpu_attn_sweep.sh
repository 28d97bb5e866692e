#!/bin/bash
# attention sweep (debug build): IR_ATTN_EMU x IR_ATTN_ORDER at B1 T4096 / B8 T1024 / B8 T4096
set -u
for o in 3 2; do for e in 2 3 4; do
  IR_ATTN_EMU=$e IR_ATTN_ORDER=$o timeout 200 python tools/gpu_attn_probe.py 2>&1 | grep "qscale 1.0" | grep -v "B2 "
done; done
