#!/bin/bash
# developer tool: PCG schedules side by side (GPU box).  MODES="kernels fused_nograph fused persistent", MINBS="5 4"
export PYTHONUNBUFFERED=1
for nx in ${NXS:-56}; do
for minb in ${MINBS:-5}; do
for mode in ${MODES:-kernels fused}; do
  echo "nx=$nx FEMBRAIN_B200_PCG=$mode MINB=$minb"; FEMBRAIN_B200_MINB=$minb FEMBRAIN_B200_PCG=$mode timeout 200 python tools/spmv_variants.py $nx default
done; done; done
