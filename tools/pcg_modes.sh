#!/bin/bash
# developer tool: PCG options side by side (GPU box).  MODES="kernels fused persistent", PDLS="1 0"
export PYTHONUNBUFFERED=1
for nx in ${NXS:-56}; do
for pdl in ${PDLS:-1}; do
for mode in ${MODES:-kernels}; do
  echo "nx=$nx FEMBRAIN_B200_PCG=$mode PDL=$pdl"; FEMBRAIN_B200_PDL=$pdl FEMBRAIN_B200_PCG=$mode timeout 200 python tools/spmv_variants.py $nx default
done; done; done
