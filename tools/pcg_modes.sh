#!/bin/bash
# developer tool: three-kernel PCG vs persistent cooperative kernel (GPU box)
export PYTHONUNBUFFERED=1
for nx in ${NXS:-56}; do
for mode in kernels persistent; do
  echo "nx=$nx FEMBRAIN_B200_PCG=$mode"; FEMBRAIN_B200_PCG=$mode timeout 200 python tools/spmv_variants.py $nx rows3_5
done; done
