#!/bin/bash
# developer tool: effect of the persisting-L2 window on the PCG iteration (GPU box)
export PYTHONUNBUFFERED=1
for nx in ${NXS:-56}; do
for pin in 0 0.5 1; do
  echo "nx=$nx FEMBRAIN_B200_L2PIN=$pin"; FEMBRAIN_B200_L2PIN=$pin timeout 300 python tools/spmv_variants.py $nx rows3_5
done; done
