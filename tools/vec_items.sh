#!/bin/bash
export PYTHONUNBUFFERED=1
for nx in ${NXS:-56}; do for it in 1 2 4 8; do echo "nx=$nx VEC_ITEMS=$it"; FEMBRAIN_B200_VEC_ITEMS=$it timeout 200 python tools/spmv_variants.py $nx default; done; done
