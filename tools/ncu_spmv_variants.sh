#!/bin/bash
# GPU box: one `ncu --set full` capture of the solver's product kernel per variant, inside a real step of the 1M-tet bench
# workload (launch 200 of that kernel), reports under gpurun_out/.  ~45 s per variant.
#   gpurun --timeout 400 -- 'bash tools/ncu_spmv_variants.sh r02 default l2evict tma sym'
tag=${1:-r02}; shift
mkdir -p gpurun_out
for var in "${@:-default}"; do
  unset FEMBRAIN_B200_SPMV FEMBRAIN_B200_L2EVICT
  case $var in
    default) kern='regex:k_spmv_rows3' ;;
    l2evict) kern='regex:k_spmv_rows3'; export FEMBRAIN_B200_L2EVICT=1 ;;
    tma)     kern='regex:k_spmv_tma';   export FEMBRAIN_B200_SPMV=tma ;;
    sym)     kern='regex:k_spmv_sym';   export FEMBRAIN_B200_SPMV=sym ;;
    *) echo "unknown variant $var"; continue ;;
  esac
  timeout 150 ncu --set full --import-source on --clock-control none -k "$kern" --launch-skip 200 -c 1 \
      -o gpurun_out/${tag}_spmv_${var} -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_${var}.log 2>&1
  echo "$var rc=$?"
done
