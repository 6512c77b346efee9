#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for name in base row3; do
  if [ $name = base ]; then v="X=1"; else v="CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_$name.so"; fi
  for d in f64; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | grep "ms/apply\|flow_" | sed "s/^/$name /"; done; done
  N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_rows" | sed "s/^/$name /"
done > gpurun_out/ab_row3.log 2>&1
cat gpurun_out/ab_row3.log
CMBL_COL_JN_RED=0 timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_kernels.py -m gpu -x -q > gpurun_out/pytest_allfallback.log 2>&1; tail -3 gpurun_out/pytest_allfallback.log
