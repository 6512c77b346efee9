#!/bin/bash
# transform length 2048: 4 instead of 2 epilogue units in flight per thread in the column kernel (one block per SM there)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
V="CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_unr2048_4.so"
{
for dt in f64 f32; do for op in 0 1; do for v in "X=1" "$V"; do
  echo "== N=2048 IQU nb=2 $dt op$op $v"
  env $v N=2048 NB=2 NPOL=3 timeout 300 python scripts/time_apply.py $dt $op 2>&1 | grep "ms/apply\|flow_cols"
done; done; done
} > gpurun_out/r02_2048_epilogue_unroll.log 2>&1
cat gpurun_out/r02_2048_epilogue_unroll.log
