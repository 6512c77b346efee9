#!/bin/bash
# column kernel: do the sweeps and the epilogue of the co-resident blocks overlap?  ablations + start-up stagger
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
for v in "X=1" "CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_colabl1.so" "CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_colabl2.so" "CMBL_FLOW_STAGGER_NS=2000" "CMBL_FLOW_STAGGER_NS=4000" "CMBL_FLOW_STAGGER_NS=6500" "CMBL_FLOW_STAGGER_NS=10000"; do
  for op in 0 1; do
    echo "== f64 op$op $v"
    env $v CMBL_B200_ALLOW_MISSING=1 timeout 300 python scripts/time_apply.py f64 $op 2>&1 | grep "ms/apply\|flow_"
  done
done
for v in "X=1" "CMBL_FLOW_STAGGER_NS=3000" "CMBL_FLOW_STAGGER_NS=6000"; do
  echo "== f32 op0 $v"
  env $v timeout 300 python scripts/time_apply.py f32 0 2>&1 | grep "ms/apply\|flow_"
done
} > gpurun_out/r02_col_overlap.log 2>&1
cat gpurun_out/r02_col_overlap.log
