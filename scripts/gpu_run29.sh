#!/bin/bash
# column kernel memory phase vs the pure-streaming ceiling: prefetch placement with and without the sweeps
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
A="CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_colabl1.so CMBL_B200_ALLOW_MISSING=1"
{
for pf in 0 1 2 3 4; do
  echo "== no sweeps (ablation 1), CMBL_FLOW_PF=$pf"
  env $A CMBL_FLOW_PF=$pf timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "flow_cols"
done
for pf in 0 3; do
  echo "== full kernel, CMBL_FLOW_PF=$pf"
  env CMBL_FLOW_PF=$pf timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols"
done
for v in "CMBL_FLOW_BLOCKS_PER_SM=2" "CMBL_FLOW_BLOCKS_PER_SM=2 CMBL_FLOW_PF=0"; do
  echo "== no sweeps (ablation 1), $v"
  env $A $v timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "flow_cols"
  echo "== full kernel, $v"
  env $v timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols"
done
} > gpurun_out/r02_col_memphase.log 2>&1
cat gpurun_out/r02_col_memphase.log
