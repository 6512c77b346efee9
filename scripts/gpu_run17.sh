#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "X=1" "CMBL_FLOW_PF=0" "CMBL_COL_PAIR=1 CMBL_FLOW_PF=0" "CMBL_COL_PAIR=1" "CMBL_FLOW_PF=1"; do
  echo "== $v"
  env $v timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct --clock-control none -s 39 -c 1 python scripts/ncu_target.py f64 fwd 2>&1 | grep -i "FastCol\|dram__\|duration\|hit_rate"
  env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols"
done > gpurun_out/ab_pf_traffic.log 2>&1
cat gpurun_out/ab_pf_traffic.log
