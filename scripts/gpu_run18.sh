#!/bin/bash
# transform length 2048 on the fast stage kernels: GPU parity, then A/B against the generic kernels (CMBL_FLOW_FAST=0)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_kernels.py -m gpu -x -q -k "fast_path or all_ops" 2>&1 | tail -3
for dt in f64 f32; do for op in 0 1; do
  for v in "X=1" "CMBL_FLOW_FAST=0"; do
    echo "== N=2048 IQU nb=2 $dt op$op $v"
    env $v N=2048 NB=2 NPOL=3 timeout 300 python scripts/time_apply.py $dt $op 2>&1 | grep "ms/apply\|flow_\|layout"
  done
done; done
echo "== N=1024 QU nb=8 f64 (regression check)"
timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_"
echo "== map_joint 2048 IQU nb=1 (bench section)"
timeout 600 python scripts/time_map_joint.py f64 2048 IP 1 2 2>&1 | tail -6
} > gpurun_out/r02_fast_2048.log 2>&1
cat gpurun_out/r02_fast_2048.log
