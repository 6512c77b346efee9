#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
R1="CMBL_B200_ALLOW_MISSING=1 CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_r1.so"
for v in "CMBL_COL_JN_RED=3" "CMBL_COL_JN_RED=1" "CMBL_COL_JN_RED=2" "$R1"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/${v:0:17} /"; done; done; done > gpurun_out/ab_red.log 2>&1
grep "ms/apply" gpurun_out/ab_red.log
for v in "X=1" "$R1"; do NB=1 NPOL=1 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_" | sed "s/^/${v:0:8} /"; NB=1 NPOL=1 env $v timeout 120 python scripts/time_apply.py f32 0 2>&1 | grep "ms/apply" | sed "s/^/${v:0:8} /";  N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply" | sed "s/^/${v:0:8} /"; N=256 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply" | sed "s/^/${v:0:8} /"; done
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_bench_contract.py > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
