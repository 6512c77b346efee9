#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "X=1" "CMBL_COL_PGROUP=0" "CMBL_FLOW_PF=0" "CMBL_COL_PGROUP=0 CMBL_FLOW_PF=0"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/$v /"; done; done; done > gpurun_out/ab_pg.log 2>&1
grep "ms/apply\|flow_" gpurun_out/ab_pg.log
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_bench_contract.py > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 7000 gpurun_out/bench_f64.json; tail -5 gpurun_out/bench_f64.err
