#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or fast_path or headline or host_pipeline or golden" > gpurun_out/pytest_ws.log 2>&1; tail -4 gpurun_out/pytest_ws.log
CMBL_COL_PIPE=0 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or lenseflow_fast_path or host_pipeline" > gpurun_out/pytest_nows.log 2>&1; tail -4 gpurun_out/pytest_nows.log
for v in "X=1" "CMBL_COL_PIPE=0" "CMBL_COL_JN_RED=0"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/$v /"; done; done; done > gpurun_out/ab_ws.log 2>&1
grep "ms/apply\|flow_" gpurun_out/ab_ws.log
N=512 timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_"
NB=1 NPOL=1 timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_"
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 6000 gpurun_out/bench_f64.json; tail -5 gpurun_out/bench_f64.err
