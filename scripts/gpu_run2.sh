#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "X=1" "CMBL_COL_JN_RED=0" "CMBL_COL_CONTIG=1" "CMBL_COL_JN_RED=1"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/$v /"; done; done; done > gpurun_out/ab_jn.log 2>&1
grep "ms/apply\|flow_" gpurun_out/ab_jn.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or fast_path or headline" > gpurun_out/pytest_gpu2.log 2>&1; tail -4 gpurun_out/pytest_gpu2.log
CMBL_COL_JN_RED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path" > gpurun_out/pytest_gpu3.log 2>&1; tail -2 gpurun_out/pytest_gpu3.log
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 5000 gpurun_out/bench_f64.json; tail -5 gpurun_out/bench_f64.err
