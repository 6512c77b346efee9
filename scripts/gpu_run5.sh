#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "X=1" "CMBL_FLOW_PF=0" "CMBL_COL_JN_RED=0"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/$v /"; done; done; done > gpurun_out/ab5.log 2>&1
grep "ms/apply\|flow_" gpurun_out/ab5.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi_c.py tests/test_kernels.py -m gpu -x -q -s -k "comm_abi or abi_from_c or concurrent or host_pipeline or fast_path or max_lensing" > gpurun_out/pytest5.log 2>&1; tail -5 gpurun_out/pytest5.log
CMBL_COL_JN_RED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lenseflow_fast_path or host_pipeline" > gpurun_out/pytest5b.log 2>&1; tail -2 gpurun_out/pytest5b.log
timeout 300 python scripts/time_cg.py f64 > gpurun_out/time_cg.log 2>&1; tail -25 gpurun_out/time_cg.log
