#!/bin/bash
# same-box A/B against the round-1 binary (after moving the rare J path out of line) + ncu captures of the final binary
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
R1="CMBL_B200_ALLOW_MISSING=1 CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_r1.so"
for v in "X=1" "$R1"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/${v:0:8} /"; done; done; done > gpurun_out/ab_r1b.log 2>&1
grep "ms/apply\|flow_cols" gpurun_out/ab_r1b.log
for v in "X=1" "$R1"; do NB=1 NPOL=1 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/${v:0:8} /"; N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/${v:0:8} /"; done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_kernels.py -m gpu -x -q -k "concurrent or fast_path or host_pipeline or pullback or cl_to_cov or headline" > gpurun_out/pytest7.log 2>&1; tail -3 gpurun_out/pytest7.log
CMBL_COL_JN_RED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lenseflow_fast_path or host_pipeline or pullback" > gpurun_out/pytest7b.log 2>&1; tail -2 gpurun_out/pytest7b.log
timeout 600 ncu --set full --clock-control none --import-source on -s 36 -c 4 -o gpurun_out/r02_ncu_flow_f64 python scripts/ncu_target.py f64 fwd > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -s 36 -c 4 -o gpurun_out/r02_ncu_flow_f32 python scripts/ncu_target.py f32 fwd > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -s 6 -c 6 -o gpurun_out/r02_ncu_adj_f64 python scripts/ncu_target.py f64 adj > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
ls -la gpurun_out/*.ncu-rep
