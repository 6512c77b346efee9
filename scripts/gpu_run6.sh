#!/bin/bash
# same-box A/B of the stage kernels against the round-1 binary; ncu captures of the final binary
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
R1="CMBL_B200_ALLOW_MISSING=1 CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_r1.so"
for rep in 1 2; do for v in "X=1" "$R1"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/[$rep] ${v:0:8} /"; done; done; done; done > gpurun_out/ab_r1.log 2>&1
grep "ms/apply\|flow_cols\|flow_rows" gpurun_out/ab_r1.log
for v in "X=1" "$R1"; do NB=1 NPOL=1 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_" | sed "s/^/${v:0:8} /"; N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_" | sed "s/^/${v:0:8} /"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_f64.csv python bench.py --steps 2 --warmup 3 --skip cpu,other,map_joint,hmc --cg-iters 1 > /dev/null 2> gpurun_out/ncu_launch.err; tail -2 gpurun_out/ncu_launch.err
ncu --set full --clock-control none --import-source on -k regex:FastColBody -s 60 -c 2 -o gpurun_out/ncu_cols_f64 python scripts/time_apply.py f64 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:TmaRowBody -s 60 -c 2 -o gpurun_out/ncu_rows_f64 python scripts/time_apply.py f64 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:FastColBody -s 60 -c 2 -o gpurun_out/ncu_cols_f32 python scripts/time_apply.py f32 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:R2CColBody|C2CRowBody|C2RColBody" -s 4 -c 4 -o gpurun_out/ncu_fft_f64 python scripts/time_cg.py f64 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_kernels.py -m gpu -x -q -k "pullback or finite_difference or logpdf_mixed or hmc or sample_joint or map_joint" > gpurun_out/pytest6.log 2>&1; tail -3 gpurun_out/pytest6.log
for v in "CMBL_GRAD_FUSED=1" "CMBL_GRAD_FUSED=0"; do env $v timeout 600 python scripts/time_map_joint.py f64 1024 P 8 1 2>&1 | grep -i "gradient\|HMC" | sed "s/^/$v /"; env $v timeout 600 python scripts/time_map_joint.py f64 512 P 8 1 2>&1 | grep -i "gradient of\|HMC" | sed "s/^/$v N=512 /"; done
