#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
R1="CMBL_B200_ALLOW_MISSING=1 CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_r1.so"
for v in "X=1" "CMBL_COL_JN_POLLS=0" "$R1"; do for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/${v:0:19} /"; done; done; done > gpurun_out/ab_r1c.log 2>&1
grep "ms/apply\|flow_cols" gpurun_out/ab_r1c.log
for v in "X=1" "$R1"; do NB=1 NPOL=1 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/${v:0:8} /"; N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/${v:0:8} /"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or host_pipeline" > gpurun_out/pytest8.log 2>&1; tail -4 gpurun_out/pytest8.log
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 1500 gpurun_out/bench_f64.json; tail -5 gpurun_out/bench_f64.err
cap() { # name, skip, count, args...
  local name=$1 s=$2 c=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -s $s -c $c -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  ls -la gpurun_out/$name.ncu-rep
}
cap r02_ncu_flow_f64 38 2 python scripts/ncu_target.py f64 fwd
cap r02_ncu_flow_f32 38 2 python scripts/ncu_target.py f32 fwd
cap r02_ncu_adj_f64 12 2 python scripts/ncu_target.py f64 adj
du -sm gpurun_out
