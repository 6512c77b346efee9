#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for v in "CMBL_COL_PAIR=1" "CMBL_COL_PAIR=0"; do for d in f64 f32; do env $v timeout 120 python scripts/time_apply.py $d 0 2>&1 | grep "ms/apply\|flow_" | sed "s/^/$v /"; done; N=512 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/$v /"; NB=1 env $v timeout 120 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/$v /"; done > gpurun_out/ab_pair.log 2>&1
cat gpurun_out/ab_pair.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "concurrent or fast_path or headline or host_pipeline or pullback or golden" > gpurun_out/pytest16.log 2>&1; tail -4 gpurun_out/pytest16.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -s 38 -c 2 python scripts/ncu_target.py f64 fwd 2>&1 | grep -i "FastCol\|TmaRow\|dram__\|duration\|hit_rate" | head -12
