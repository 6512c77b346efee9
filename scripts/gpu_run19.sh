#!/bin/bash
# 2048 fast path: bench sections that use it (map_joint at Nside=2048 IQU), pullback parity on the GPU, HMC sanity at 2048
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_kernels.py -m gpu -x -q -k "pullback" 2>&1 | tail -3
timeout 900 python bench.py --skip cpu,other,hmc,cg > gpurun_out/bench_2048.json 2> gpurun_out/bench_2048.err; tail -2 gpurun_out/bench_2048.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2048.json')); print(d['value'], d['e2e']['value']); print(json.dumps(d['map_joint']))"
echo "== time_map_joint f64 2048 P nb=1 (fast kernels)"
timeout 900 python scripts/time_map_joint.py f64 2048 P 1 1 2>&1 | tail -7
echo "== time_map_joint f64 2048 IP nb=1, generic kernels (CMBL_FLOW_FAST=0)"
CMBL_FLOW_FAST=0 timeout 900 python scripts/time_map_joint.py f64 2048 IP 1 1 2>&1 | tail -6
} > gpurun_out/r02_fast_2048_b.log 2>&1
cat gpurun_out/r02_fast_2048_b.log
