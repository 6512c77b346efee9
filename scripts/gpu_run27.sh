#!/bin/bash
# row tile of 64 KB at Nx = 1024 fp64 (G = 8: 256-byte runs in the column kernel) against the default 32 KB (G = 4: 128-byte runs)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
V="CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_rowtile64.so"
{
env $V timeout 600 python -m pytest tests/test_kernels.py -m gpu -x -q -k "fast_path or pullback or precompute" 2>&1 | tail -2
for v in "X=1" "$V"; do
  for op in 0 1; do
    echo "== f64 op$op $v"
    env $v timeout 300 python scripts/time_apply.py f64 $op 2>&1 | grep "ms/apply\|flow_\|layout"
  done
  echo "== CG f64 $v"
  env $v timeout 300 python scripts/time_cg.py f64 2>&1 | grep "per CG iteration"
  echo "== N=1024 NB=1 NPOL=1 f64 op0 $v"
  env $v N=1024 NB=1 NPOL=1 timeout 300 python scripts/time_apply.py f64 0 2>&1 | grep "ms/apply"
done
} > gpurun_out/r02_row_tile_64k.log 2>&1
cat gpurun_out/r02_row_tile_64k.log
