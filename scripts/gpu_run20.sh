#!/bin/bash
# programmatic dependent launch between the stage kernels (CMBL_PDL=1), by workload shape
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
for shape in "N=1024 NB=1 NPOL=1" "N=1024 NB=8 NPOL=2" "N=512 NB=8 NPOL=2" "N=2048 NB=1 NPOL=3" "N=256 NB=1 NPOL=2"; do
  for dt in f64 f32; do for op in 0 1; do for pdl in 0 1; do
    echo "== $shape $dt op$op CMBL_PDL=$pdl"
    env $shape CMBL_PDL=$pdl timeout 300 python scripts/time_apply.py $dt $op 2>&1 | grep "ms/apply"
  done; done; done
done
} > gpurun_out/r02_pdl_by_shape.log 2>&1
cat gpurun_out/r02_pdl_by_shape.log
