#!/bin/bash
# general-transform column passes on the persistent tile kernels (fft2d_fast.cuh) + PDL defaults: GPU suite, then A/B
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for dt in f64 f32; do for v in "CMBL_FFT_FAST=1" "CMBL_FFT_FAST=0"; do
  echo "== CG iteration $dt $v"
  env $v timeout 300 python scripts/time_cg.py $dt 2>&1 | grep "per CG iteration\|fft2_rows\|fft_cols\|rfft2_cols\|irfft2_cols\|flow_rows\|flow_cols\|layout"
done; done
for shape in "N=1024 NB=8 NPOL=2" "N=1024 NB=1 NPOL=1"; do for dt in f64 f32; do for op in 0 1; do
  echo "== $shape $dt op$op (default PDL rule)"
  env $shape timeout 300 python scripts/time_apply.py $dt $op 2>&1 | grep "ms/apply"
done; done; done
} > gpurun_out/r02_fft_fast_cols.log 2>&1
cat gpurun_out/r02_fft_fast_cols.log
