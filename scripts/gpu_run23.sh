#!/bin/bash
# row pass of the general transforms: strided loads in flight per thread (4 vs 16)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
for shape in "N=1024 C=16" "N=2048 C=6"; do for dt in f64 f32; do for v in "X=1" "CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_rowunr16.so"; do
  echo "== $shape $dt $v"
  env $shape $v timeout 300 python scripts/time_fft.py $dt 2>&1 | grep "us\|Error"
done; done; done
} > gpurun_out/r02_fft_row_unroll.log 2>&1
cat gpurun_out/r02_fft_row_unroll.log
