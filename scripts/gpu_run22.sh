#!/bin/bash
# standalone 2-D transforms: persistent column kernels on/off, row pass with the skewed interleaved tile
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "rfft2 or fft or gradientf" 2>&1 | tail -3
for shape in "N=1024 C=16" "N=2048 C=6" "N=512 C=16" "N=256 C=16"; do for dt in f64 f32; do for v in "CMBL_FFT_FAST=1" "CMBL_FFT_FAST=1 CMBL_FFT_ROW_PF=0" "CMBL_FFT_FAST=0 CMBL_FFT_ROW_PF=0"; do
  echo "== $shape $dt $v"
  env $shape $v timeout 300 python scripts/time_fft.py $dt 2>&1 | grep "us\|Error"
done; done; done
} > gpurun_out/r02_fft_standalone.log 2>&1
cat gpurun_out/r02_fft_standalone.log
