#!/bin/bash
# copy what the final GPU pass (scripts/gpu_final.sh) brought back in gpurun_out/ into the tracked profiles/
cd "$(dirname "$0")/.."
cp gpurun_out/profiles_out/*.txt profiles/
cp gpurun_out/profiles_out/traffic.json profiles/traffic.json
cp gpurun_out/bench_f64.json profiles/r02_bench_f64.json
cp gpurun_out/bench_f32.json profiles/r02_bench_f32.json
cp gpurun_out/bench_ref.json profiles/r02_bench_reference_arm.json
cp gpurun_out/pytest_gpu.log profiles/r02_pytest_gpu.log
cp gpurun_out/smoke.log profiles/r02_smoke.log
cp gpurun_out/r02_launches_f64.csv profiles/
cp gpurun_out/r02_compute_sanitizer.log gpurun_out/r02_gradient_timing.log gpurun_out/r02_fft_final.log profiles/
python scripts/launch_summary.py profiles/r02_launches_f64.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv  python bench.py --steps 2 --warmup 3 --skip cpu,other,map_joint,hmc --cg-iters 1   (first 1200 launches of the process: setup, warm-up and timed applies, e2e applies)" > profiles/r02_launches_f64_summary.txt
rm -f profiles/r02_stage_traffic_by_kind.log
sha256sum cmblensing.jl_b200/libcmbl_b200.so | cut -c1-16; grep binary_sha16 profiles/traffic.json
