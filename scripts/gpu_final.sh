#!/bin/bash
# final single-GPU pass of round 2: smoke, full GPU suite, both bench arms, both precisions, launch list, ncu captures of the final binary
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
sha256sum cmblensing.jl_b200/libcmbl_b200.so | cut -c1-16 > gpurun_out/binary_sha16.txt; cat gpurun_out/binary_sha16.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -2 gpurun_out/bench_f64.err
timeout 900 python bench.py --dtype f32 --skip cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; tail -2 gpurun_out/bench_f32.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_f64.csv python bench.py --steps 2 --warmup 3 --skip cpu,other,map_joint,hmc --cg-iters 1 > /dev/null 2> gpurun_out/ncu_launch.err; tail -1 gpurun_out/ncu_launch.err
cap() { local name=$1 s=$2 c=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -s $s -c $c -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1; ls -la gpurun_out/$name.ncu-rep; }
cap r02_ncu_flow_f64 38 2 python scripts/ncu_target.py f64 fwd
cap r02_ncu_flow_f32 38 2 python scripts/ncu_target.py f32 fwd
cap r02_ncu_adj_f64 12 2 python scripts/ncu_target.py f64 adj
cap r02_ncu_fft_f64 2 3 python scripts/ncu_target.py f64 adj
# DRAM traffic of one whole RK4 step (4 row + 4 column launches: stage kinds 0,1,1,2 in some rotation), metrics only, both precisions
for dt in f64 f32; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_op_read_hit_rate.pct --clock-control none -s 20 -c 8 --csv --log-file gpurun_out/r02_stage_traffic_by_kind_$dt.csv python scripts/ncu_target.py $dt fwd > /dev/null 2>&1
done
python scripts/make_profiles.py r02 gpurun_out/profiles_out > gpurun_out/make_profiles.log 2>&1; tail -2 gpurun_out/make_profiles.log
rm -f gpurun_out/r02_ncu_flow_f32.ncu-rep gpurun_out/r02_ncu_adj_f64.ncu-rep gpurun_out/r02_ncu_fft_f64.ncu-rep      # keep one .ncu-rep (64 MiB limit on what travels back)
# gradient of logpdf(Mixed) / HMC update at the bench workload, and the standalone transforms
timeout 600 python scripts/time_map_joint.py f64 1024 P 8 1 2>&1 | tail -5 > gpurun_out/r02_gradient_timing.log; tail -4 gpurun_out/r02_gradient_timing.log | cut -c1-200
{ for dt in f64 f32; do timeout 300 python scripts/time_fft.py $dt 2>&1 | grep "us"; done; } > gpurun_out/r02_fft_final.log 2>&1
# compute-sanitizer on the small workload (all kernel families, incl. transform length 2048 and the persistent transform column kernels)
{ echo "compute-sanitizer --tool {memcheck,racecheck,synccheck} python scripts/sanitize_target.py   (256x256 QU + 64x64 IQU generic + 2048x256 I fp64 + 256x2048 QU fp32: all four flow ops, pullback, 3 CG iterations, get_max_lensing_step)"
  echo "libcmbl_b200.so sha256[:16] = $(cat gpurun_out/binary_sha16.txt)"
  for tool in memcheck racecheck synccheck; do
    timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_target.py 2>&1 | grep -E " ok |ERROR SUMMARY|RACECHECK SUMMARY|Invalid|Race|hazard" | head -20 | sed "s/^/[$tool] /"
  done; } > gpurun_out/r02_compute_sanitizer.log 2>&1; tail -3 gpurun_out/r02_compute_sanitizer.log
python - <<'PY'
import json
for n in ("ref", "f64", "f32"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["value"], 3), d["unit"], "e2e", round(d["e2e"]["value"], 3), "roofline", (d.get("roofline") or {}).get("frac"), "cg", (d.get("cg") or {}).get("value"), "mj", (d.get("map_joint") or {}).get("value"), "hmc", (d.get("hmc") or {}).get("value"))
    except Exception as e:
        print(n, "FAILED", e)
PY
du -sm gpurun_out
