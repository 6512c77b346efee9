cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
{ timeout 600 python scripts/time_map_joint.py f64 1024 P 8 2; timeout 600 python scripts/time_map_joint.py f32 1024 P 8 2; } > gpurun_out/map_joint.log 2>&1
cat gpurun_out/map_joint.log
timeout 600 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 300 gpurun_out/bench_f64.err
