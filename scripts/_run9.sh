cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
{ timeout 600 python scripts/time_map_joint.py f32 512 P 8 2; timeout 600 python scripts/time_map_joint.py f64 512 P 8 2; } > gpurun_out/cfg5_hmc.log 2>&1
cat gpurun_out/cfg5_hmc.log
timeout 600 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 300 gpurun_out/bench_f64.err; python -c "
import json; d=json.load(open('gpurun_out/bench_f64.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])"
