cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 600 gpurun_out/bench_f64.err
{
for op in 0 1; do timeout 300 python scripts/time_apply.py f64 $op; done
timeout 300 python scripts/time_cg.py f64
N=2048 NPOL=3 NB=1 timeout 300 python scripts/time_apply.py f64 0
N=2048 NPOL=3 NB=1 timeout 300 python scripts/time_apply.py f32 0
N=512 NPOL=2 NB=8 timeout 300 python scripts/time_apply.py f64 0
} > gpurun_out/times2.log 2>&1
cat gpurun_out/times2.log
