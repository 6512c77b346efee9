cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for dt in f64 f32; do
  timeout 600 ncu --set full --clock-control none --import-source on -s 36 -c 4 -f -o gpurun_out/r01_tma_$dt python scripts/ncu_target.py $dt fwd > gpurun_out/ncu_$dt.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches_f64.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --cg-iters 1 > gpurun_out/b.log 2>&1
ls -la gpurun_out | tail -8
