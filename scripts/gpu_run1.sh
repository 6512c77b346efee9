#!/bin/bash
# GPU pass 1 of round 2: parity suite, A/B of the column-kernel tile order, bench line
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q -s --deselect tests/test_bench_contract.py > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for c in 1 0; do for d in f64 f32; do for op in 0 1; do CMBL_COL_CONTIG=$c timeout 120 python scripts/time_apply.py $d $op 2>&1 | sed "s/^/contig=$c /"; done; done; done > gpurun_out/ab_contig.log 2>&1
cat gpurun_out/ab_contig.log | grep "ms/apply\|flow_"
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -c 3000 gpurun_out/bench_f64.json; tail -5 gpurun_out/bench_f64.err
