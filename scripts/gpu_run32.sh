#!/bin/bash
# bench lines with the standalone-transform section (same binary as the final pass)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bench_contract.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err; tail -1 gpurun_out/bench_f64.err
timeout 900 python bench.py --dtype f32 --skip cpu > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; tail -1 gpurun_out/bench_f32.err
python - <<'PY'
import json
for n in ("f64", "f32"):
    d = json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "traffic", d["roofline"]["traffic"], "cg", round(d["cg"]["value"], 1), "mj", round(d["map_joint"]["value"], 3), "hmc", round(d["hmc"]["value"], 1), "fft", d["fft"])
PY
