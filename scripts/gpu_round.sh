#!/bin/bash
# One GPU-box pass used at the end of a round: smoke, GPU parity tests, both bench arms and the two precisions.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 python bench.py > gpurun_out/bench_f64.json 2> gpurun_out/bench_f64.err
timeout 600 python bench.py --dtype f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err
python - <<'PY'
import json
for n in ("ref", "f64", "f32"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, round(d["value"], 3), d["unit"], "e2e", round(d["e2e"]["value"], 3), "roofline", (d.get("roofline") or {}).get("frac"), "cg", (d.get("cg") or {}).get("value"), d.get("clocks"))
    except Exception as e:
        print(n, "FAILED", e, open(f"gpurun_out/bench_{n}.err").read()[-400:])
PY
