#!/bin/bash
# eight B200s with the final kernels: the bench line at N=8 (and N=4)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for n in 8 4; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n"
  timeout 900 $TR bench.py --gpus $n --skip cpu > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; tail -2 gpurun_out/bench_n$n.err
  python - $n <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_n{sys.argv[1]}.json"))
print(d["n_gpus"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "copies", d["e2e"].get("copies_alone_ms_per_step"), "cg", round(d["cg"]["value"], 1), "mj", d["map_joint"]["value"], "hmc", round(d["hmc"]["value"], 1))
PY
done
