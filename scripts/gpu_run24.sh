#!/bin/bash
# two B200s with the final kernels: the C-ABI communicator on real ranks, and the bench line at N=2
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/comm_2gpu.py > gpurun_out/comm_2gpu.log 2>&1; tail -6 gpurun_out/comm_2gpu.log
timeout 900 $TR bench.py --gpus 2 --skip cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_n2.json"))
print(d["n_gpus"], round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "cg", round(d["cg"]["value"], 1), "mj", d["map_joint"]["value"], d["map_joint"]["collectives"], "hmc", round(d["hmc"]["value"], 1))
PY
