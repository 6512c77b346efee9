#!/bin/bash
# eight B200s: BASELINE configs 4 and 5 at their named scale (map_joint: 8 items sharded over 8 GPUs with NCCL all-reduces; hmc: 64 chains), e2e host-path scaling
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
nvidia-smi topo -m > gpurun_out/topo_n8.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/lscpu_n8.txt
timeout 900 $TR bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 600 gpurun_out/bench_n8.json; tail -3 gpurun_out/bench_n8.err
CMBL_BENCH_NUMA=off timeout 600 $TR bench.py --gpus 8 --skip cg,map_joint,hmc,cpu,other > gpurun_out/bench_n8_nonuma.json 2> gpurun_out/bench_n8_nonuma.err; tail -2 gpurun_out/bench_n8_nonuma.err
python - <<'PY'
import json
for n in ("n8", "n8_nonuma"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json")); e = d["e2e"]
        print(n, "value", round(d["value"], 1), "e2e", round(e["value"], 1), "sync", round(e["synchronous"]["value"], 1), "copies_ms", round(e["copies_alone_ms_per_step"], 2), e["host_memory_policy"],
              "cg", (d.get("cg") or {}).get("value"), "mj", (d.get("map_joint") or {}).get("value"), "hmc", (d.get("hmc") or {}).get("value"))
    except Exception as ex:
        print(n, "FAILED", ex)
PY
cat gpurun_out/lscpu_n8.txt
