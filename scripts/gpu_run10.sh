#!/bin/bash
# two B200s: the C-ABI communicator on real ranks, and the bench line at N=2 (map_joint with its all-reduces inside the timed region)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR scripts/comm_2gpu.py > gpurun_out/comm_2gpu.log 2>&1; tail -6 gpurun_out/comm_2gpu.log
timeout 900 $TR bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 2500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; cat gpurun_out/bench_n2_ref.json | cut -c1-600
