#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 python scripts/repro_small.py > gpurun_out/repro.log 2>&1; tail -8 gpurun_out/repro.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scripts/repro_small.py > gpurun_out/repro_memcheck.log 2>&1; grep -m40 "Invalid\|at \|by thread\|Address\|ERROR SUMMARY\|ok" gpurun_out/repro_memcheck.log | head -60
