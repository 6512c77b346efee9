#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log)"; grep "SUMMARY\|ok " gpurun_out/sanitize_$tool.log | tail -4; grep -m5 "Error\|hazard\|Race" gpurun_out/sanitize_$tool.log
done
