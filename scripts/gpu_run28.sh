#!/bin/bash
# what the memory system delivers for the column kernel's access pattern without any of its arithmetic (scripts/ubench/stream_pattern.cu)
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
timeout 300 scripts/ubench/stream_pattern > gpurun_out/r02_stream_pattern.log 2>&1
cat gpurun_out/r02_stream_pattern.log
