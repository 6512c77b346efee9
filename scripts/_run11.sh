cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for st in 1 2; do for ch in 2 3 4; do CMBL_HOST_STREAMS=$st CMBL_HOST_CHUNKS=$ch timeout 120 python scripts/time_e2e.py f64 || echo "FAILED st=$st ch=$ch rc=$?"; done; done
for ch in 3 4; do CMBL_HOST_STREAMS=2 CMBL_HOST_CHUNKS=$ch timeout 120 python scripts/time_e2e.py f32 || echo "FAILED f32 ch=$ch"; done
} > gpurun_out/e2e3.log 2>&1
cat gpurun_out/e2e3.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
