cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for ch in 2 3 4 5; do CMBL_HOST_CHUNKS=$ch timeout 300 python scripts/time_e2e.py f64; done
for ch in 2 3 4; do CMBL_HOST_CHUNKS=$ch timeout 300 python scripts/time_e2e.py f32; done
} > gpurun_out/e2e2.log 2>&1
cat gpurun_out/e2e2.log
