cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
for dt in f64 f32; do for op in 0 1; do NB=1 NPOL=1 timeout 300 python scripts/time_apply.py $dt $op; done; done
NB=1 NPOL=1 CMBL_PDL=1 timeout 300 python scripts/time_apply.py f64 0
NB=2 NPOL=1 timeout 300 python scripts/time_apply.py f64 0
} > gpurun_out/cfg2.log 2>&1
cat gpurun_out/cfg2.log
