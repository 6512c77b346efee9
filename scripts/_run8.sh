cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ timeout 600 python scripts/time_map_joint.py f64 1024 P 8 3; timeout 600 python scripts/time_map_joint.py f32 1024 P 8 3; timeout 600 python scripts/time_map_joint.py f32 2048 IP 1 3; } > gpurun_out/map_joint.log 2>&1
cat gpurun_out/map_joint.log
