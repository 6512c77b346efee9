cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{ timeout 600 python scripts/time_map_joint.py f64 1024 P 8 1; timeout 600 python scripts/time_map_joint.py f32 1024 P 8 1; } > gpurun_out/map_joint2.log 2>&1
grep -E "gradient|MAP_joint|HMC" gpurun_out/map_joint2.log
