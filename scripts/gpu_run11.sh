#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
for name in base adj4 adj8 f2u4 f2u2; do
  if [ $name = base ]; then v="X=1"; else v="CMBL_B200_LIB=$PWD/scripts/ubench/libcmbl_$name.so"; fi
  for d in f64 f32; do for op in 0 1; do env $v timeout 120 python scripts/time_apply.py $d $op 2>&1 | grep "ms/apply\|flow_cols" | sed "s/^/$name /"; done; done
  N=512 env $v timeout 120 python scripts/time_apply.py f64 1 2>&1 | grep "ms/apply" | sed "s/^/$name /"
done > gpurun_out/ab_unr.log 2>&1
cat gpurun_out/ab_unr.log
