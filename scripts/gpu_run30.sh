#!/bin/bash
# ncu --set full captures of the kernels added late in round 2: irfft2's persistent column kernel, the stage kernels at transform length 2048
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
SHA=$(sha256sum cmblensing.jl_b200/libcmbl_b200.so | cut -c1-16)
cap() { local name=$1 desc=$2 s=$3 c=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -s $s -c $c -o gpurun_out/$name "$@" > gpurun_out/$name.log 2>&1
  { echo "ncu --set full --clock-control none --import-source on -s $s -c $c  $*   ($desc)"; echo "libcmbl_b200.so sha256[:16] = $SHA"
    python scripts/ncu_summary.py gpurun_out/$name.ncu-rep; python scripts/ncu_hot.py gpurun_out/$name.ncu-rep 12; } > gpurun_out/$name.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep; head -5 gpurun_out/$name.txt | cut -c1-200; }
cap r02_ncu_irfft2_f64 "inverse row pass + persistent column kernel of the irfft2 inside precompute, Nside=1024, 40 planes" 8 2 python scripts/ncu_target.py f64 adj
N=2048 NB=2 NPOL=3 cap r02_ncu_flow_2048_f64 "row + column kernel of a middle RK4 stage at Nside=2048 IQU batch 2" 12 2 env N=2048 NB=2 NPOL=3 python scripts/ncu_target.py f64 fwd
