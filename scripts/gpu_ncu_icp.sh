#!/bin/bash
# one ncu --set full capture of the ICP kernel (C2 batch), source counters included
mkdir -p gpurun_out
LIBV=${1:-pose_refine_b200/variants/lib_base.so}
CL=${2:-2}
OUT=${3:-icp_hyp}
PR_LIB=$PWD/$LIBV timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_ -s 1 -c 1 -f -o gpurun_out/$OUT \
   python scripts/time_icp.py 512 1 $CL > gpurun_out/ncu_icp.log 2>&1
tail -3 gpurun_out/ncu_icp.log
