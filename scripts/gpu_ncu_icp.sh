#!/bin/bash
# one ncu --set full capture of the ICP kernel (C2 batch), source counters included
mkdir -p gpurun_out
LIBV=${1:-pose_refine_b200/variants/lib_w4b4i4.so}
CL=${2:-2}
PR_LIB=$PWD/$LIBV timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_hyp -s 1 -c 1 -f -o gpurun_out/icp_hyp \
   python scripts/time_icp.py 512 1 $CL > gpurun_out/ncu_icp.log 2>&1
tail -5 gpurun_out/ncu_icp.log
