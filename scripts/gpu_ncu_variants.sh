#!/bin/bash
# full ncu capture (with source) of the persistent ICP kernel for the named variant libraries
mkdir -p gpurun_out
for name in "$@"; do
  PR_LIB=$PWD/pose_refine_b200/variants/lib_$name.so timeout -k 10 600 ncu --set full --clock-control none --import-source on \
    -k regex:icp_persistent -s 1 -c 1 -o gpurun_out/icp_$name -f python scripts/time_icp.py 512 1 > gpurun_out/ncu_$name.log 2>&1
  tail -2 gpurun_out/ncu_$name.log
done
ls -la gpurun_out/*.ncu-rep
