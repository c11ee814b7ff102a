#!/bin/bash
mkdir -p gpurun_out
for f in pose_refine_b200/variants/lib_*.so; do
  PR_LIB=$PWD/$f timeout 120 python scripts/time_icp.py 512 10 ${CLUSTERS:-1,2,4,8}
done
