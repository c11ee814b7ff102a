#!/bin/bash
# time every variant library for the cluster sizes in $CLUSTERS (default 1,2,4,8)
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for f in pose_refine_b200/variants/lib_*.so; do
  PR_LIB=$PWD/$f timeout 120 python scripts/time_icp.py 512 8 ${CLUSTERS:-1,2,4,8} >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
done
cut -c1-200 gpurun_out/variants.jsonl; grep -iE "error|assert" gpurun_out/variants.err | sort | uniq -c | head -5
