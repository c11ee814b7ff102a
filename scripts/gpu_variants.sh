#!/bin/bash
# time every variant library; a name like lib_x.so@N runs it with PR_ICP_WARPS=N
mkdir -p gpurun_out
: > gpurun_out/variants.jsonl
for f in pose_refine_b200/variants/lib_*.so; do
  for w in ${WARPS_LIST:-0}; do
    if [ "$w" = "0" ]; then unset PR_ICP_WARPS; else export PR_ICP_WARPS=$w; fi
    echo -n "{\"warps\": $w, \"r\": " >> gpurun_out/variants.jsonl
    PR_LIB=$PWD/$f timeout 40 python scripts/time_icp.py 512 8 >> gpurun_out/variants.jsonl 2>> gpurun_out/variants.err
    echo "}" >> gpurun_out/variants.jsonl
  done
done
cut -c1-260 gpurun_out/variants.jsonl; grep -iE "error|assert" gpurun_out/variants.err | sort | uniq -c | head -5
