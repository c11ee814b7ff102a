#!/bin/bash
mkdir -p gpurun_out
CLUSTERS=2,4 timeout 400 bash scripts/gpu_quick_icp.sh 2>&1 | grep -E "icp_ms|rror"
