#!/bin/bash
mkdir -p gpurun_out
CLUSTERS=2,4,8 timeout 900 bash scripts/gpu_variants.sh
timeout 600 python -m pytest tests -m gpu -x -q -k "full_size" -s 2>&1 | grep -E "passed|failed|C2 full|statistics|^E  " | cut -c1-300
