#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "icp or pass_sums or correspond or solver or refiner or reference_arith or full_size or c3 or c4" -s > gpurun_out/pytest_icp.log 2>&1
echo "pytest icp exit $?" ; grep -E "passed|failed|FAILED|mismatch|C2 full|fast vs exact|^E  " gpurun_out/pytest_icp.log | cut -c1-250 | head -40
CLUSTERS=1,2,4 timeout 900 bash scripts/gpu_variants.sh
PR_LIB=$PWD/pose_refine_b200/variants/lib_base.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_hyp -s 1 -c 1 -f -o gpurun_out/icp_hyp2 python scripts/time_icp.py 512 1 2 > gpurun_out/ncu_icp.log 2>&1
tail -3 gpurun_out/ncu_icp.log
