#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "icp or pass_sums or correspond or solver or refiner or reference_arith or full_size" -s > gpurun_out/pytest_icp.log 2>&1
echo "pytest icp exit $?" ; grep -E "passed|failed|FAILED|mismatch|C2 full|fast vs exact" gpurun_out/pytest_icp.log | head -30
timeout 900 bash scripts/gpu_variants.sh
