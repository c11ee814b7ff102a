#!/bin/bash
# round 2, first GPU visit: parity tests of the new ICP driver, variant timings, a bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "icp or pass_sums or correspond or solver or refiner or reference_arith or full_size" -s > gpurun_out/pytest_icp.log 2>&1
echo "pytest icp exit $?" ; tail -25 gpurun_out/pytest_icp.log
timeout 600 bash scripts/gpu_variants.sh
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json
