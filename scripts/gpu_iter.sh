#!/bin/bash
# quick iteration: gpu tests (ICP-related only unless FULL=1), ICP timing, optional ncu
mkdir -p gpurun_out
if [ "$FULL" = "1" ]; then SEL=""; else SEL="icp or refiner or pcd2ab or full_size"; fi
timeout -k 10 600 python -m pytest tests -m gpu -q ${SEL:+-k "$SEL"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 200 python scripts/time_icp.py 512 8 > gpurun_out/time_icp.json 2> gpurun_out/time_icp.err
PR_ICP_IMPL=warp timeout -k 10 200 python scripts/time_icp.py 512 8 >> gpurun_out/time_icp.json 2>> gpurun_out/time_icp.err
timeout -k 10 300 python scripts/time_stages.py 512 6 > gpurun_out/time_persistent.json 2> gpurun_out/time_persistent.err
if [ "$NCU" = "1" ]; then timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"icp_p" -s 1 -c 1 -o gpurun_out/icp_persist -f python scripts/profile_step.py 2 > gpurun_out/ncu_icp2.log 2>&1; fi
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/time_icp.json; tail -2 gpurun_out/time_icp.err; cut -c1-400 gpurun_out/time_persistent.json
