#!/bin/bash
# quick iteration: gpu tests (ICP-related only unless FULL=1), stage timing, launch list, optional ncu
mkdir -p gpurun_out
if [ "$FULL" = "1" ]; then SEL=""; else SEL="icp or refiner or pcd2ab or full_size"; fi
timeout 900 python -m pytest tests -m gpu -q ${SEL:+-k "$SEL"} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/time_stages.py 512 6 > gpurun_out/time_persistent.json 2> gpurun_out/time_persistent.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 13 --csv --log-file gpurun_out/launches2.csv python scripts/profile_step.py 3 > gpurun_out/launches2.log 2>&1
if [ "$NCU" = "1" ]; then timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_persistent -s 1 -c 1 -o gpurun_out/icp_persist -f python scripts/profile_step.py 2 > gpurun_out/ncu_icp2.log 2>&1; fi
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/time_persistent.json; tail -3 gpurun_out/time_persistent.err; grep -E "icp_persistent" gpurun_out/launches2.csv | awk -F'","' '{print $5, $(NF-1), $NF}' | head -3
