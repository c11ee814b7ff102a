#!/bin/bash
# launch list of one warm step + full capture of the ICP pass kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
# step = 1 memset + 4 render + 3 count + 1 fill + 1 plan + 31 pass = 41 launches (+ scene setup before)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 3 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_pass -s 40 -c 2 -o gpurun_out/icp_pass -f python scripts/profile_step.py 2 > gpurun_out/ncu_icp.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/launches.log; tail -3 gpurun_out/ncu_icp.log; ls -la gpurun_out
