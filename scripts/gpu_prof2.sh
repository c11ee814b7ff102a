#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 24 --csv --log-file gpurun_out/launches2.csv python scripts/profile_step.py 3 > gpurun_out/launches2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_persistent -s 1 -c 1 -o gpurun_out/icp_persist -f python scripts/profile_step.py 2 > gpurun_out/ncu_icp2.log 2>&1
tail -3 gpurun_out/ncu_icp2.log
