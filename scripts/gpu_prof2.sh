#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:icp_persistent -s 1 -c 1 -o gpurun_out/icp_persist -f python scripts/profile_step.py 2 > gpurun_out/ncu_icp2.log 2>&1
tail -2 gpurun_out/ncu_icp2.log
