#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv > gpurun_out/smi0.txt
timeout 300 python scripts/time_stages.py 512 8 > gpurun_out/time_persistent.json 2> gpurun_out/time_persistent.err
PR_ICP_IMPL=pass timeout 300 python scripts/time_stages.py 512 8 > gpurun_out/time_pass.json 2> gpurun_out/time_pass.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 30 --csv --log-file gpurun_out/launches2.csv python scripts/profile_step.py 3 > gpurun_out/launches2.log 2>&1
cat gpurun_out/time_persistent.json gpurun_out/time_pass.json; tail -3 gpurun_out/time_persistent.err; grep -E "icp_persistent|raster_tile" gpurun_out/launches2.csv | cut -d, -f5,13-15 | head
