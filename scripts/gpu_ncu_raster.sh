#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_tile -s 1 -c 1 -f -o gpurun_out/raster_tile2 python scripts/time_step.py 512 1 > gpurun_out/ncu_raster.log 2>&1
tail -3 gpurun_out/ncu_raster.log
timeout 300 python scripts/time_step.py 512 10
