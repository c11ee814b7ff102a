#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"raster_tile|bin_smem|vertex_kernel|cloud_count|cloud_fill" -s 5 -c 6 -o gpurun_out/raster_kernels -f python scripts/profile_step.py 2 > gpurun_out/ncu_raster.log 2>&1
tail -2 gpurun_out/ncu_raster.log
