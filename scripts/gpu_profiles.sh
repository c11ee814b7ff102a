#!/bin/bash
# the ncu evidence of round 2: full captures of the three hot kernels + the launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:icp_hyp -s 1 -c 1 -f -o gpurun_out/r02_icp_hyp python scripts/time_icp.py 512 1 0 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:raster_tile -s 1 -c 1 -f -o gpurun_out/r02_raster_tile python scripts/time_step.py 512 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:icp_hyp -s 1 -c 1 -f -o gpurun_out/r02_icp_nn python scripts/time_nn.py 512 1 0 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log
