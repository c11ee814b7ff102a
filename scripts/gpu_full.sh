#!/bin/bash
# full validation + bench + profiles (1 GPU)
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
# launch list of the bench command itself (2 steps)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 40 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"icp_persistent|raster_tile|bin_kernel" -s 4 -c 4 -o gpurun_out/step_kernels -f python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-1500 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.json
