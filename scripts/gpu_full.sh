#!/bin/bash
# full validation + bench + profiles (1 GPU); artefacts land in gpurun_out/ (copied to profiles/ by scripts/collect_profiles.py)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout -k 10 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout -k 10 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout -k 10 600 python scripts/time_ref_cuda.py 512 1,8,16 > gpurun_out/ref_cuda.json 2> gpurun_out/ref_cuda.err
# launch list of the bench command itself (2 steps)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 25 -c 40 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"icp_persistent|raster_tile|bin_smem|cloud_fill_tiles|vertex_kernel" -s 5 -c 6 -o gpurun_out/step_kernels -f python scripts/profile_step.py 2 > gpurun_out/ncu_step.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-1800 gpurun_out/bench.json; tail -2 gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/ref_cuda.json
