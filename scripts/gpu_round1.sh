#!/bin/bash
# first GPU pass: tests, smoke, short bench, sanitizer on smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/sanitizer.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/sanitizer.log
