#!/bin/bash
# round-end evidence on one GPU: bench line, reference arm, ncu captures of the three hot kernels, launch list
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm exit $?"; tail -c 600 gpurun_out/bench_reference_arm.json
bash scripts/gpu_profiles.sh
