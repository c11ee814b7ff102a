#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/nn.jsonl
for f in pose_refine_b200/variants/lib_nn_*.so; do
  PR_LIB=$PWD/$f timeout 300 python scripts/time_nn.py 512 3 ${CLUSTERS:-2,8} >> gpurun_out/nn.jsonl 2>> gpurun_out/nn.err
done
cat gpurun_out/nn.jsonl; tail -3 gpurun_out/nn.err
timeout 600 python -m pytest tests -m gpu -x -q -k "nn or c3 or pass_sums or refiner" -s 2>&1 | grep -E "passed|failed|FAILED|mismatch|^E  " | cut -c1-300 | head
