#!/bin/bash
# variants timing first (cheap), then the GPU test-suite with the default library
mkdir -p gpurun_out
bash scripts/gpu_variants.sh > gpurun_out/variants.log 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/variants.jsonl | cut -c1-200; tail -5 gpurun_out/pytest_gpu.log
