#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
PR_ICP_IMPL=pass timeout 300 python scripts/time_stages.py > gpurun_out/time_pass.json 2> gpurun_out/time_pass.err
timeout 300 python scripts/time_stages.py > gpurun_out/time_persistent.json 2> gpurun_out/time_persistent.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/time_pass.json gpurun_out/time_persistent.json; tail -3 gpurun_out/time_persistent.err
