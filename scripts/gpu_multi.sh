#!/bin/bash
# multi-GPU check: the C++ host demo (one thread per GPU) and the bench line at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_cpp_dropin.py -m gpu -x -q -k multi_gpu -s 2>&1 | grep -E "ranks|passed|failed|Error|error" | head
bash scripts/gpu_bench.sh $N
