#!/bin/bash
# A/B of two library variants on the refiner step: gpu_ab.sh <a> <b>
for rep in 1 2; do
for v in "$@"; do
  PR_LIB=$PWD/pose_refine_b200/variants/lib_$v.so timeout 120 python scripts/time_step.py 512 10 | tail -1 | cut -c1-330
done
done
