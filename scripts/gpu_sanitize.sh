#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over the smoke run and the small parity tests of the new kernels
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_smoke.log 2>&1; echo "memcheck smoke rc=$?" | tee -a gpurun_out/sanitizer_smoke.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "cloud_fused or clustered or scene_nn_build or icp_projective_fixture or refiner_end_to_end" > gpurun_out/sanitizer_tests.log 2>&1; echo "memcheck tests rc=$?" | tee -a gpurun_out/sanitizer_tests.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/racecheck_smoke.log 2>&1; echo "racecheck smoke rc=$?" | tee -a gpurun_out/racecheck_smoke.log
tail -4 gpurun_out/sanitizer_smoke.log; tail -6 gpurun_out/sanitizer_tests.log; tail -4 gpurun_out/racecheck_smoke.log
