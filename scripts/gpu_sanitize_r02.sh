#!/bin/bash
# round 2: compute-sanitizer over the new kernels (cluster / DSMEM ICP driver, hash grid + queue, folded raster outputs)
mkdir -p gpurun_out
run() { # name tool test-filter
  timeout 1500 compute-sanitizer --tool $2 --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "$3" > gpurun_out/san_$1.log 2>&1
  echo "$1 ($2) rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/san_$1.log | tail -3
}
run mem_smoke memcheck "refiner_end_to_end"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/san_mem_entry.log 2>&1; echo "mem_entry rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/san_mem_entry.log | tail -1
run mem_icp memcheck "icp_projective_fixture or icp_nn or pass_sums or pose_renderer or capacity"
run mem_corr memcheck "correspondences"
run race_icp racecheck "icp_projective_fixture or refiner_end_to_end"
run sync_icp synccheck "icp_projective_fixture or refiner_end_to_end"
