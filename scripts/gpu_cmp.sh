#!/bin/bash
mkdir -p gpurun_out
PR_ICP_IMPL=pass python scripts/dump_results.py gpurun_out/res_pass.npy > gpurun_out/dump_pass.json 2>&1
python scripts/dump_results.py gpurun_out/res_persist.npy > gpurun_out/dump_persist.json 2>&1
python scripts/oracle_subset.py gpurun_out/res_oracle.npy 0,8,16,24,32,40,48,56,64,72,80,88,96,104,112,120,128,136,144,152,160,168,176,184,192,200,208,216,224,232,240,248,256,264,272,280,288,296,304,312,320,328,336,344,352,360,368,376,384,392,400,408,416,424,432,440,448,456,464,472,480,488,496,504 > gpurun_out/oracle_subset.log 2>&1
cat gpurun_out/dump_pass.json gpurun_out/dump_persist.json; tail -2 gpurun_out/oracle_subset.log
