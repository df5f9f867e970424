#!/bin/bash
# ncu --set full of the tensor-core sweep, k_gen, k_kabsch (one launch each, cfg-3 pair)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/prof_score_tc_r2 python tools/ransac_once.py 0 3 > gpurun_out/c_ncu_tc.log 2>&1
tail -3 gpurun_out/c_ncu_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gen|k_kabsch" -s 2 -c 2 -f -o gpurun_out/prof_gen_kabsch_r2 python tools/ransac_once.py 0 3 > gpurun_out/c_ncu_gen.log 2>&1
tail -3 gpurun_out/c_ncu_gen.log
