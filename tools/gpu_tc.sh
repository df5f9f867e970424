#!/bin/bash
# tensor-sweep iteration: parity tests that exercise the sweep, the A/B timing, one ncu --set full capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_score_tc.py tests/test_gpu_ransac.py tests/test_gpu_random_sweep.py -m gpu -q -x > gpurun_out/tc_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/tc_pytest.log
tail -15 gpurun_out/tc_pytest.log
timeout 300 python tools/score_ab.py 2 0 > gpurun_out/tc_score_ab.json 2> gpurun_out/tc_score_ab.err
cat gpurun_out/tc_score_ab.json | tr -d ' \n' ; echo; tail -3 gpurun_out/tc_score_ab.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/prof_score_tc_r2 python tools/ransac_once.py 0 3 > gpurun_out/tc_ncu.log 2>&1
tail -2 gpurun_out/tc_ncu.log
