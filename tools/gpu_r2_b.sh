#!/bin/bash
# round 2, GPU pass B: all GPU tests after the nps==1 fix, the bench line, launch list, ncu of the tensor-core sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 > gpurun_out/b_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/b_pytest.log
tail -30 gpurun_out/b_pytest.log
timeout 300 python tools/score_ab.py 2 > gpurun_out/b_score_ab.json 2> gpurun_out/b_score_ab.err
timeout 600 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -c 3000 gpurun_out/b_bench.json; tail -5 gpurun_out/b_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 2 -c 1 -f -o gpurun_out/prof_score_tc_r2 python tools/ransac_once.py 0 2 > gpurun_out/b_ncu_tc.log 2>&1
tail -3 gpurun_out/b_ncu_tc.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/b_ncu_bench.log 2>&1
tail -3 gpurun_out/b_ncu_bench.log
