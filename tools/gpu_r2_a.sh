#!/bin/bash
# round 2, first GPU pass: parity tests of the new kernels, sweep A/B, the existing bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x -k "score_tc" > gpurun_out/a_pytest_tc.log 2>&1
echo "rc_tc=$?" >> gpurun_out/a_pytest_tc.log
tail -25 gpurun_out/a_pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --deselect tests/test_gpu_sharded.py -k "not score_tc" > gpurun_out/a_pytest_rest.log 2>&1
echo "rc_rest=$?" >> gpurun_out/a_pytest_rest.log
tail -25 gpurun_out/a_pytest_rest.log
timeout 300 python tools/score_ab.py 2 > gpurun_out/a_score_ab.json 2> gpurun_out/a_score_ab.err
cat gpurun_out/a_score_ab.json | head -60
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/a_pytest_shard.log 2>&1
echo "rc_shard=$?" >> gpurun_out/a_pytest_shard.log
tail -15 gpurun_out/a_pytest_shard.log
