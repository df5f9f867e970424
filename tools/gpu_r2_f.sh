#!/bin/bash
# f4 pass: new GPU parity tests (device ICP, seed scoring), pair breakdown for 1 rank and for rank 0 of 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_f4.py tests/test_gpu_fr.py -m gpu -x -q -s > gpurun_out/f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest.log; tail -15 gpurun_out/f_pytest.log
for w in 1 2 4 8; do timeout 120 python tools/pair_breakdown.py 1 $w > gpurun_out/f_breakdown_w$w.json 2>&1; cat gpurun_out/f_breakdown_w$w.json; done
