#!/bin/bash
# evidence pass: full GPU suite, bench line, ncu captures of the final kernels, launch list, pipeline traces, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/g_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/g_pytest.log; tail -4 gpurun_out/g_pytest.log
timeout 900 python bench.py > gpurun_out/g_bench1.json 2> gpurun_out/g_bench1.err; tail -c 300 gpurun_out/g_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/prof_score_tc_r2_final python tools/ransac_once.py 0 3 > gpurun_out/g_ncu_tc.log 2>&1; tail -1 gpurun_out/g_ncu_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_nn_tc|k_rerank" -s 4 -c 2 -f -o gpurun_out/prof_nn_tc_r2 python tools/match_once.py 0 3 > gpurun_out/g_ncu_nn.log 2>&1; tail -1 gpurun_out/g_ncu_nn.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/g_ncu_bench.log 2>&1; tail -1 gpurun_out/g_ncu_bench.log
python tools/tcs_trace.py 32 1 > gpurun_out/trace_tn32_elc.txt 2>&1; python tools/tcs_trace.py 32 0 > gpurun_out/trace_tn32_noelc.txt 2>&1; grep "^# " gpurun_out/trace_tn32_elc.txt | tail -4
timeout 900 compute-sanitizer --tool memcheck python tests/sanitize_workload.py > gpurun_out/g_san_mem.log 2>&1; tail -2 gpurun_out/g_san_mem.log
timeout 900 compute-sanitizer --tool racecheck python tests/sanitize_workload.py > gpurun_out/g_san_race.log 2>&1; tail -2 gpurun_out/g_san_race.log
