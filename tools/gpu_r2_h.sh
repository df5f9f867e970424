#!/bin/bash
# evidence pass of the final build: full GPU suite, bench line, reference arm, ncu captures, launch list, pipeline traces, sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/h_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/h_pytest.log; tail -4 gpurun_out/h_pytest.log
timeout 900 python bench.py > gpurun_out/h_bench1.json 2> gpurun_out/h_bench1.err; tail -c 300 gpurun_out/h_bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.json 2> gpurun_out/h_bench_ref.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_score_tc -s 1 -c 1 -f -o gpurun_out/prof_score_tc_r2_tn64 python tools/ransac_once.py 0 3 > gpurun_out/h_ncu_tc.log 2>&1; tail -1 gpurun_out/h_ncu_tc.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_icp_eval|k_grid_insert|k_kabsch_weighted|k_seed_end" -c 6 -f -o gpurun_out/prof_f4_r2 python tools/f4_once.py > gpurun_out/h_ncu_f4.log 2>&1; tail -1 gpurun_out/h_ncu_f4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 --skip-extras > gpurun_out/h_ncu_bench.log 2>&1; tail -1 gpurun_out/h_ncu_bench.log
timeout 300 python tools/tcs_trace.py 64 1 > gpurun_out/trace_tn64_elc.txt 2>&1; timeout 300 python tools/tcs_trace.py 64 0 > gpurun_out/trace_tn64_noelc.txt 2>&1; grep "^# " gpurun_out/trace_tn64_elc.txt | tail -4
timeout 600 compute-sanitizer --tool memcheck python tests/sanitize_workload.py > gpurun_out/h_san_mem.log 2>&1; tail -2 gpurun_out/h_san_mem.log
timeout 600 compute-sanitizer --tool racecheck python tests/sanitize_workload.py > gpurun_out/h_san_race.log 2>&1; tail -2 gpurun_out/h_san_race.log
for w in 1 2 4 8; do timeout 120 python tools/pair_breakdown.py 1 $w > gpurun_out/h_breakdown_w$w.json 2>&1; done; cat gpurun_out/h_breakdown_w8.json; timeout 120 python tools/pdl_ab.py > gpurun_out/h_pdl_ab.json 2>&1
