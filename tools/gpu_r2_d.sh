#!/bin/bash
# 2-GPU pass: sharded parity test on two real GPUs, the bench line at N = 1 and N = 2 (pairs sharded + hypothesis sharding)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/d_smi.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/d_pytest_shard.log 2>&1; echo "rc=$?" >> gpurun_out/d_pytest_shard.log
tail -5 gpurun_out/d_pytest_shard.log
timeout 900 python bench.py > gpurun_out/d_bench1.json 2> gpurun_out/d_bench1.err; tail -c 600 gpurun_out/d_bench1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/d_bench2.json 2> gpurun_out/d_bench2.err; tail -c 600 gpurun_out/d_bench2.err
python - <<'PY'
import json
for f in ("gpurun_out/d_bench1.json", "gpurun_out/d_bench2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], d.get("e2e_numpy", {}).get("value"), d["roofline"]["avg_launch_ms"], d["roofline"]["issue_frac"])
        print(" hyp", json.dumps(d.get("hypothesis_sharding")))
        print(" fr", json.dumps(d.get("fr_e2e")))
        print(" faithful", json.dumps(d.get("reference_faithful")))
        print(" mnn", json.dumps(d.get("mnn_match", {}).get("fractions_of_tensor_peak")), d.get("mnn_match", {}).get("ms"))
        print(" other", json.dumps(d.get("other_regime")))
        print(" cpu", json.dumps(d.get("cpu_baseline")), json.dumps(d.get("speedup_vs_cpu")))
    except Exception as e:
        print(f, "ERR", e)
PY
