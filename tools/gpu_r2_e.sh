#!/bin/bash
# 8-GPU pass: sharded parity on real GPUs (2 ranks spawned; 8 ranks under torchrun), bench at N = 2, 4, 8 (pairs sharded +
# hypothesis sharding), cfg 5 at its stated size with the oracle subsample
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/e_smi.txt; nproc >> gpurun_out/e_smi.txt
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q > gpurun_out/e_pytest_shard2.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest_shard2.log; tail -3 gpurun_out/e_pytest_shard2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 -m pytest tests/test_gpu_sharded.py -m gpu -q -k under_torchrun > gpurun_out/e_pytest_shard8.log 2>&1; echo "rc=$?" >> gpurun_out/e_pytest_shard8.log; grep -E "passed|failed|rc=" gpurun_out/e_pytest_shard8.log | tail -4
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --skip-extras > gpurun_out/e_bench$n.json 2> gpurun_out/e_bench$n.err; tail -c 300 gpurun_out/e_bench$n.err
done
timeout 600 python bench.py --steps 10 --warmup 3 --skip-extras > gpurun_out/e_bench1.json 2> gpurun_out/e_bench1.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/eval_cfg5.py --pairs 5000 --oracle_every 25 --out gpurun_out/e_cfg5.json > gpurun_out/e_cfg5.log 2> gpurun_out/e_cfg5.err; tail -c 400 gpurun_out/e_cfg5.err
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/e_bench%d.json" % n).read().strip().splitlines()[-1])
        print(n, "pairs/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "clocks", d.get("clocks"))
        print("   hyp", json.dumps(d.get("hypothesis_sharding")))
    except Exception as e:
        print(n, "ERR", e)
try:
    print(open("gpurun_out/e_cfg5.json").read())
except Exception as e:
    print("cfg5 ERR", e)
PY
