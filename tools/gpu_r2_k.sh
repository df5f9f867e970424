#!/bin/bash
# 8-GPU pass of the final build: sharded parity under torchrun (8 ranks), bench at N = 8 and 4 (pairs sharded + hypothesis
# sharding), cfg 5 at its stated size (FR + ICP per pair, oracle subsample)
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/k_smi.txt; nproc >> gpurun_out/k_smi.txt
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 -m pytest tests/test_gpu_sharded.py -m gpu -q -k under_torchrun > gpurun_out/k_pytest_shard8.log 2>&1; echo "rc=$?" >> gpurun_out/k_pytest_shard8.log; grep -E "passed|failed|rc=" gpurun_out/k_pytest_shard8.log | tail -3
for n in 8 4; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --skip-extras > gpurun_out/k_bench$n.json 2> gpurun_out/k_bench$n.err; tail -c 200 gpurun_out/k_bench$n.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/eval_cfg5.py --pairs 5000 --oracle_every 50 --out gpurun_out/k_cfg5.json > gpurun_out/k_cfg5.log 2> gpurun_out/k_cfg5.err; tail -c 300 gpurun_out/k_cfg5.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open("gpurun_out/k_bench%d.json" % n).read().strip().splitlines()[-1])
        print(n, "pairs/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "per-pair-call", round(d["e2e_per_pair_call"]["value"], 1), "clocks", d.get("clocks"))
        print("   hyp", json.dumps(d.get("hypothesis_sharding")))
    except Exception as e:
        print(n, "ERR", e)
try:
    print(open("gpurun_out/k_cfg5.json").read())
except Exception as e:
    print("cfg5 ERR", e)
PY
