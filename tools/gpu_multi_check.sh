#!/bin/bash
# 2+ GPU check: the multi-GPU pytest, then bench.py under torchrun (replica metric + config-4 sharded arm)
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_n$N.json") if l.startswith("{")][-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "n_gpus", d["n_gpus"])
print(json.dumps(d.get("sharded_esdf"), indent=1))
PY
