timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_esdf.py 2>&1 | grep -v Warning | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 tools/shard_esdf.py 2>&1 | grep -v Warning | tail -3
